"""Array form of the graph half of the oracle (TEST INFRASTRUCTURE) for graphs too large for NetworkX.

Same semantics as ``oracle/ref_search.py`` (which is pinned against the unmodified reference), restated over the
CSR arrays of ``ah_rag_b200.csr`` / ``bench_workloads.make_c5_graph`` with dense node ids:
  anchor_rerank  -- hierarchical_graph.py:719-789 with the default search params (:37-47)
  expand         -- environment.py:393-480
``tests/test_oracle.py::test_array_oracle_equals_networkx_oracle`` proves the two forms equal on the golden graphs.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

PART_OUT, PART_IN, BEL_OUT, BEL_IN, REL_OUT, REL_IN = range(6)
DEFAULTS = {"alpha": 0.6, "beta": 0.2, "gamma": 0.1, "delta": 0.1, "boost": (0.0, 1.0, 0.0, 0.0)}


def nbrs(g: Dict, slot: int, v: int) -> List[int]:
    rp = g["row_ptr"][slot]
    return [int(x) for x in g["col_idx"][slot][rp[v]:rp[v + 1]]]


def anchor_rerank(g: Dict, seed_nodes: Sequence[int], seed_scores: Sequence[float], top_k: int, member_top_m: int = 5,
                  params: Dict = DEFAULTS) -> List[Tuple[int, float, float]]:
    E: Dict[int, float] = {}
    for v, s in zip(seed_nodes, seed_scores):
        v = int(v)
        if v < 0:
            continue
        sem = 1.0 - float(np.float32(1.0) - np.float32(s))  # :723 on hnswlib's fp32 distance
        E[v] = sem  # overwrite keeps the slot (:732)
        t = int(g["node_type"][v])
        if t == 0:
            for p in nbrs(g, BEL_OUT, v)[:2]:
                if p not in E:
                    E[p] = sem * 0.9
        elif t == 1:
            for c in nbrs(g, BEL_IN, v)[:member_top_m]:
                if c not in E:
                    E[c] = sem * 0.85
    res = []
    for node, sem in E.items():
        judge, conf, t = g["judge"][node], g["conf"][node], int(g["node_type"][node])
        jt = float(g["judge_term"][node]) if judge == judge else 0.0
        ct = float(conf) / 10.0 if conf == conf else 0.0
        score = params["alpha"] * sem + params["beta"] * jt + params["gamma"] * ct + params["delta"] * params["boost"][t]
        res.append((node, round(sem, 4), round(float(score), 4)))
    res.sort(key=lambda x: x[2], reverse=True)  # stable, on the rounded score (:781)
    return res[:top_k]


def expand(g: Dict, op: int, ids: Sequence[int], limit: int) -> List[int]:
    X: List[int] = []
    n = len(g["node_type"])
    for v in ids:
        v = int(v)
        if v < 0 or v >= n:
            continue
        if op == 0:
            X += nbrs(g, BEL_OUT, v)
        elif op == 1:
            X += nbrs(g, BEL_IN, v)
        else:
            X += nbrs(g, REL_OUT, v)
            X += nbrs(g, REL_IN, v)
            if int(g["node_type"][v]) == 0:
                for h in nbrs(g, PART_OUT, v):
                    X.append(h)
                    X += [u for u in nbrs(g, PART_IN, h) if u != v]
        if len(X) >= limit:
            break
    return list(dict.fromkeys(X))[:limit]
