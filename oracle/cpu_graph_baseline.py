"""CPU baseline for the GRAPH half of the path (TEST / BENCH INFRASTRUCTURE).

Two forms, both sequential -- one environment after the other on one core, the shape of the reference's rollout loop
(/root/reference/src/ah_rag/agent/policy_ppo.py:163-205, "sequential roll" :144):

  networkx_loop   the reference's algorithm as ``oracle/ref_search.py`` restates it (pinned against the unmodified
                  reference), over a NetworkX DiGraph exactly as the reference holds its graph, on a REDUCED c5 graph
                  (NetworkX needs ~1 KB per edge: the full 500M-edge graph does not fit);
  array_loop      the same semantics over the CSR arrays (``oracle/array_graph.py``), at the FULL c5 size.

Per environment: anchor expansion + rerank of 50 seeds (hierarchical_graph.py:719-789), then one expand_* call on the
first ids of the result (environment.py:393-480) -- the two hops bench.py times on the GPU.
"""
from __future__ import annotations

import json
import time
from typing import Dict, Sequence

import numpy as np

from . import array_graph, ref_search

_TYPES = {0: "entity", 1: "summary", 2: "hyperedge"}
_EDGE = {0: "participates_in", 2: "belongs_to", 4: "related_to"}  # the OUT slot of each edge type
_OPS = ("parents", "children", "related")


def networkx_graph(host: Dict) -> "ref_search.OracleGraph":
    """CSR arrays (dense ids) -> the NetworkX DiGraph the reference would have loaded (node ids = str(dense id))."""
    import networkx as nx
    G = nx.DiGraph()
    nt, lv, judge, conf = host["node_type"], host["level"], host["judge"], host["conf"]
    for v in range(len(nt)):
        d = {"node_type": _TYPES.get(int(nt[v]))}
        if lv[v]:
            d["level"] = int(lv[v])
        if judge[v] == judge[v]:
            d["judge_scores"] = json.dumps({"overall": float(judge[v])})
        if conf[v] == conf[v]:
            d["confidence"] = float(conf[v])
        G.add_node(str(v), **d)
    for slot, et in _EDGE.items():
        rp, ci = host["row_ptr"][slot], host["col_idx"][slot]
        for v in np.nonzero(np.diff(rp))[0]:
            for w in ci[rp[v]:rp[v + 1]]:
                G.add_edge(str(int(v)), str(int(w)), edge_type=et)
    G.graph["search_params"] = {}
    og = object.__new__(ref_search.OracleGraph)
    og.G = G
    og._fallback_sp = {}
    return og


def networkx_loop(host: Dict, seeds: np.ndarray, scores: np.ndarray, ops: Sequence[int], top_k: int, member_top_m: int,
                  limit: int, budget_s: float) -> Dict[str, object]:
    t0 = time.perf_counter()
    og = networkx_graph(host)
    build_s = time.perf_counter() - t0
    prm = ref_search.resolve_params(og, member_top_m=member_top_m)
    n, t0 = 0, time.perf_counter()
    while n < len(seeds) and time.perf_counter() - t0 < budget_s:
        sd = ref_search.seeds_from_topk([str(int(v)) for v in seeds[n]], scores[n])
        out = ref_search.anchor_rerank(og, sd, top_k, prm)
        take = 1 if ops[n] == 2 else 2
        ref_search.expand(og, _OPS[int(ops[n])], [r["node_id"] for r in out["reranked"][:take]], limit)
        n += 1
    dt = time.perf_counter() - t0
    return {"envs_per_s": n / dt, "envs": n, "seconds": dt, "graph_build_s": build_s,
            "nodes": og.G.number_of_nodes(), "edges": og.G.number_of_edges()}


def array_loop(host: Dict, seeds: np.ndarray, scores: np.ndarray, ops: Sequence[int], top_k: int, member_top_m: int,
               limit: int, budget_s: float) -> Dict[str, object]:
    n, t0 = 0, time.perf_counter()
    while n < len(seeds) and time.perf_counter() - t0 < budget_s:
        out = array_graph.anchor_rerank(host, seeds[n], scores[n], top_k, member_top_m)
        take = 1 if ops[n] == 2 else 2
        array_graph.expand(host, int(ops[n]), [r[0] for r in out[:take]], limit)
        n += 1
    dt = time.perf_counter() - t0
    return {"envs_per_s": n / dt, "envs": n, "seconds": dt}
