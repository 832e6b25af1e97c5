"""Graph -> CSR exporter (host side of the boundary; numpy only, no CUDA).

Dense node id = position in ``G.nodes`` order of the LOADED graph.  For each (edge type, direction) one
CSR whose neighbour order is the adjacency iteration order NetworkX gives after
``HierarchicalGraph.load`` (hierarchical_graph.py:196) -- that order, not a sorted one, is what the
reference's expansion semantics depend on (hierarchical_graph.py:441-453, environment.py:401-467).
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional

import numpy as np

EDGE_TYPES = ("participates_in", "belongs_to", "related_to")  # slot = 2 * type + (0 out | 1 in)
NODE_TYPE_CODE = {"entity": 0, "summary": 1, "hyperedge": 2}
NODE_TYPE_NAME = {0: "entity", 1: "summary", 2: "hyperedge", 3: None}


def judge_overall(d: Dict[str, Any]) -> Optional[float]:
    """``judge`` of a node exactly as hierarchical_graph.py:746-752 / environment.py:99-105 derive it."""
    js = d.get("judge_scores")
    if not js:
        return None
    try:
        if isinstance(js, str):
            js = json.loads(js)
        return float(js.get("overall", 0.0))
    except Exception:
        return None


def _csr_from_lists(n: int, adj: List[List[int]]) -> (np.ndarray, np.ndarray):
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    for i, a in enumerate(adj):
        row_ptr[i + 1] = row_ptr[i] + len(a)
    col = np.fromiter((v for a in adj for v in a), dtype=np.int32, count=int(row_ptr[-1]))
    return row_ptr, col


def csr_from_networkx(G) -> Dict[str, Any]:
    """Export a loaded ``nx.DiGraph`` to the arrays ``ahr_graph_create`` takes."""
    node_ids = list(G.nodes)
    dense = {nid: i for i, nid in enumerate(node_ids)}
    n = len(node_ids)
    lists = [[[] for _ in range(n)] for _ in range(6)]
    tcode = {t: i for i, t in enumerate(EDGE_TYPES)}
    for u in node_ids:
        ui = dense[u]
        for _, v, d in G.out_edges(u, data=True):  # adjacency order of the loaded graph
            t = tcode.get(d.get("edge_type"))
            if t is not None:
                lists[2 * t][ui].append(dense[v])
        for w, _, d in G.in_edges(u, data=True):
            t = tcode.get(d.get("edge_type"))
            if t is not None:
                lists[2 * t + 1][ui].append(dense[w])
    row_ptr, col_idx = [], []
    for s in range(6):
        rp, ci = _csr_from_lists(n, lists[s])
        row_ptr.append(rp)
        col_idx.append(ci)

    node_type = np.full(n, 3, dtype=np.uint8)
    level = np.zeros(n, dtype=np.uint8)
    judge = np.full(n, np.nan, dtype=np.float64)
    judge_term = np.zeros(n, dtype=np.float64)
    conf = np.full(n, np.nan, dtype=np.float64)
    for i, nid in enumerate(node_ids):
        d = G.nodes[nid]
        node_type[i] = NODE_TYPE_CODE.get(d.get("node_type"), 3)
        lv = d.get("level")
        if isinstance(lv, (int, float)) and not isinstance(lv, bool) and 0 <= lv < 256:
            level[i] = int(lv)
        j = judge_overall(d)
        if j is not None:
            judge[i] = j
            # same expression, same library call as hierarchical_graph.py:766 (np.exp on a Python float)
            judge_term[i] = 1.0 / (1.0 + np.exp(-((j or 0.0) / 10.0)))
        c = d.get("confidence")
        if c is not None:
            conf[i] = float(c)
    return {"n_nodes": n, "node_ids": node_ids, "dense": dense, "row_ptr": row_ptr, "col_idx": col_idx,
            "node_type": node_type, "level": level, "judge": judge, "judge_term": judge_term, "conf": conf}


def csr_from_edges(n_nodes: int, src: np.ndarray, dst: np.ndarray, etype: np.ndarray, node_type: np.ndarray,
                   level: Optional[np.ndarray] = None, judge: Optional[np.ndarray] = None,
                   conf: Optional[np.ndarray] = None) -> Dict[str, Any]:
    """Array-level twin for graphs too large for NetworkX (``synth.make_csr_graph``).

    ``src/dst/etype`` are in link order; a node's adjacency order is the order its links appear, which is
    what NetworkX would produce for the same link list (first occurrence of a duplicated pair keeps its
    slot; duplicates are assumed absent here).  A stable sort by source (resp. target) preserves it.
    """
    n = int(n_nodes)
    row_ptr, col_idx = [], []
    for t in range(3):
        m = etype == t
        s, d = src[m], dst[m]
        for direction in (0, 1):
            key, val = (s, d) if direction == 0 else (d, s)
            order = np.argsort(key, kind="stable")
            counts = np.bincount(key, minlength=n).astype(np.int64)
            rp = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(counts, out=rp[1:])
            row_ptr.append(rp)
            col_idx.append(val[order].astype(np.int32))
    judge = np.full(n, np.nan) if judge is None else np.asarray(judge, dtype=np.float64)
    conf = np.full(n, np.nan) if conf is None else np.asarray(conf, dtype=np.float64)
    jt = np.zeros(n, dtype=np.float64)
    has = ~np.isnan(judge)
    jt[has] = 1.0 / (1.0 + np.exp(-(judge[has] / 10.0)))
    return {"n_nodes": n, "node_ids": None, "dense": None, "row_ptr": row_ptr, "col_idx": col_idx,
            "node_type": np.asarray(node_type, dtype=np.uint8),
            "level": np.zeros(n, np.uint8) if level is None else np.asarray(level, dtype=np.uint8),
            "judge": judge, "judge_term": jt, "conf": conf}
