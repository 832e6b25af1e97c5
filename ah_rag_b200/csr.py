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


_ABSENT = object()


def _node_attr_arrays(node_ids: List[Any], attrs: List[Dict[str, Any]]) -> Dict[str, np.ndarray]:
    n = len(node_ids)
    node_type = np.full(n, 3, dtype=np.uint8)
    level = np.zeros(n, dtype=np.uint8)
    judge = np.full(n, np.nan, dtype=np.float64)
    judge_term = np.zeros(n, dtype=np.float64)
    conf = np.full(n, np.nan, dtype=np.float64)
    for i, d in enumerate(attrs):
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
    return {"node_type": node_type, "level": level, "judge": judge, "judge_term": judge_term, "conf": conf}


def csr_from_node_link(data: Dict[str, Any], edges: str = "links") -> Dict[str, Any]:
    """CSR export straight from the on-disk node-link document (``structure.json``, hierarchical_graph.py:175-218),
    without walking a NetworkX graph: equal, array for array, to
    ``csr_from_networkx(nx.node_link_graph(data, directed=True, edges="links"))`` (tests/test_host.py), in vectorised
    numpy -- the per-edge Python loop of the NetworkX route is what dominates start-up on large graphs.

    Rules of ``node_link_graph`` / ``DiGraph.add_edge`` reproduced here: nodes in document order, a repeated node id
    merges its attributes into the first entry; a link naming an unknown node appends it (source before target); a
    repeated (source, target) pair keeps the adjacency slot of its first link and takes the ``edge_type`` of the last
    link that carries that key.
    """
    node_ids: List[Any] = []
    dense: Dict[Any, int] = {}
    attrs: List[Dict[str, Any]] = []
    for d in data.get("nodes", []):
        nid = d["id"]
        rest = {k: v for k, v in d.items() if k != "id"}
        if nid in dense:
            attrs[dense[nid]].update(rest)
        else:
            dense[nid] = len(node_ids)
            node_ids.append(nid)
            attrs.append(rest)
    links = data.get(edges, [])
    m = len(links)
    tcode = {t: i for i, t in enumerate(EDGE_TYPES)}
    src_l = [l["source"] for l in links]
    dst_l = [l["target"] for l in links]
    # -2: the link has no edge_type key (keeps whatever an earlier duplicate set); -1: key present, not a known type
    et = np.fromiter((-2 if (t := l.get("edge_type", _ABSENT)) is _ABSENT else tcode.get(t, -1) if isinstance(t, str) else -1
                      for l in links), dtype=np.int64, count=m)
    src = np.fromiter((dense.get(u, -1) for u in src_l), dtype=np.int64, count=m)
    dst = np.fromiter((dense.get(v, -1) for v in dst_l), dtype=np.int64, count=m)
    if m and (src.min() < 0 or dst.min() < 0):  # rare: endpoints missing from "nodes" are appended in link order
        for i in np.flatnonzero((src < 0) | (dst < 0)):
            for arr, name in ((src, src_l[i]), (dst, dst_l[i])):
                if name not in dense:
                    dense[name] = len(node_ids)
                    node_ids.append(name)
                    attrs.append({})
                arr[i] = dense[name]
    n = len(node_ids)
    if m:
        key = src * n + dst
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        n_pairs = first.shape[0]
        last_typed = np.full(n_pairs, -1, dtype=np.int64)
        typed = np.flatnonzero(et != -2)
        np.maximum.at(last_typed, inv[typed], typed)
        pair_type = np.where(last_typed >= 0, et[np.maximum(last_typed, 0)], -1)
        order = np.argsort(first, kind="stable")          # pairs in the order of their first link
        p_src, p_dst, p_type = src[first][order], dst[first][order], pair_type[order]
    else:
        p_src = p_dst = p_type = np.zeros(0, dtype=np.int64)
    # in-adjacency order of DiGraph: _pred[v] gets u when the pair (u, v) is first added -> also first-link order
    row_ptr, col_idx = [], []
    for t in range(3):
        sel = p_type == t
        s_, d_ = p_src[sel], p_dst[sel]
        for direction in (0, 1):
            k_, v_ = (s_, d_) if direction == 0 else (d_, s_)
            o = np.argsort(k_, kind="stable")
            rp = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.bincount(k_, minlength=n), out=rp[1:])
            row_ptr.append(rp)
            col_idx.append(v_[o].astype(np.int32))
    out = {"n_nodes": n, "node_ids": node_ids, "dense": dense, "row_ptr": row_ptr, "col_idx": col_idx}
    out.update(_node_attr_arrays(node_ids, attrs))
    return out


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
