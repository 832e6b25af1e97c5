"""numpy restatement of the staged expansion over a node-range-sharded CSR (TEST INFRASTRUCTURE).

Same record layout and the same three steps as csrc/graph_shard.cu (stage 1 on the owner of an input node, stage 2 on the
owner of a hyperedge, merge), written with plain Python loops over the arrays of ``ah_rag_b200.csr``.  Used by
  * tests/test_dist_gloo.py  -- world_size 2 on CPU: the stages injected into ``dist.ShardedGraph``, result compared with
                                ``oracle/array_graph.expand`` on the whole graph (environment.py:393-480);
  * tests/test_gpu_shard.py  -- the CUDA stages' records against these, record for record.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

PART_OUT, PART_IN, BEL_OUT, BEL_IN, REL_OUT, REL_IN = range(6)
PRESENT, ENTITY = 1, 2


def geom(limit: int, max_ids: int) -> Tuple[int, int, int]:
    cap = 2 * limit - 1
    return cap, 4 + cap + limit, 4 + cap  # cap, ints of a stage-1 record, ints of a stage-2 record


def _nbrs(g: Dict, slot: int, v: int) -> List[int]:
    rp = g["row_ptr"][slot]
    return [int(x) for x in g["col_idx"][slot][rp[v]:rp[v + 1]]]


def _prefix(values: Sequence[int], cap: int, limit: int) -> Tuple[List[int], int]:
    """Order-preserving unique prefix (at most cap values) and the un-deduplicated length saturated at `limit`."""
    uniq = list(dict.fromkeys(values))[:cap]
    return uniq, min(len(values), limit)


def stage1(g: Dict, lo: int, hi: int, ops, in_ids, in_offsets, max_ids: int, limit: int) -> np.ndarray:
    cap, r1, _ = geom(limit, max_ids)
    n_env = len(ops)
    out = np.zeros((n_env, max_ids * r1), dtype=np.int32)
    for e in range(n_env):
        ids = in_ids[in_offsets[e]:in_offsets[e + 1]]
        for idx in range(min(len(ids), max_ids)):
            v = int(ids[idx])
            if not (lo <= v < hi):
                continue
            op = int(ops[e])
            if op == 0:
                vals = _nbrs(g, BEL_OUT, v)
            elif op == 1:
                vals = _nbrs(g, BEL_IN, v)
            else:
                vals = _nbrs(g, REL_OUT, v) + _nbrs(g, REL_IN, v)
            uniq, raw = _prefix(vals, cap, limit)
            entity = int(g["node_type"][v]) == 0
            hs = _nbrs(g, PART_OUT, v)[:limit] if (op == 2 and entity) else []
            rec = out[e, idx * r1:(idx + 1) * r1]
            rec[0], rec[1], rec[2], rec[3] = raw, PRESENT | (ENTITY if entity else 0), len(uniq), len(hs)
            rec[4:4 + len(uniq)] = uniq
            rec[4 + cap:4 + cap + len(hs)] = hs
    return out


def _owner(bounds, v: int) -> int:
    r = 0
    while r + 1 < len(bounds) - 1 and v >= bounds[r + 1]:
        r += 1
    return r


def stage2(g: Dict, lo: int, hi: int, bounds, ops, in_ids, in_offsets, max_ids: int, limit: int, g1: np.ndarray) -> np.ndarray:
    cap, r1, r2 = geom(limit, max_ids)
    n_env, n = len(ops), len(g["node_type"])
    out = np.zeros((n_env, max_ids * limit * r2), dtype=np.int32)
    for e in range(n_env):
        if int(ops[e]) != 2:
            continue
        ids = in_ids[in_offsets[e]:in_offsets[e + 1]]
        for idx in range(min(len(ids), max_ids)):
            v = int(ids[idx])
            if v < 0 or v >= n:
                continue
            rec1 = g1[_owner(bounds, v), e, idx * r1:(idx + 1) * r1]
            if not (rec1[1] & ENTITY):
                continue
            for j in range(int(rec1[3])):
                h = int(rec1[4 + cap + j])
                if not (lo <= h < hi):
                    continue
                uniq, raw = _prefix([u for u in _nbrs(g, PART_IN, h) if u != v], cap, limit)
                rec = out[e, (idx * limit + j) * r2:(idx * limit + j + 1) * r2]
                rec[0], rec[1], rec[2] = raw, PRESENT, len(uniq)
                rec[4:4 + len(uniq)] = uniq
    return out


def merge(n_nodes: int, bounds, ops, in_ids, in_offsets, max_ids: int, limit: int, g1: np.ndarray, g2: np.ndarray):
    cap, r1, r2 = geom(limit, max_ids)
    n_env = len(ops)
    out_ids = np.full((n_env, limit), -1, dtype=np.int32)
    out_cnt = np.zeros(n_env, dtype=np.int32)
    for e in range(n_env):
        X: List[int] = []
        raw = 0
        ids = in_ids[in_offsets[e]:in_offsets[e + 1]]
        for idx in range(min(len(ids), max_ids)):
            v = int(ids[idx])
            if v < 0 or v >= n_nodes:
                continue
            rec1 = g1[_owner(bounds, v), e, idx * r1:(idx + 1) * r1]
            X += [int(x) for x in rec1[4:4 + rec1[2]]]
            raw += int(rec1[0])
            if int(ops[e]) == 2 and (rec1[1] & ENTITY):
                for j in range(int(rec1[3])):
                    h = int(rec1[4 + cap + j])
                    rec2 = g2[_owner(bounds, h), e, (idx * limit + j) * r2:(idx * limit + j + 1) * r2]
                    X += [h] + [int(x) for x in rec2[4:4 + rec2[2]]]
                    raw += 1 + int(rec2[0])
            if raw >= limit:
                break
        res = list(dict.fromkeys(X))[:limit]
        out_ids[e, :len(res)] = res
        out_cnt[e] = len(res)
    return out_ids, out_cnt


# ---------------------------------------------------------------------------------------------
# anchor expansion + rerank over the sharded graph (csrc/graph_shard.cu: shard_anchor_stage1/2; merge = the rerank of
# oracle/array_graph.anchor_rerank fed from the records).  Record: [present, type, n_nbr, level, judge, judge_term, conf
# (3 float64 as 6 int32)] + nbr[fan] (+ pad to an even count); stage 2: the 10-int head only.
# ---------------------------------------------------------------------------------------------
HEAD = 10


def anchor_geom(n_seeds: int, parent_cap: int, member_top_m: int) -> Tuple[int, int, int]:
    fan = max(parent_cap, member_top_m)
    return fan, HEAD + fan + (fan & 1), HEAD


def _write_head(rec: np.ndarray, g: Dict, v: int, n_nbr: int) -> None:
    rec[0], rec[1], rec[2], rec[3] = PRESENT, int(g["node_type"][v]), n_nbr, int(g["level"][v])
    rec[4:10] = np.array([g["judge"][v], g["judge_term"][v], g["conf"][v]], dtype=np.float64).view(np.int32)


def anchor_stage1(g: Dict, lo: int, hi: int, seeds: np.ndarray, parent_cap: int, member_top_m: int) -> np.ndarray:
    nq, ns = seeds.shape
    fan, r1, _ = anchor_geom(ns, parent_cap, member_top_m)
    out = np.zeros((nq, ns * r1), dtype=np.int32)
    for q in range(nq):
        for r in range(ns):
            v = int(seeds[q, r])
            rec = out[q, r * r1:(r + 1) * r1]
            if not (lo <= v < hi):
                rec[1] = 3
                continue
            t = int(g["node_type"][v])
            nb = _nbrs(g, BEL_OUT, v)[:parent_cap] if t == 0 else (_nbrs(g, BEL_IN, v)[:member_top_m] if t == 1 else [])
            rec[HEAD:HEAD + len(nb)] = nb
            _write_head(rec, g, v, len(nb))
    return out


def anchor_stage2(g: Dict, lo: int, hi: int, bounds, seeds: np.ndarray, parent_cap: int, member_top_m: int, g1: np.ndarray) -> np.ndarray:
    nq, ns = seeds.shape
    fan, r1, r2 = anchor_geom(ns, parent_cap, member_top_m)
    n = len(g["node_type"])
    out = np.zeros((nq, max(1, ns * fan * r2)), dtype=np.int32)
    for q in range(nq):
        for r in range(ns):
            v = int(seeds[q, r])
            if v < 0 or v >= n:
                continue
            rec1 = g1[_owner(bounds, v), q, r * r1:(r + 1) * r1]
            for j in range(int(rec1[2])):
                u = int(rec1[HEAD + j])
                rec = out[q, (r * fan + j) * r2:(r * fan + j + 1) * r2]
                if lo <= u < hi:
                    _write_head(rec, g, u, 0)
                else:
                    rec[1] = 3
    return out


def anchor_merge(n_nodes: int, bounds, seeds: np.ndarray, scores: np.ndarray, top_k: int, parent_cap: int, member_top_m: int,
                 g1: np.ndarray, g2: np.ndarray, params=None):
    """Per query the reranked [(node, round(sem, 4), round(score, 4))] list, exactly as array_graph.anchor_rerank builds it."""
    from . import array_graph
    P = params or array_graph.DEFAULTS
    nq, ns = seeds.shape
    fan, r1, r2 = anchor_geom(ns, parent_cap, member_top_m)
    res = []
    for q in range(nq):
        E: Dict[int, float] = {}
        attr: Dict[int, np.ndarray] = {}
        for r in range(ns):
            v, s = int(seeds[q, r]), scores[q, r]
            if v < 0 or v >= n_nodes:
                continue
            sem = 1.0 - float(np.float32(1.0) - np.float32(s))
            rec1 = g1[_owner(bounds, v), q, r * r1:(r + 1) * r1]
            E[v] = sem
            attr.setdefault(v, rec1[:HEAD])
            t = int(rec1[1])
            for j in range(int(rec1[2])):
                u = int(rec1[HEAD + j])
                if u not in E:
                    E[u] = sem * (0.9 if t == 0 else 0.85)
                    attr[u] = g2[_owner(bounds, u), q, (r * fan + j) * r2:(r * fan + j + 1) * r2]
        out = []
        for node, sem in E.items():
            h = attr[node]
            judge, jt, conf = h[4:10].view(np.float64)
            t = int(h[1])
            jterm = float(jt) if judge == judge else 0.0
            ct = float(conf) / 10.0 if conf == conf else 0.0
            score = P["alpha"] * sem + P["beta"] * jterm + P["gamma"] * ct + P["delta"] * P["boost"][t]
            out.append((node, round(sem, 4), round(float(score), 4)))
        out.sort(key=lambda x: x[2], reverse=True)
        res.append(out[:top_k])
    return res
