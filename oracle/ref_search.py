"""CPU restatement of the graph half of the AH-RAG hot path (TEST INFRASTRUCTURE, not product code).

Each function cites the reference lines (under /root/reference) it follows.  It is checked against the
UNMODIFIED reference (run through ``oracle/stubs.py``) by ``tests/test_oracle_vs_reference.py`` in the
dev container and against the committed golden vectors in ``tests/golden/`` everywhere else.

The graph is held in NetworkX exactly as the reference's loader produces it
(``nx.node_link_graph(data, directed=True, edges="links")``, hierarchical_graph.py:196), because the
adjacency iteration order of the LOADED graph is part of the semantics being pinned.
"""
from __future__ import annotations

import json
import math
import os
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import networkx as nx
import numpy as np

from . import exact


# ---------------------------------------------------------------------------------------------
# graph container
# ---------------------------------------------------------------------------------------------
class OracleGraph:
    """Loaded graph + search defaults (hierarchical_graph.py:191-218)."""

    def __init__(self, node_link: Dict[str, Any], meta: Optional[Dict[str, Any]] = None) -> None:
        self.G: nx.DiGraph = nx.node_link_graph(node_link, directed=True, edges="links")
        sp = {"alpha": 0.6, "beta": 0.2, "gamma": 0.1, "delta": 0.1, "judge_overall_min": None,
              "confidence_min": None, "member_top_m": 5, "type_filter": None,
              "layer_boost": {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0}}  # :37-47
        # load(): hg = cls() (defaults) ; hg.G = node_link_graph(...) REPLACES the graph incl. G.graph,
        # then meta.json search_params are merged over G.graph["search_params"] (:209-211)
        base = self.G.graph.get("search_params", {})
        if meta and isinstance(meta.get("search_params"), dict):
            base = {**base, **meta["search_params"]}
        self.G.graph["search_params"] = base
        self._fallback_sp = sp

    @classmethod
    def load(cls, directory: str) -> "OracleGraph":
        with open(os.path.join(directory, "structure.json"), "r", encoding="utf-8") as f:
            data = json.load(f)
        meta = None
        mp = os.path.join(directory, "meta.json")
        if os.path.exists(mp):
            with open(mp, "r", encoding="utf-8") as f:
                meta = json.load(f)
        return cls(data, meta)

    # hierarchical_graph.py:441-453
    def neighbors(self, v: str, direction: str, edge_type: str) -> List[str]:
        if direction == "out":
            return [w for _, w, d in self.G.out_edges(v, data=True) if d.get("edge_type") == edge_type]
        return [u for u, _, d in self.G.in_edges(v, data=True) if d.get("edge_type") == edge_type]


def judge_overall(d: Dict[str, Any]) -> Optional[float]:
    """hierarchical_graph.py:746-752 / environment.py:99-105."""
    js = d.get("judge_scores")
    if not js:
        return None
    try:
        if isinstance(js, str):
            js = json.loads(js)
        return float(js.get("overall", 0.0))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# A.1  anchor expansion + rerank  (hierarchical_graph.py:681-699, 719-789)
# ---------------------------------------------------------------------------------------------
def resolve_params(og: OracleGraph, member_top_m=5, alpha=None, beta=None, gamma=None, delta=None,
                   judge_overall_min=None, confidence_min=None, type_filter=None) -> Dict[str, Any]:
    sp = og.G.graph.get("search_params", {})
    p = {
        "alpha": alpha if alpha is not None else sp.get("alpha", 0.6),
        "beta": beta if beta is not None else sp.get("beta", 0.2),
        "gamma": gamma if gamma is not None else sp.get("gamma", 0.1),
        "delta": delta if delta is not None else sp.get("delta", 0.1),
        "judge_overall_min": judge_overall_min if judge_overall_min is not None else sp.get("judge_overall_min"),
        "confidence_min": confidence_min if confidence_min is not None else sp.get("confidence_min"),
        "member_top_m": member_top_m if member_top_m is not None else sp.get("member_top_m", 5),
    }
    if type_filter is None:
        tf = sp.get("type_filter")
        type_filter = list(tf) if isinstance(tf, (set, tuple)) else tf
    p["type_filter"] = type_filter
    p["layer_boost"] = sp.get("layer_boost", {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0})
    return p


def seeds_from_topk(ids: Sequence[Optional[str]], scores: Sequence[float]) -> List[Dict[str, Any]]:
    """hierarchical_graph.py:719-724 with dist = fl32(1 - s) as hnswlib's cosine space reports it."""
    seeds = []
    for nid, s in zip(ids, scores):
        if nid is None:
            continue
        dist = np.float32(1.0) - np.float32(s)
        seeds.append({"node_id": nid, "semantic": 1.0 - float(dist)})
    return seeds


def anchor_rerank(og: OracleGraph, seeds: List[Dict[str, Any]], top_k: int, params: Dict[str, Any]) -> Dict[str, Any]:
    G = og.G
    m = params["member_top_m"]
    expanded: Dict[str, Dict[str, Any]] = {}
    for s in seeds:  # rank order
        nid = s["node_id"]
        nt = G.nodes.get(nid, {}).get("node_type")
        expanded[nid] = {"node_id": nid, "semantic": s["semantic"], "node_type": nt}  # :732 overwrite, keeps slot
        if nt == "entity":
            for parent in og.neighbors(nid, "out", "belongs_to")[:2]:  # :733-736
                if parent not in expanded:
                    expanded[parent] = {"node_id": parent, "semantic": s["semantic"] * 0.9, "node_type": "summary"}
        elif nt == "summary":
            for child in og.neighbors(nid, "in", "belongs_to")[:m]:  # :737-740
                if child not in expanded:
                    expanded[child] = {"node_id": child, "semantic": s["semantic"] * 0.85, "node_type": "entity"}

    alpha, beta, gamma, delta = params["alpha"], params["beta"], params["gamma"], params["delta"]
    tf, jmin, cmin, boost = params["type_filter"], params["judge_overall_min"], params["confidence_min"], params["layer_boost"]
    results = []
    for nid, info in expanded.items():  # :744-779
        d = G.nodes.get(nid, {})
        judge = judge_overall(d)
        conf = d.get("confidence")
        ntype = d.get("node_type")
        layer = 0 if ntype == "entity" else (1 if ntype == "summary" else 0)
        if tf and ntype not in tf:
            continue
        if jmin is not None and (judge is None or judge < jmin):
            continue
        if cmin is not None and (conf is None or conf < cmin):
            continue
        layer_boost = float(boost.get(ntype, 0.0))
        semantic = float(info.get("semantic", 0.0))
        judge_term = (1.0 / (1.0 + np.exp(-((judge or 0.0) / 10.0)))) if judge is not None else 0.0
        conf_term = float(conf) / 10.0 if conf is not None else 0.0
        score = alpha * semantic + beta * judge_term + gamma * conf_term + delta * layer_boost
        results.append({"node_id": nid, "node_type": ntype, "layer": layer, "semantic": round(semantic, 4),
                        "judge_overall": judge, "confidence": conf, "score": round(float(score), 4),
                        "name": d.get("name"), "title": d.get("title")})
    results.sort(key=lambda x: x["score"], reverse=True)  # stable, key = ROUNDED score (:781)
    return {"seeds": seeds, "expanded": list(expanded.values()), "reranked": results[:top_k]}


# ---------------------------------------------------------------------------------------------
# A.2  frontier expansion ops  (environment.py:393-480)
# ---------------------------------------------------------------------------------------------
def expand(og: OracleGraph, op: str, node_ids: Iterable[str], limit: int = 10) -> List[str]:
    G = og.G
    X: List[str] = []
    for v in node_ids:
        if v not in G:
            continue
        if op == "parents":  # :424-426
            X += og.neighbors(v, "out", "belongs_to")
        elif op == "children":  # :401-403
            X += og.neighbors(v, "in", "belongs_to")
        elif op == "related":  # :448-467
            X += og.neighbors(v, "out", "related_to")
            X += og.neighbors(v, "in", "related_to")
            if G.nodes.get(v, {}).get("node_type") == "entity":
                for h in og.neighbors(v, "out", "participates_in"):
                    X.append(h)
                    X += [u for u in og.neighbors(h, "in", "participates_in") if u != v]
        else:
            raise ValueError(op)
        if len(X) >= limit:  # checked after finishing v, on the un-deduplicated length
            break
    return list(dict.fromkeys(X))[:limit]


# ---------------------------------------------------------------------------------------------
# observation / environment state  (environment.py:92-150, 155-241, 340-355)
# ---------------------------------------------------------------------------------------------
def node_brief(og: OracleGraph, nid: str) -> Dict[str, Any]:
    d = og.G.nodes.get(nid, {})
    nt = d.get("node_type")
    layer = 0 if nt == "entity" else (d.get("level") or (1 if nt == "summary" else 0))
    return {"node_id": nid, "node_type": nt, "entity_type": d.get("entity_type"), "layer": layer,
            "title": d.get("title"), "name": d.get("name"), "judge_overall": judge_overall(d),
            "confidence": d.get("confidence")}


class OracleEnv:
    """State machine of GraphEnvironment restricted to the hot path (no logging, no LCA/details)."""

    def __init__(self, og: OracleGraph, index_ids: Sequence[str], index_rows: np.ndarray) -> None:
        self.og = og
        self.index_ids = list(index_ids)
        self.index_rows = index_rows  # fp32 [n, d] or bf16 bits
        self.step_count = 0
        self.selection_set: set = set()
        self.frontier_set: set = set()
        self.filters = {"judge_overall_min": None, "confidence_min": None, "type_filter": None}
        self.weights = {"alpha": None, "beta": None, "gamma": None, "delta": None}

    def _obs(self, seeds, reranked) -> Dict[str, Any]:
        def brief(res):
            b = node_brief(self.og, res.get("node_id"))
            b.update({"score": res.get("score"), "semantic": res.get("semantic")})
            return b
        return {"selection": [brief(x) for x in reranked], "seeds": [brief(x) for x in seeds],
                "state": {"selection_ids": sorted(self.selection_set),
                          "frontier_ids": sorted(self.frontier_set)[:50]},
                "counts": {"n_nodes": self.og.G.number_of_nodes(), "n_edges": self.og.G.number_of_edges()},
                "step": self.step_count}

    def search(self, q_vec: np.ndarray, top_k: int, member_top_m: int = 5, **kw) -> Dict[str, Any]:
        n = len(self.index_ids)
        if n == 0:
            seeds = []
        else:
            k = min(top_k, n)
            sc, rows = exact.exact_topk_np(self.index_rows, np.asarray(q_vec)[None, :], k)
            seeds = seeds_from_topk([self.index_ids[r] for r in rows[0]], sc[0])
        params = resolve_params(self.og, member_top_m=member_top_m, **kw)
        return anchor_rerank(self.og, seeds, top_k, params)

    def reset(self, q_vec: Optional[np.ndarray], top_k: int = 5):
        self.step_count = 0
        self.selection_set.clear()
        self.frontier_set.clear()
        return self.semantic_anchor(q_vec, top_k=top_k)

    def semantic_anchor(self, q_vec, top_k: int = 5, member_top_m: int = 5):
        self.step_count += 1
        cluster = self.search(q_vec, top_k, member_top_m, judge_overall_min=self.filters["judge_overall_min"],
                              confidence_min=self.filters["confidence_min"], type_filter=self.filters["type_filter"],
                              **self.weights)
        reranked = cluster["reranked"]
        self.frontier_set = set(x["node_id"] for x in reranked if x.get("node_id"))  # :225 replace
        obs = self._obs(cluster["seeds"], reranked)
        return obs, {"action": "semantic_anchor", "top_k": top_k, "returned": len(reranked)}

    def _expand(self, op: str, node_ids, limit):
        self.step_count += 1
        out = expand(self.og, op, node_ids, limit)
        seeds = [{"node_id": n, "semantic": 0.0} for n in out]
        obs = self._obs(seeds, [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in out])  # before update (:409-411)
        self.frontier_set.update(out)
        return obs, {"action": f"expand_{op}", "inputs": list(node_ids), "returned": len(out)}

    def expand_parents(self, node_ids, limit: int = 10):
        return self._expand("parents", node_ids, limit)

    def expand_children(self, node_ids, limit: int = 10):
        return self._expand("children", node_ids, limit)

    def expand_related(self, node_ids, limit: int = 10):
        return self._expand("related", node_ids, limit)

    def commit_selection(self, node_ids):
        self.step_count += 1
        added = []
        for nid in node_ids:
            if nid in self.og.G and nid not in self.selection_set:
                self.selection_set.add(nid)
                added.append(nid)
                self.frontier_set.discard(nid)
        obs = self._obs([], [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in added])
        return obs, {"action": "commit_selection", "added": added, "total_selection": len(self.selection_set)}


# ---------------------------------------------------------------------------------------------
# featurizer / reward  (featurizer.py:7-56, reward.py:6-26) and gym step (gym_env.py:85-125)
# ---------------------------------------------------------------------------------------------
def featurize(obs: Dict[str, Any], k_nodes: int = 10) -> np.ndarray:
    state = obs.get("state") or {}
    feats = [float(obs.get("step") or 0), float(len(state.get("selection_ids") or [])),
             float(len(state.get("frontier_ids") or [])), float(len(obs.get("seeds") or []))]
    sel = (obs.get("selection") or [])[:k_nodes]
    for n in sel:
        nt = str(n.get("node_type") or "")
        feats += [1.0 if nt == "entity" else 0.0, 1.0 if nt == "summary" else 0.0,
                  1.0 if nt not in {"entity", "summary"} else 0.0, float(n.get("layer") or 0),
                  float(n.get("score") or 0.0), float(n.get("semantic") or 0.0),
                  float(n.get("judge_overall") or 0.0), float(n.get("confidence") or 0.0)]
    feats += [0.0] * (8 * (k_nodes - len(sel)))
    return np.asarray(feats, dtype=np.float32)


def step_reward(prev: Optional[Dict[str, Any]], cur: Dict[str, Any]) -> float:
    if prev is None:
        return 0.0
    ps, cs = prev.get("state") or {}, cur.get("state") or {}
    add_sel = len(set(cs.get("selection_ids") or []) - set(ps.get("selection_ids") or []))
    add_fr = max(0, len(set(cs.get("frontier_ids") or [])) - len(set(ps.get("frontier_ids") or [])))
    return float(1.0 * add_sel + 0.05 * min(add_fr, 10) - 0.05)


class OracleGymEnv:
    """gym_env.py:27-125 over OracleEnv; action 4 (query_details_top) keeps only what the featurizer reads."""

    def __init__(self, env: OracleEnv, max_steps: int = 6, repeat_penalty: float = 0.02) -> None:
        self.env = env
        self.max_steps = max(1, int(max_steps))
        self.repeat_penalty = float(repeat_penalty)
        self._cur = 0
        self._raw: Optional[Dict[str, Any]] = None
        self._last_action: Optional[int] = None

    def reset(self, q_vec):
        self._cur = 0
        raw, _ = self.env.reset(q_vec)  # top_k=5 default (gym_env.py:44 -> environment.py:155)
        self._raw = raw
        self._last_action = None
        return featurize(raw), raw

    def _top(self, k: int) -> List[str]:
        return [n["node_id"] for n in (self._raw.get("selection") or [])[:k] if n.get("node_id")]

    def step(self, action: int):
        self._cur += 1
        prev, done = self._raw, False
        if action == 0:
            obs, _ = self.env.expand_parents(self._top(2))
        elif action == 1:
            obs, _ = self.env.expand_children(self._top(2))
        elif action == 2:
            obs, _ = self.env.expand_related(self._top(1))
        elif action == 3:
            obs, _ = self.env.commit_selection(self._top(3))
        elif action == 4:
            ids = self._top(1)
            if ids:  # query_node_details (environment.py:301-335): one detailed brief, no seeds, no state
                self.env.step_count += 1
                obs = {"selection": [node_brief(self.env.og, ids[0])], "seeds": [],
                       "counts": prev.get("counts"), "step": self.env.step_count}
            else:
                obs = prev
        else:
            done, obs = True, prev
        r = 0.0 if done else step_reward(prev, obs)
        if not done and self._last_action is not None and int(action) == self._last_action and self.repeat_penalty > 0:
            r -= self.repeat_penalty
        done = done or self._cur >= self.max_steps
        self._raw = obs
        self._last_action = int(action)
        return featurize(obs), float(r), bool(done), obs
