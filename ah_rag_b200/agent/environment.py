"""Drop-in for ``ah_rag.agent.environment.GraphEnvironment`` (environment.py:15-511) on the hot path.

Same verbs, same ``(observation, info)`` schema, same state rules; underneath, ``semantic_anchor`` is one
fused top-k + one expansion/rerank launch (``HierarchicalGraph.search_batch``) and ``expand_*`` is one
``ahr_expand`` launch over the device CSR instead of NetworkX walks.  ``VecGraphEnvironment`` is the additive
batched form the RL rollout loop (policy_ppo.py:163-205) should drive: B environments per launch.

``expand_to_lca`` (environment.py:243-299) runs its closure on the device CSR (``ahr_lca_closure``,
one sync-free launch) and reproduces the reference's quirk (descendant sets).  Out of scope: the
structlog session logger -- a plain JSONL writer keeps the file layout (session.json / events.jsonl / summary.json) when
logging is enabled.
"""
from __future__ import annotations

import json
import os
import time
import uuid
from datetime import datetime, timezone
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..csr import judge_overall
from ..graph.hierarchical_graph import HierarchicalGraph

_OPS = {"parents": 0, "children": 1, "related": 2}


def _utcnow() -> datetime:
    return datetime.now(timezone.utc).replace(tzinfo=None)


class GraphEnvironment:
    def __init__(self, graph_dir: str = "graph", random_state: int = 42, logging_enabled: bool = True,
                 log_dir: str = "artifacts/phase2", session_id: Optional[str] = None, debug: bool = False,
                 log_level: str = "normal", redact: bool = True, hg: Optional[HierarchicalGraph] = None) -> None:
        self.graph_dir = graph_dir
        self.random_state = random_state
        self.hg: Optional[HierarchicalGraph] = hg
        self.last_query: Optional[str] = None
        self.last_results: Optional[Dict[str, Any]] = None
        self.step_count = 0
        self.selection_set: set = set()
        self.frontier_set: set = set()
        self.current_filters: Dict[str, Any] = {"judge_overall_min": None, "confidence_min": None, "type_filter": None}
        self.current_weights: Dict[str, Any] = {"alpha": None, "beta": None, "gamma": None, "delta": None,
                                                "member_top_m": None, "top_k": 5}
        self.logging_enabled = logging_enabled
        self.log_dir = log_dir
        self.log_level = log_level
        self.session_id = session_id or _utcnow().strftime("%Y%m%dT%H%M%S") + "-" + uuid.uuid4().hex[:6]
        self.session_path = os.path.join(self.log_dir, self.session_id)
        self.debug = debug
        self.stats: Dict[str, Any] = {"actions": [], "cumulative": {"steps": 0, "queries": 0, "expansions": 0, "time_s": 0.0}}
        if self.logging_enabled:
            try:
                os.makedirs(self.session_path, exist_ok=True)
                with open(os.path.join(self.session_path, "session.json"), "w", encoding="utf-8") as f:
                    json.dump({"session_id": self.session_id, "created_at": _utcnow().isoformat() + "Z"}, f)
            except Exception:
                pass
        self._ensure_graph_loaded()

    # ------------------------------------------------------------------ plumbing
    def _ensure_graph_loaded(self) -> None:
        if self.hg is None:
            self.hg = HierarchicalGraph.load(self.graph_dir)

    def _log(self, event: Dict[str, Any]) -> None:
        if not self.logging_enabled or self.log_level == "off":
            return
        try:
            with open(os.path.join(self.session_path, "events.jsonl"), "a", encoding="utf-8") as f:
                f.write(json.dumps({**event, "step": self.step_count, "ts": _utcnow().isoformat() + "Z"}, default=str) + "\n")
        except Exception:
            pass

    def _counts(self) -> Dict[str, int]:
        return {"n_nodes": self.hg.G.number_of_nodes(), "n_edges": self.hg.G.number_of_edges()}

    def _node_brief(self, node_id: str) -> Dict[str, Any]:
        d = self.hg.G.nodes.get(node_id, {})
        nt = d.get("node_type")
        layer = 0 if nt == "entity" else (d.get("level") or (1 if nt == "summary" else 0))
        return {"node_id": node_id, "node_type": nt, "entity_type": d.get("entity_type"), "layer": layer,
                "title": d.get("title"), "name": d.get("name"), "judge_overall": judge_overall(d),
                "confidence": d.get("confidence")}

    def _observation(self, seeds: List[Dict[str, Any]], reranked: List[Dict[str, Any]]) -> Dict[str, Any]:
        def brief(res: Dict[str, Any]) -> Dict[str, Any]:
            b = self._node_brief(res.get("node_id"))
            b.update({"score": res.get("score"), "semantic": res.get("semantic")})
            return b

        obs: Dict[str, Any] = {
            "selection": [brief(x) for x in reranked],
            "seeds": [brief(x) for x in seeds],
            "state": {"selection_ids": sorted(self.selection_set), "frontier_ids": sorted(self.frontier_set)[:50]},
            "counts": self._counts(),
            "step": self.step_count,
        }
        if self.debug:
            obs["diagnostics"] = {"filters": self.current_filters, "weights": self.current_weights,
                                  "last_query": self.last_query, "frontier_size": len(self.frontier_set),
                                  "selection_size": len(self.selection_set)}
        return obs

    def _record(self, info: Dict[str, Any], steps: int = 1, **cumulative: float) -> None:
        self.stats["actions"].append(info)
        self.stats["cumulative"]["steps"] += steps
        for key, val in cumulative.items():
            self.stats["cumulative"][key] += val

    # ------------------------------------------------------------------ core API (environment.py:155-241)
    def reset(self, seed_query: Optional[str] = None, top_k: int = 5) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._ensure_graph_loaded()
        self.last_query = None
        self.last_results = None
        self.step_count = 0
        self.selection_set.clear()
        self.frontier_set.clear()
        if seed_query:
            return self.semantic_anchor(seed_query, top_k=top_k)
        obs = {"selection": [], "seeds": [], "counts": self._counts(), "step": self.step_count}
        self._log({"action": "reset", "message": "reset without seed_query"})
        return obs, {"message": "reset without seed_query"}

    def _anchor_kwargs(self, top_k, member_top_m, judge_overall_min, confidence_min, type_filter) -> Dict[str, Any]:
        f, w = self.current_filters, self.current_weights
        return {"top_k": top_k, "member_top_m": member_top_m,
                "judge_overall_min": f.get("judge_overall_min") if judge_overall_min is None else judge_overall_min,
                "confidence_min": f.get("confidence_min") if confidence_min is None else confidence_min,
                "type_filter": f.get("type_filter") if type_filter is None else type_filter,
                "alpha": w.get("alpha"), "beta": w.get("beta"), "gamma": w.get("gamma"), "delta": w.get("delta")}

    def _finish_anchor(self, query: Any, cluster: Dict[str, Any], top_k: int, dur: float):
        seeds, reranked = cluster.get("seeds", []), cluster.get("reranked", [])
        self.frontier_set = set(x.get("node_id") for x in reranked if x.get("node_id"))  # replaces (:225)
        obs = self._observation(seeds, reranked)
        info = {"action": "semantic_anchor", "query": query, "top_k": top_k, "returned": len(reranked),
                "time_s": round(dur, 4)}
        self.last_results = cluster
        self._log({**info, "filters": self.current_filters, "weights": self.current_weights})
        self._record(info, queries=1, time_s=dur)
        return obs, info

    def semantic_anchor(self, query: str, top_k: int = 5, member_top_m: int = 5,
                        judge_overall_min: Optional[float] = None, confidence_min: Optional[float] = None,
                        type_filter: Optional[List[str]] = None) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._ensure_graph_loaded()
        self.step_count += 1
        self.last_query = query
        t0 = time.perf_counter()
        kw = self._anchor_kwargs(top_k, member_top_m, judge_overall_min, confidence_min, type_filter)
        cluster = self.hg.search(query=query, return_cluster=True, **kw)
        return self._finish_anchor(query, cluster, top_k, time.perf_counter() - t0)

    @staticmethod
    def _belongs_to_is_dag(arrays: Dict[str, Any]) -> bool:
        """``nx.is_directed_acyclic_graph`` of the belongs_to subgraph: a property of the loaded graph, evaluated once
        per export on the host arrays (Kahn's algorithm, level by level) and cached with them."""
        if "bel_is_dag" not in arrays:
            rp_out, ci_out = arrays["row_ptr"][2], arrays["col_idx"][2]
            indeg = np.diff(arrays["row_ptr"][3]).astype(np.int64)
            alive = np.ones(arrays["n_nodes"], dtype=bool)
            frontier = np.flatnonzero(indeg == 0)
            while frontier.size:
                alive[frontier] = False
                starts, ends = rp_out[frontier], rp_out[frontier + 1]
                if int((ends - starts).sum()) == 0:
                    break
                tgt = np.concatenate([ci_out[a:b] for a, b in zip(starts, ends) if b > a])
                np.subtract.at(indeg, tgt, 1)
                cand = np.unique(tgt)
                frontier = cand[(indeg[cand] == 0) & alive[cand]]
            arrays["bel_is_dag"] = not bool(alive.any())
        return arrays["bel_is_dag"]

    def expand_to_lca(self, node_ids: List[str], max_results: int = 5):
        """environment.py:243-299 over the device-resident CSR instead of a DiGraph rebuilt per
        call (``ahr_lca_closure``, csrc/lca.cu).  As in the reference, "ancestors" are taken with ``nx.ancestors`` on
        child -> parent edges, i.e. the nodes BELOW each input (its transitive members) plus the input itself; the
        result is the set of nodes common to all inputs that have no belongs_to parent inside that common set, ordered
        by ``(level or 1, node id)``.  For distinct entities that set is empty -- the reference's behaviour, kept.  An
        input that has no belongs_to edge at all raises ``networkx.NetworkXError`` like ``nx.ancestors`` does."""
        import networkx as nx
        self._ensure_graph_loaded()
        self.step_count += 1
        dg = self.hg._device_graph()
        arrays = dg["arrays"]
        dense, names = arrays["dense"], arrays["node_ids"]
        rp_in, rp_out = arrays["row_ptr"][3], arrays["row_ptr"][2]
        ids: List[int] = []
        for nid in node_ids:  # membership in the belongs_to subgraph, checked in input order like the reference's loop
            v = dense.get(nid, -1)
            if v < 0 or (rp_in[v + 1] == rp_in[v] and rp_out[v + 1] == rp_out[v]):
                raise nx.NetworkXError(f"The node {nid} is not in the graph.")
            ids.append(v)
        lcas = [names[v] for v in dg["csr"].lca_closure(ids).cpu().tolist()]
        t0 = time.perf_counter()
        G = self.hg.G
        lcas = sorted(lcas, key=lambda x: (G.nodes[x].get("level") or 1, x))[:max_results]
        obs = self._observation([{"node_id": n, "semantic": 0.0} for n in lcas],
                                [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in lcas])
        info = {"action": "expand_to_lca", "inputs": node_ids, "lca_count": len(lcas),
                "dag": self._belongs_to_is_dag(arrays), "time_s": round(time.perf_counter() - t0, 4)}
        self._log(info)
        self._record(info, expansions=1)
        return obs, info

    def query_node_details(self, node_id: str) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._ensure_graph_loaded()
        self.step_count += 1
        if node_id not in self.hg.G:
            return ({"selection": [], "seeds": [], "counts": self._counts(), "step": self.step_count},
                    {"error": "node_not_found", "node_id": node_id})
        d = self.hg.G.nodes[node_id]
        details = {**self._node_brief(node_id), "title": d.get("title"), "name": d.get("name"),
                   "summary_text": (d.get("summary_text") or d.get("summary") or "")[:500],
                   "description": (d.get("description") or "")[:500], "top_words": d.get("top_words"),
                   "members": d.get("members")}
        obs = {"selection": [details], "seeds": [], "counts": self._counts(), "step": self.step_count}
        info = {"action": "query_node_details", "node_id": node_id}
        self._log(info)
        self._record(info)
        return obs, info

    # ------------------------------------------------------------------ state + dynamic params (:340-388)
    def commit_selection(self, node_ids: List[str]) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._ensure_graph_loaded()
        self.step_count += 1
        added = []
        for nid in node_ids:
            if nid in self.hg.G and nid not in self.selection_set:
                self.selection_set.add(nid)
                added.append(nid)
                self.frontier_set.discard(nid)
        obs = self._observation([], [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in added])
        info = {"action": "commit_selection", "added": added, "total_selection": len(self.selection_set)}
        self._log(info)
        self._record(info)
        return obs, info

    def set_filters(self, judge_overall_min: Optional[float] = None, confidence_min: Optional[float] = None,
                    type_filter: Optional[List[str]] = None) -> Dict[str, Any]:
        if judge_overall_min is not None:
            self.current_filters["judge_overall_min"] = judge_overall_min
        if confidence_min is not None:
            self.current_filters["confidence_min"] = confidence_min
        if type_filter is not None:
            self.current_filters["type_filter"] = list(type_filter)
        info = {"action": "set_filters", **self.current_filters}
        self._log(info)
        self.stats["actions"].append(info)
        return info

    def set_search_weights(self, alpha: Optional[float] = None, beta: Optional[float] = None,
                           gamma: Optional[float] = None, delta: Optional[float] = None,
                           member_top_m: Optional[int] = None, top_k: Optional[int] = None) -> Dict[str, Any]:
        # member_top_m / top_k are stored but never consumed by semantic_anchor, as in the reference (:200-210)
        for key, val in (("alpha", alpha), ("beta", beta), ("gamma", gamma), ("delta", delta),
                         ("member_top_m", member_top_m), ("top_k", top_k)):
            if val is not None:
                self.current_weights[key] = val
        info = {"action": "set_search_weights", **self.current_weights}
        self._log(info)
        self.stats["actions"].append(info)
        return info

    # ------------------------------------------------------------------ graph expansions (:393-480) on the GPU
    def _expand_device(self, op: str, node_ids: Sequence[str], limit: int) -> List[str]:
        import torch
        dg = self.hg._device_graph()
        dense, names, csr = dg["arrays"]["dense"], dg["arrays"]["node_ids"], dg["csr"]
        ids = torch.tensor([dense.get(n, -1) for n in node_ids], dtype=torch.int32, device=csr.device)
        ops = torch.tensor([_OPS[op]], dtype=torch.int32, device=csr.device)
        offs = torch.tensor([0, len(node_ids)], dtype=torch.int32, device=csr.device)
        out_ids, out_cnt = csr.expand(ops, ids, offs, int(limit))
        n = int(out_cnt[0].item())
        return [names[i] for i in out_ids[0, :n].tolist()]

    def _expand(self, op: str, node_ids: List[str], limit: int) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        self._ensure_graph_loaded()
        self.step_count += 1
        expanded = self._expand_device(op, node_ids, limit) if limit >= 1 else []
        seeds = [{"node_id": n, "semantic": 0.0} for n in expanded]
        obs = self._observation(seeds, [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in expanded])
        info = {"action": f"expand_{op}", "inputs": node_ids, "returned": len(expanded)}
        self.frontier_set.update(expanded)  # after the observation is built, as in the reference (:409-411)
        self._log(info)
        self._record(info, expansions=1)
        return obs, info

    def expand_children(self, node_ids: List[str], limit: int = 10):
        return self._expand("children", node_ids, limit)

    def expand_parents(self, node_ids: List[str], limit: int = 10):
        return self._expand("parents", node_ids, limit)

    def expand_related(self, node_ids: List[str], limit: int = 10):
        return self._expand("related", node_ids, limit)

    # ------------------------------------------------------------------ debug + end (:485-509)
    def set_debug(self, enabled: bool = True) -> Dict[str, Any]:
        self.debug = enabled
        info = {"action": "set_debug", "debug": self.debug}
        self._log(info)
        self.stats["actions"].append(info)
        return info

    def end_episode(self) -> Dict[str, Any]:
        summary = {"session_id": self.session_id, "created_at": _utcnow().isoformat() + "Z",
                   "selection_size": len(self.selection_set), "frontier_size": len(self.frontier_set),
                   "stats": self.stats, "filters": self.current_filters, "weights": self.current_weights,
                   "last_query": self.last_query}
        if self.logging_enabled:
            try:
                with open(os.path.join(self.session_path, "summary.json"), "w", encoding="utf-8") as f:
                    json.dump(summary, f, ensure_ascii=False, indent=2, default=str)
            except Exception:
                pass
        self._log({"action": "end_episode"})
        return summary


class VecGraphEnvironment:
    """B environments over ONE shared device graph/index (the reference loads a graph copy per env and rolls
    them sequentially, policy_ppo.py:144-205).  Each verb takes per-env inputs and issues one launch for the
    whole batch; per-env state and observations are exactly those of ``GraphEnvironment``."""

    def __init__(self, hg: HierarchicalGraph, n_envs: int) -> None:
        self.hg = hg
        self.envs = [GraphEnvironment(hg=hg, logging_enabled=False) for _ in range(n_envs)]

    def __len__(self) -> int:
        return len(self.envs)

    def reset_batch(self, query_embeddings, top_k: int = 5, member_top_m: int = 5, queries: Optional[Sequence[Any]] = None):
        """One fused top-k + one expansion/rerank launch for all environments."""
        t0 = time.perf_counter()
        for e in self.envs:
            e.last_query, e.last_results, e.step_count = None, None, 0
            e.selection_set.clear()
            e.frontier_set.clear()
            e.step_count += 1
        e0 = self.envs[0]
        kw = e0._anchor_kwargs(top_k, member_top_m, None, None, None)
        clusters = self.hg.search_batch(query_embeddings, return_cluster=True, **kw)
        dur = (time.perf_counter() - t0) / max(1, len(self.envs))
        out = []
        for b, (e, cl) in enumerate(zip(self.envs, clusters)):
            q = queries[b] if queries is not None else None
            e.last_query = q
            out.append(e._finish_anchor(q, cl, top_k, dur))
        return out

    def expand_batch(self, ops: Sequence[str], node_ids: Sequence[Sequence[str]], limit: int = 10):
        """ops[b] in {"parents","children","related"}; one ahr_expand launch for the batch."""
        import torch
        dg = self.hg._device_graph()
        dense, names, csr = dg["arrays"]["dense"], dg["arrays"]["node_ids"], dg["csr"]
        flat, offs = [], [0]
        for ids in node_ids:
            flat += [dense.get(n, -1) for n in ids]
            offs.append(len(flat))
        dev = csr.device
        out_ids, out_cnt = csr.expand(torch.tensor([_OPS[o] for o in ops], dtype=torch.int32, device=dev),
                                      torch.tensor(flat, dtype=torch.int32, device=dev),
                                      torch.tensor(offs, dtype=torch.int32, device=dev), int(limit))
        out_ids, out_cnt = out_ids.cpu().numpy(), out_cnt.cpu().numpy()
        res = []
        for b, e in enumerate(self.envs[:len(ops)]):
            e.step_count += 1
            expanded = [names[i] for i in out_ids[b, :out_cnt[b]]]
            seeds = [{"node_id": n, "semantic": 0.0} for n in expanded]
            obs = e._observation(seeds, [{"node_id": n, "score": 0.0, "semantic": 0.0} for n in expanded])
            info = {"action": f"expand_{ops[b]}", "inputs": list(node_ids[b]), "returned": len(expanded)}
            e.frontier_set.update(expanded)
            e._record(info, expansions=1)
            res.append((obs, info))
        return res
