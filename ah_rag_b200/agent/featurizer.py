"""Observation -> fixed-length float32 vector (drop-in for ah_rag/agent/featurizer.py:7-56).

Layout: 4 globals (step, |selection_ids|, |frontier_ids|, n_seeds) + k_nodes blocks of 8
(one-hot entity/summary/other, layer, score, semantic, judge_overall, confidence), zero padded.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

NODE_FEATS = 8
GLOBAL_FEATS = 4


def _node_feats(n: Dict[str, Any]) -> List[float]:
    nt = str(n.get("node_type") or "")
    return [1.0 if nt == "entity" else 0.0, 1.0 if nt == "summary" else 0.0,
            1.0 if nt not in ("entity", "summary") else 0.0, float(n.get("layer") or 0),
            float(n.get("score") or 0.0), float(n.get("semantic") or 0.0),
            float(n.get("judge_overall") or 0.0), float(n.get("confidence") or 0.0)]


def featurize_observation(obs: Dict[str, Any], k_nodes: int = 10) -> Tuple[np.ndarray, Dict[str, Any]]:
    state = obs.get("state") or {}
    feats: List[float] = [float(obs.get("step") or 0), float(len(state.get("selection_ids") or [])),
                          float(len(state.get("frontier_ids") or [])), float(len(obs.get("seeds") or []))]
    shown = (obs.get("selection") or [])[:k_nodes]
    ids: List[str] = []
    for n in shown:
        feats.extend(_node_feats(n))
        ids.append(str(n.get("node_id")))
    feats.extend([0.0] * (NODE_FEATS * (k_nodes - len(shown))))
    return np.asarray(feats, dtype=np.float32), {"top_node_ids": ids}
