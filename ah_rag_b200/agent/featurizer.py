"""Observation -> fixed-length float32 vector (drop-in for ah_rag/agent/featurizer.py:7-56).

Layout of the 4 + 8 * k_nodes values:
    [0] obs["step"]   [1] |state.selection_ids|   [2] |state.frontier_ids|   [3] |obs["seeds"]|
    then, for each of the first k_nodes entries of obs["selection"]:
    is_entity, is_summary, is_other, layer, score, semantic, judge_overall, confidence      (missing / None -> 0)
    zero padded to k_nodes blocks.
Values are gathered in float64 and rounded once to float32, as a Python list of floats turned into a float32 array is.
The device twin (all environments in one launch) is ``env_featurize`` in csrc/env.cu.
"""
from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple

import numpy as np

GLOBAL_FEATS = 4
NODE_FEATS = 8
_NUMERIC_KEYS = ("layer", "score", "semantic", "judge_overall", "confidence")  # columns 3..7 of a node block


def _fill(row: np.ndarray, obs: Dict[str, Any], k_nodes: int) -> list:
    state = obs.get("state") or {}
    row[0] = float(obs.get("step") or 0)
    row[1] = len(state.get("selection_ids") or ())
    row[2] = len(state.get("frontier_ids") or ())
    row[3] = len(obs.get("seeds") or ())
    shown_ids = []
    for slot, node in enumerate((obs.get("selection") or ())[:k_nodes]):
        base = GLOBAL_FEATS + slot * NODE_FEATS
        kind = str(node.get("node_type") or "")
        row[base + (0 if kind == "entity" else 1 if kind == "summary" else 2)] = 1.0
        for col, key in enumerate(_NUMERIC_KEYS, start=3):
            row[base + col] = float(node.get(key) or 0)
        shown_ids.append(str(node.get("node_id")))
    return shown_ids


def featurize_observation(obs: Dict[str, Any], k_nodes: int = 10) -> Tuple[np.ndarray, Dict[str, Any]]:
    row = np.zeros(GLOBAL_FEATS + NODE_FEATS * k_nodes, dtype=np.float64)
    shown_ids = _fill(row, obs, k_nodes)
    return row.astype(np.float32), {"top_node_ids": shown_ids}


def featurize_batch(observations: Sequence[Dict[str, Any]], k_nodes: int = 10) -> np.ndarray:
    """[B, 4 + 8 * k_nodes] float32 for a list of observations (host path of the vectorised environments)."""
    out = np.zeros((len(observations), GLOBAL_FEATS + NODE_FEATS * k_nodes), dtype=np.float64)
    for b, obs in enumerate(observations):
        _fill(out[b], obs, k_nodes)
    return out.astype(np.float32)
