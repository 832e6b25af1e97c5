"""Dense step reward / final reward (drop-in for ah_rag/agent/reward.py:6-38).

step:  +1.0 for every id that is in the current observation's ``state.selection_ids`` but not in the previous one's,
       +0.05 for every id by which ``state.frontier_ids`` grew (growth counted on the list sizes, capped at 10),
       -0.05 flat.  An observation without a ``state`` (query_node_details) counts as two empty lists.
The arithmetic keeps the reference's evaluation order ((1.0 * a + 0.05 * b) - 0.05 in float64); the device twin in
csrc/env.cu uses the same order without FMA contraction.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

_FINAL_WEIGHTS = (("f1", 0.4), ("faithfulness", 0.3), ("answer_relevancy", 0.2), ("contextual_recall", 0.1))


def _id_sets(obs: Dict[str, Any]):
    state = obs.get("state") or {}
    return set(state.get("selection_ids") or ()), set(state.get("frontier_ids") or ())


def step_reward(prev_obs: Optional[Dict[str, Any]], cur_obs: Dict[str, Any]) -> float:
    if prev_obs is None:
        return 0.0
    sel_before, frontier_before = _id_sets(prev_obs)
    sel_now, frontier_now = _id_sets(cur_obs)
    newly_selected = len(sel_now.difference(sel_before))
    frontier_growth = min(max(len(frontier_now) - len(frontier_before), 0), 10)
    return float(1.0 * newly_selected + 0.05 * frontier_growth - 0.05)


def final_reward(metrics: Dict[str, float]) -> float:
    """0.4 f1 + 0.3 faithfulness + 0.2 answer_relevancy + 0.1 contextual_recall (missing metrics count as 0)."""
    total = 0.0
    for name, weight in _FINAL_WEIGHTS:
        total += weight * float(metrics.get(name, 0.0))
    return total
