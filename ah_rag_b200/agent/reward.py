"""Dense step reward / final reward (drop-in for ah_rag/agent/reward.py:6-38)."""
from __future__ import annotations

from typing import Any, Dict, Optional


def step_reward(prev_obs: Optional[Dict[str, Any]], cur_obs: Dict[str, Any]) -> float:
    """+1 per node newly committed to the selection, +0.05 per net new frontier id (at most 10), -0.05 per step."""
    if prev_obs is None:
        return 0.0
    ps, cs = prev_obs.get("state") or {}, cur_obs.get("state") or {}
    new_sel = len(set(cs.get("selection_ids") or []) - set(ps.get("selection_ids") or []))
    grown = max(0, len(set(cs.get("frontier_ids") or [])) - len(set(ps.get("frontier_ids") or [])))
    return float(1.0 * new_sel + 0.05 * min(grown, 10) - 0.05)


def final_reward(metrics: Dict[str, float]) -> float:
    return (0.4 * float(metrics.get("f1", 0.0)) + 0.3 * float(metrics.get("faithfulness", 0.0))
            + 0.2 * float(metrics.get("answer_relevancy", 0.0)) + 0.1 * float(metrics.get("contextual_recall", 0.0)))
