"""Drop-in for ``ah_rag.agent.gym_env.AHRAGGymEnv`` (gym_env.py:11-125) plus its batched form.

Discrete actions: 0 expand_parents(top 2) | 1 expand_children(top 2) | 2 expand_related(top 1) |
3 commit_top(top 3) | 4 query_details_top(top 1) | 5 end_episode.  Observations are the featurizer's
float32[84]; rewards the dense step reward minus a repeat-action penalty.

``VecAHRAGGymEnv`` serves the rollout loops of collect_trajectories / ppo_train (policy_ppo.py:163-205):
``reset(query_embeddings)`` is one fused top-k + one rerank launch for all environments, ``step(actions)`` one
``ahr_expand`` launch for every environment that chose an expand action.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .environment import GraphEnvironment, VecGraphEnvironment
from .featurizer import featurize_observation
from .reward import step_reward
from ..graph.hierarchical_graph import HierarchicalGraph

N_ACTIONS = 6
_EXPAND = {0: ("parents", 2), 1: ("children", 2), 2: ("related", 1)}


def _action_mask(raw: Optional[Dict[str, Any]]) -> List[int]:
    """gym_env.py:54-75: without any top node only end_episode is valid."""
    if len((raw or {}).get("selection") or []) == 0:
        return [0, 0, 0, 0, 0, 1]
    return [1] * N_ACTIONS


def _top_ids(raw: Dict[str, Any], k: int) -> List[str]:
    return [n.get("node_id") for n in (raw.get("selection") or [])[:k] if n.get("node_id")]


class AHRAGGymEnv:
    def __init__(self, graph_dir: str = "graph", max_steps: int = 6, debug: bool = False,
                 enable_action_mask: bool = True, repeat_penalty: float = 0.02,
                 hg: Optional[HierarchicalGraph] = None) -> None:
        # the reference leaves file logging on here (gym_env.py:27); it is not a result, so it is off
        self.env = GraphEnvironment(graph_dir=graph_dir, debug=debug, log_level="off", redact=True,
                                    logging_enabled=False, hg=hg)
        self.max_steps = max(1, int(max_steps))
        self._cur_step = 0
        self._last_obs: Optional[Dict[str, Any]] = None
        self._last_raw_obs: Optional[Dict[str, Any]] = None
        self._query: Optional[str] = None
        self.enable_action_mask = bool(enable_action_mask)
        self.repeat_penalty = float(repeat_penalty)
        self._last_action: Optional[int] = None

    @property
    def action_size(self) -> int:
        return N_ACTIONS

    def get_action_mask(self) -> List[int]:
        return _action_mask(self._last_raw_obs)

    def _pick_top_ids(self, raw_obs: Dict[str, Any], k: int = 2) -> List[str]:
        return _top_ids(raw_obs, k)

    def _after_reset(self, raw_obs: Dict[str, Any]) -> Tuple[Any, Dict[str, Any]]:
        vec, aux = featurize_observation(raw_obs)
        self._last_obs = {"vec": vec, "aux": aux}
        self._last_raw_obs = raw_obs
        info: Dict[str, Any] = {"raw_obs": raw_obs, "aux": aux}
        if self.enable_action_mask:
            info["action_mask"] = self.get_action_mask()
        self._last_action = None
        return vec, info

    def reset(self, query: str) -> Tuple[Any, Dict[str, Any]]:
        self._query = query
        self._cur_step = 0
        raw_obs, _ = self.env.reset(seed_query=query)
        return self._after_reset(raw_obs)

    def _finish_step(self, action: int, prev_raw: Dict[str, Any], obs: Dict[str, Any], done: bool):
        r = 0.0 if done else step_reward(prev_raw, obs)
        if not done and self._last_action is not None and int(action) == int(self._last_action) and self.repeat_penalty > 0:
            r -= self.repeat_penalty
        done = done or (self._cur_step >= self.max_steps)
        vec, aux = featurize_observation(obs)
        self._last_obs = {"vec": vec, "aux": aux}
        self._last_raw_obs = obs
        info: Dict[str, Any] = {"action": int(action)}
        if self.enable_action_mask:
            info["action_mask"] = self.get_action_mask()
        self._last_action = int(action)
        info.update({"raw_obs": obs, "aux": aux, "step": self._cur_step})
        return vec, float(r), bool(done), info

    def _host_action(self, action: int, prev_raw: Dict[str, Any]) -> Tuple[Dict[str, Any], bool]:
        """Actions that never touch the device (commit / details / end)."""
        if action == 3:
            return self.env.commit_selection(_top_ids(prev_raw, 3))[0], False
        if action == 4:
            ids = _top_ids(prev_raw, 1)
            return (self.env.query_node_details(ids[0])[0] if ids else prev_raw), False
        return prev_raw, True

    def step(self, action: int) -> Tuple[Any, float, bool, Dict[str, Any]]:
        assert self._last_raw_obs is not None, "Call reset() first"
        self._cur_step += 1
        prev_raw = self._last_raw_obs
        if action in _EXPAND:
            op, k = _EXPAND[action]
            obs, done = getattr(self.env, f"expand_{op}")(_top_ids(prev_raw, k))[0], False
        else:
            obs, done = self._host_action(int(action), prev_raw)
        return self._finish_step(action, prev_raw, obs, done)


class VecAHRAGGymEnv:
    """B gym environments sharing one device-resident graph + index."""

    def __init__(self, hg: HierarchicalGraph, n_envs: int, max_steps: int = 6, repeat_penalty: float = 0.02) -> None:
        self.hg = hg
        self.gyms = [AHRAGGymEnv(hg=hg, max_steps=max_steps, repeat_penalty=repeat_penalty) for _ in range(n_envs)]
        self.vec = VecGraphEnvironment.__new__(VecGraphEnvironment)
        self.vec.hg = hg
        self.vec.envs = [g.env for g in self.gyms]

    def __len__(self) -> int:
        return len(self.gyms)

    def reset(self, query_embeddings, queries: Optional[Sequence[Any]] = None):
        """-> (obs [B, 84] float32, infos)."""
        outs = self.vec.reset_batch(query_embeddings, top_k=5, member_top_m=5, queries=queries)
        vecs, infos = [], []
        for g, (raw, _) in zip(self.gyms, outs):
            g._query = g.env.last_query
            g._cur_step = 0
            v, info = g._after_reset(raw)
            vecs.append(v)
            infos.append(info)
        return np.stack(vecs), infos

    def action_masks(self) -> np.ndarray:
        return np.asarray([g.get_action_mask() for g in self.gyms], dtype=np.int64)

    def step(self, actions: Sequence[int]):
        """-> (obs [B, 84], rewards [B], dones [B], infos); every expand action of the batch is one launch."""
        B = len(self.gyms)
        assert len(actions) == B
        prevs = [g._last_raw_obs for g in self.gyms]
        for g in self.gyms:
            g._cur_step += 1
        exp_idx = [b for b in range(B) if int(actions[b]) in _EXPAND]
        obs_list: List[Any] = [None] * B
        dones = [False] * B
        if exp_idx:
            ops = [_EXPAND[int(actions[b])][0] for b in exp_idx]
            ids = [_top_ids(prevs[b], _EXPAND[int(actions[b])][1]) for b in exp_idx]
            sub = VecGraphEnvironment.__new__(VecGraphEnvironment)
            sub.hg, sub.envs = self.hg, [self.gyms[b].env for b in exp_idx]
            for b, (o, _) in zip(exp_idx, sub.expand_batch(ops, ids, limit=10)):
                obs_list[b] = o
        for b in range(B):
            if obs_list[b] is None:
                obs_list[b], dones[b] = self.gyms[b]._host_action(int(actions[b]), prevs[b])
        vecs, rewards, done_out, infos = [], [], [], []
        for b, g in enumerate(self.gyms):
            v, r, d, info = g._finish_step(int(actions[b]), prevs[b], obs_list[b], dones[b])
            vecs.append(v); rewards.append(r); done_out.append(d); infos.append(info)
        return np.stack(vecs), np.asarray(rewards, dtype=np.float64), np.asarray(done_out), infos


class DeviceVecAHRAGGymEnv:
    """B gym environments whose whole state lives on the device: ``reset`` = fused top-k + rerank + one
    ``ahr_env_reset`` launch, ``step`` = ONE ``ahr_env_step`` launch (expansion walk, selection / frontier sets,
    reward, featurizer and action mask of gym_env.py:85-125 for every environment).  Nothing is copied to the
    host: features ``[B, 84]`` float32, rewards ``[B]`` float64, done / mask uint8 are CUDA tensors a policy
    network can consume directly (the rollout loop of policy_ppo.py:163-205 without its per-env Python).

    Per item the numbers equal ``AHRAGGymEnv`` driven with the same queries and actions (bit-exact, see
    tests/test_gpu_env.py).  Observations carry what the featurizer and reward read, not the text fields.
    """

    FEATS = 84

    def __init__(self, hg: HierarchicalGraph, n_envs: int, max_steps: int = 6, repeat_penalty: float = 0.02,
                 cap_selection: Optional[int] = None, cap_frontier: Optional[int] = None) -> None:
        from ..engine import EnvBatchState
        self.hg = hg
        self.csr = hg._device_graph()["csr"]
        self._batch_args = (n_envs, max_steps, repeat_penalty, cap_selection, cap_frontier)
        self.batch = EnvBatchState(self.csr, n_envs, max_steps, repeat_penalty, cap_selection, cap_frontier)
        self.n_envs, self.max_steps, self.repeat_penalty = self.batch.n_envs, self.batch.max_steps, self.batch.repeat_penalty

    def __len__(self) -> int:
        return self.n_envs

    @property
    def action_size(self) -> int:
        return N_ACTIONS

    @property
    def state(self):
        return self.batch.state

    def reset(self, query_embeddings):
        """``query_embeddings`` [B, d] -> (features [B, 84] float32, action mask [B, 6] uint8), device tensors."""
        import torch
        if len(query_embeddings) != self.n_envs:
            raise ValueError(f"expected {self.n_envs} query embeddings, got {len(query_embeddings)}")
        hg = self.hg
        dg = hg._device_graph()
        csr = dg["csr"]
        if csr is not self.csr:  # the graph changed since the last episode: new device graph, fresh environment state
            from ..engine import EnvBatchState
            self.csr = csr
            self.batch = EnvBatchState(csr, *self._batch_args)
        # GraphEnvironment.reset -> semantic_anchor(top_k=5, member_top_m=5), no weights / filters set (:155-241)
        params = hg._resolve(5, None, None, None, None, None, None, None)
        db_path, _ = hg._resolve_index(None, "unused")
        di = hg._device_index(db_path, dg)
        n_index = len(di["store"].ids)
        top_k = 5
        dev = self.csr.device
        if n_index == 0:  # un-built index: empty anchor, as the reference's empty collection returns
            out = {"nodes": torch.full((self.n_envs, top_k), -1, dtype=torch.int32, device=dev),
                   "semantic": torch.zeros((self.n_envs, top_k), dtype=torch.float64, device=dev),
                   "score": torch.zeros((self.n_envs, top_k), dtype=torch.float64, device=dev),
                   "count": torch.zeros((self.n_envs,), dtype=torch.int32, device=dev)}
            n_seeds = 0
        else:
            out, rows = hg._device_search(query_embeddings, top_k, params, False, db_path, dg, di)
            n_seeds = int(rows.shape[1])
        return self.batch.reset(out, n_seeds)

    def action_masks(self):
        return self.batch.mask

    def step(self, actions):
        """``actions`` [B] ints in 0..5 (device int32 tensor preferred) ->
        (features [B, 84], rewards [B] f64, done [B] u8, mask [B, 6] u8) -- device tensors, overwritten every step."""
        import torch
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions, dtype=np.int32))
        if tuple(a.shape) != (self.n_envs,):
            raise ValueError(f"actions must have shape ({self.n_envs},)")
        return self.batch.step(a.to(device=self.csr.device, dtype=torch.int32).contiguous())

    def check_overflow(self) -> None:
        self.batch.check_overflow()

    def _sets(self, ids_key: str, cnt_key: str) -> List[List[str]]:
        names = self.hg._device_graph()["arrays"]["node_ids"]
        ids, cnt = self.batch.state[ids_key].cpu().numpy(), self.batch.state[cnt_key].cpu().numpy()
        return [sorted(names[int(v)] for v in ids[b, :int(cnt[b])]) for b in range(self.n_envs)]

    def selection_sets(self) -> List[List[str]]:
        """Selected node ids per environment (host copy; what `end_episode` hands to the evaluator)."""
        return self._sets("sel_ids", "sel_cnt")

    def frontier_sets(self) -> List[List[str]]:
        return self._sets("fr_ids", "fr_cnt")
