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
        import ctypes
        import torch
        from .. import _lib
        self._torch, self._ctypes, self._libmod = torch, ctypes, _lib
        self.hg = hg
        self.n_envs = int(n_envs)
        self.max_steps = max(1, int(max_steps))
        self.repeat_penalty = float(repeat_penalty)
        self.csr = hg._device_graph()["csr"]
        dev = self.csr.device
        B = self.n_envs
        # an episode adds at most 3 selected and 10 frontier nodes per step (gym_env.py:94-100)
        self.cap_sel = int(cap_selection or max(3 * (self.max_steps + 2), 32))
        self.cap_fr = int(cap_frontier or (10 + 10 * (self.max_steps + 2)))
        i32 = dict(dtype=torch.int32, device=dev)
        self.state = {
            "sel_ids": torch.zeros((B, self.cap_sel), **i32), "sel_cnt": torch.zeros((B,), **i32),
            "fr_ids": torch.zeros((B, self.cap_fr), **i32), "fr_cnt": torch.zeros((B,), **i32),
            "shown_ids": torch.full((B, 10), -1, **i32),
            "shown_score": torch.zeros((B, 10), dtype=torch.float64, device=dev),
            "shown_sem": torch.zeros((B, 10), dtype=torch.float64, device=dev),
            "shown_cnt": torch.zeros((B,), **i32), "n_seeds": torch.zeros((B,), **i32),
            "obs_step": torch.zeros((B,), **i32), "obs_sel": torch.zeros((B,), **i32),
            "obs_fr": torch.zeros((B,), **i32), "env_steps": torch.zeros((B,), **i32),
            "gym_steps": torch.zeros((B,), **i32), "last_action": torch.full((B,), -1, **i32),
            "done": torch.zeros((B,), dtype=torch.uint8, device=dev), "overflow": torch.zeros((1,), **i32),
        }
        self.features = torch.zeros((B, self.FEATS), dtype=torch.float32, device=dev)
        self.mask = torch.zeros((B, N_ACTIONS), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((B,), dtype=torch.float64, device=dev)
        self.done = torch.zeros((B,), dtype=torch.uint8, device=dev)
        eb = _lib.EnvBatch()
        eb.n_env, eb.cap_sel, eb.cap_fr, eb.reserved = B, self.cap_sel, self.cap_fr, 0
        for name, t in self.state.items():
            setattr(eb, name, t.data_ptr())
        self._eb = eb
        self._ready = False

    def __len__(self) -> int:
        return self.n_envs

    @property
    def action_size(self) -> int:
        return N_ACTIONS

    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.csr.device).cuda_stream

    def reset(self, query_embeddings):
        """``query_embeddings`` [B, d] -> (features [B, 84] float32, action mask [B, 6] uint8), device tensors."""
        torch, ct = self._torch, self._ctypes
        if len(query_embeddings) != self.n_envs:
            raise ValueError(f"expected {self.n_envs} query embeddings, got {len(query_embeddings)}")
        hg = self.hg
        # GraphEnvironment.reset -> semantic_anchor(top_k=5, member_top_m=5), no weights / filters set (:155-241)
        params = hg._resolve(5, None, None, None, None, None, None, None)
        db_path, _ = hg._resolve_index(None, "unused")
        n_index = len(hg._device_index(db_path)["store"].ids)
        top_k = 5
        dev = self.csr.device
        if n_index == 0:
            out = {"nodes": torch.full((self.n_envs, top_k), -1, dtype=torch.int32, device=dev),
                   "semantic": torch.zeros((self.n_envs, top_k), dtype=torch.float64, device=dev),
                   "score": torch.zeros((self.n_envs, top_k), dtype=torch.float64, device=dev),
                   "count": torch.zeros((self.n_envs,), dtype=torch.int32, device=dev)}
            n_seeds = 0
        else:
            out, rows = hg._device_search(query_embeddings, top_k, params, False, db_path)
            n_seeds = int(rows.shape[1])
        if self.n_envs:
            with torch.cuda.device(dev):
                self._libmod.check(self.csr.lib.ahr_env_reset(
                    self.csr._h, ct.byref(self._eb), out["nodes"].data_ptr(), out["semantic"].data_ptr(),
                    out["score"].data_ptr(), out["count"].data_ptr(), top_k, n_seeds, self.features.data_ptr(),
                    self.mask.data_ptr(), self._stream()), "ahr_env_reset")
        self._ready = True
        return self.features, self.mask

    def action_masks(self):
        return self.mask

    def step(self, actions):
        """``actions`` [B] ints in 0..5 (device int32 tensor preferred) ->
        (features [B, 84], rewards [B] f64, done [B] u8, mask [B, 6] u8) -- device tensors, overwritten every step."""
        torch, ct = self._torch, self._ctypes
        if not self._ready:
            raise RuntimeError("Call reset() first")
        dev = self.csr.device
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions, dtype=np.int32))
        a = a.to(device=dev, dtype=torch.int32).contiguous()
        if a.shape != (self.n_envs,):
            raise ValueError(f"actions must have shape ({self.n_envs},)")
        if self.n_envs:
            with torch.cuda.device(dev):
                self._libmod.check(self.csr.lib.ahr_env_step(
                    self.csr._h, ct.byref(self._eb), a.data_ptr(), self.max_steps, self.repeat_penalty,
                    self.features.data_ptr(), self.rewards.data_ptr(), self.done.data_ptr(), self.mask.data_ptr(),
                    self._stream()), "ahr_env_step")
        return self.features, self.rewards, self.done, self.mask

    def check_overflow(self) -> None:
        """Raises if a selection / frontier list outgrew its capacity (synchronises)."""
        if int(self.state["overflow"].item()):
            raise RuntimeError("DeviceVecAHRAGGymEnv: selection/frontier capacity exceeded; raise cap_selection/cap_frontier")

    def selection_sets(self) -> List[List[str]]:
        """Selected node ids per environment (host copy; what `end_episode` hands to the evaluator)."""
        ids = self.hg._device_graph()["arrays"]["node_ids"]
        sel, cnt = self.state["sel_ids"].cpu().numpy(), self.state["sel_cnt"].cpu().numpy()
        return [sorted(ids[int(v)] for v in sel[b, :int(cnt[b])]) for b in range(self.n_envs)]

    def frontier_sets(self) -> List[List[str]]:
        ids = self.hg._device_graph()["arrays"]["node_ids"]
        fr, cnt = self.state["fr_ids"].cpu().numpy(), self.state["fr_cnt"].cpu().numpy()
        return [sorted(ids[int(v)] for v in fr[b, :int(cnt[b])]) for b in range(self.n_envs)]
