"""Batched rollout collection over the device-resident gym environments.

The reference collects experience one environment at a time: ``scripts/collect_trajectories.py:31-53`` (one
``AHRAGGymEnv``, uniform random actions until ``done``) and the rollout half of ``ppo_train``
(``src/ah_rag/agent/policy_ppo.py:163-205``: ``n_envs`` environments rolled sequentially, a masked policy per step).
Here all B episodes advance together -- one ``ahr_env_step`` launch per step -- and everything a learner needs comes back
as device tensors; nothing is copied to the host inside the loop.

An episode that has ended keeps being fed action 5 (``end_episode``: no state change, zero reward) until the whole batch is
done; ``alive[t, b]`` says whether step ``t`` belonged to episode ``b``, so per-environment results equal what the
sequential loops of the reference produce for the same actions.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

END_EPISODE = 5


def random_policy(generator=None) -> Callable:
    """Uniform over the 6 actions, mask ignored -- what collect_trajectories.py:45 does (``random.randrange``)."""
    import torch

    def act(obs, mask):
        return torch.randint(0, 6, (obs.shape[0],), device=obs.device, generator=generator, dtype=torch.int64).to(torch.int32)

    return act


def masked_random_policy(generator=None) -> Callable:
    """Uniform over the actions the environment's mask allows (gym_env.py:54-75)."""
    import torch

    def act(obs, mask):
        w = mask.to(torch.float32)
        return torch.multinomial(w, 1, generator=generator).view(-1).to(torch.int32)

    return act


def collect_rollouts(venv, query_embeddings, policy: Optional[Callable] = None, max_steps: Optional[int] = None) -> Dict[str, Any]:
    """Reset ``venv`` (a ``DeviceVecAHRAGGymEnv``) on ``query_embeddings`` [B, d] and roll every episode to its end.

    ``policy(obs [B, 84] float32, mask [B, 6] uint8) -> actions [B] int32`` on the device (default: uniform random).
    Returns device tensors, T = number of steps taken by the longest episode (<= max_steps):
      obs [T, B, 84] (the observation each action was taken on), actions [T, B] int32, rewards [T, B] float64,
      dones [T, B] bool, alive [T, B] bool, final_obs [B, 84], episode_reward [B] float64, episode_length [B] int32.
    """
    import torch
    policy = policy or random_policy()
    limit = int(max_steps if max_steps is not None else venv.max_steps)
    obs, mask = venv.reset(query_embeddings)
    B = obs.shape[0]
    dev = obs.device
    alive = torch.ones((B,), dtype=torch.bool, device=dev)
    end = torch.full((B,), END_EPISODE, dtype=torch.int32, device=dev)
    rec: Dict[str, list] = {"obs": [], "actions": [], "rewards": [], "dones": [], "alive": []}
    for _ in range(limit):
        actions = torch.where(alive, policy(obs, mask).to(torch.int32), end)
        rec["obs"].append(obs.clone())
        rec["actions"].append(actions)
        rec["alive"].append(alive.clone())
        obs, rewards, done, mask = venv.step(actions)
        rec["rewards"].append(torch.where(alive, rewards, torch.zeros_like(rewards)))
        done_b = done.to(torch.bool)
        rec["dones"].append(done_b & alive)
        alive = alive & ~done_b
        if not bool(alive.any()):  # the only host synchronisation of the loop: one flag per step
            break
    out = {k: torch.stack(v) if v else torch.empty((0, B), device=dev) for k, v in rec.items()}
    out["final_obs"] = obs.clone()
    out["episode_reward"] = out["rewards"].sum(dim=0) if rec["rewards"] else torch.zeros((B,), dtype=torch.float64, device=dev)
    out["episode_length"] = out["alive"].sum(dim=0).to(torch.int32) if rec["alive"] else torch.zeros((B,), dtype=torch.int32, device=dev)
    return out
