"""GPU parity of the device-resident batched gym environment (ahr_env_reset / ahr_env_step) against
(1) the gym traces written by the UNMODIFIED reference (tests/golden, gym_env.py:41-125) and
(2) the oracle's OracleGymEnv on larger synthetic graphs with random action sequences.
Bit-exact: float32 feature vectors, float64 rewards, done flags, action masks, final selection / frontier sets."""
import copy

import numpy as np
import pytest

from ah_rag_b200 import synth
from oracle import ref_search
from tests import golden_util as gu
from tests.test_gpu_dropin import _make_hg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gu.fixture_names())
def test_device_env_matches_reference_gym_traces(name, tmp_path):
    import torch
    from ah_rag_b200.agent.gym_env import DeviceVecAHRAGGymEnv
    fx = gu.load(name)
    hg, _ = _make_hg(fx, tmp_path)
    eps = fx["gym"]
    env = DeviceVecAHRAGGymEnv(hg, len(eps))
    feats, mask = env.reset(np.stack([fx["query_vecs"][ep["query"]] for ep in eps]))
    assert feats.is_cuda and feats.dtype == torch.float32 and tuple(feats.shape) == (len(eps), 84)
    f, m = feats.cpu().numpy(), mask.cpu().numpy()
    for b, ep in enumerate(eps):
        assert [float(x) for x in f[b]] == ep["steps"][0]["vec"], (name, b)
        assert [int(x) for x in m[b]] == ep["steps"][0]["mask"]
    for t in range(1, max(len(ep["steps"]) for ep in eps)):
        acts = [ep["steps"][t]["action"] if t < len(ep["steps"]) else 5 for ep in eps]
        feats, rew, done, mask = env.step(torch.tensor(acts, dtype=torch.int32, device="cuda"))
        f, r, d, m = feats.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), mask.cpu().numpy()
        for b, ep in enumerate(eps):
            if t < len(ep["steps"]):
                st = ep["steps"][t]
                assert [float(x) for x in f[b]] == st["vec"], (name, b, t)
                assert float(r[b]) == st["reward"] and bool(d[b]) == st["done"], (name, b, t)
                assert [int(x) for x in m[b]] == st["mask"]
    env.check_overflow()


@pytest.mark.parametrize("seed,n_envs,steps", [(21, 96, 9), (22, 257, 14)])
def test_device_env_random_rollouts_match_oracle(seed, n_envs, steps, tmp_path):
    import torch
    from ah_rag_b200.agent.gym_env import DeviceVecAHRAGGymEnv
    nl = synth.make_graph(seed=seed, n_entities=400, n_l1=24, n_l2=4, n_hyperedges=260, related_density=0.3)
    rng = np.random.default_rng(seed)
    og = ref_search.OracleGraph(copy.deepcopy(nl))
    index_ids = [n for n, d in og.G.nodes(data=True) if d.get("node_type") in ("entity", "summary")]
    dim = 48
    rows = synth.unit_rows(len(index_ids), dim, seed + 1)
    qv = synth.unit_rows(n_envs, dim, seed + 2)
    queries = [f"q{i}" for i in range(n_envs)]
    fx = {"node_link": nl, "index_ids": index_ids, "index_rows": rows, "queries": queries, "query_vecs": qv}
    hg, _ = _make_hg(fx, tmp_path)
    env = DeviceVecAHRAGGymEnv(hg, n_envs, max_steps=steps, repeat_penalty=0.02)
    oracles = []
    for b in range(n_envs):
        oe = ref_search.OracleEnv(og, index_ids, rows)
        oracles.append(ref_search.OracleGymEnv(oe, max_steps=steps, repeat_penalty=0.02))
    feats, mask = env.reset(qv)
    f, m = feats.cpu().numpy(), mask.cpu().numpy()
    for b, o in enumerate(oracles):
        vec, raw = o.reset(qv[b])
        assert np.array_equal(f[b].view(np.uint32), vec.view(np.uint32)), b
        assert list(m[b]) == ([1] * 6 if raw["selection"] else [0, 0, 0, 0, 0, 1])
    # actions: mostly expansions and commits, a few details, no early end (action 5 closes the comparison at the end)
    for t in range(steps):
        acts = rng.choice(6, size=n_envs, p=[0.22, 0.22, 0.22, 0.2, 0.1, 0.04]).astype(np.int32)
        feats, rew, done, mask = env.step(torch.from_numpy(acts).cuda())
        f, r, d, m = feats.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), mask.cpu().numpy()
        for b, o in enumerate(oracles):
            vec, wr, wd, raw = o.step(int(acts[b]))
            assert np.array_equal(f[b].view(np.uint32), vec.view(np.uint32)), (b, t, int(acts[b]))
            assert float(r[b]) == wr and bool(d[b]) == wd, (b, t, int(acts[b]), float(r[b]), wr)
            assert list(m[b]) == ([1] * 6 if raw["selection"] else [0, 0, 0, 0, 0, 1])
    env.check_overflow()
    sel, fr = env.selection_sets(), env.frontier_sets()
    for b, o in enumerate(oracles):
        assert sel[b] == sorted(o.env.selection_set) and fr[b] == sorted(o.env.frontier_set), b


def test_device_env_capacity_overflow_is_reported(tmp_path):
    import torch
    from ah_rag_b200.agent.gym_env import DeviceVecAHRAGGymEnv
    fx = gu.load("g_dense_s11")
    hg, _ = _make_hg(fx, tmp_path)
    env = DeviceVecAHRAGGymEnv(hg, 4, max_steps=50, cap_selection=3, cap_frontier=5)
    env.reset(fx["query_vecs"][:4])
    for a in (1, 0, 2, 1, 0, 2):
        env.step(torch.full((4,), a, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="capacity"):
        env.check_overflow()
    with pytest.raises(ValueError):
        env.step([0, 1])


def test_collect_rollouts_matches_sequential_oracle_episodes(tmp_path):
    """agent.rollout.collect_rollouts (all episodes together, finished ones idling on end_episode) against the oracle gym
    environment stepped one episode at a time with the same actions, as collect_trajectories.py / ppo_train do."""
    import torch
    from ah_rag_b200.agent.gym_env import DeviceVecAHRAGGymEnv
    from ah_rag_b200.agent.rollout import collect_rollouts, random_policy
    fx = gu.load("g_dense_s11")
    hg, _ = _make_hg(fx, tmp_path)
    B = 48
    qv = np.stack([fx["query_vecs"][b % len(fx["query_vecs"])] for b in range(B)])
    venv = DeviceVecAHRAGGymEnv(hg, B, max_steps=6)
    gen = torch.Generator(device="cuda").manual_seed(123)
    out = collect_rollouts(venv, qv, policy=random_policy(gen))
    T = out["actions"].shape[0]
    assert 1 <= T <= 6 and tuple(out["obs"].shape) == (T, B, 84) and out["rewards"].dtype == torch.float64
    acts, rew, alive = out["actions"].cpu().numpy(), out["rewards"].cpu().numpy(), out["alive"].cpu().numpy()
    obs, dones = out["obs"].cpu().numpy(), out["dones"].cpu().numpy()
    og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
    for b in range(B):
        gym = ref_search.OracleGymEnv(ref_search.OracleEnv(og, fx["index_ids"], fx["index_rows"]), max_steps=6)
        vec, _ = gym.reset(qv[b])
        total, t, done = 0.0, 0, False
        while not done:
            assert alive[t, b] and np.array_equal(obs[t, b].view(np.uint32), vec.view(np.uint32)), (b, t)
            vec, r, done, _ = gym.step(int(acts[t, b]))
            assert float(rew[t, b]) == r and bool(dones[t, b]) == done, (b, t)
            total += r
            t += 1
        assert int(out["episode_length"][b]) == t and not alive[t:, b].any() and (acts[t:, b] == 5).all()
        assert float(out["episode_reward"][b]) == pytest.approx(total, abs=1e-12)
        assert np.array_equal(out["final_obs"][b].cpu().numpy().view(np.uint32), vec.view(np.uint32))


def test_rollout_captured_in_a_cuda_graph_equals_step_by_step(tmp_path):
    """EnvBatchState.capture_rollout: reset + T steps as one CUDA graph; a replay (also with NEW actions written into the
    captured buffer) reproduces, bit for bit, what the same calls give when issued one by one."""
    import torch
    from ah_rag_b200 import engine
    nl = synth.make_graph(seed=33, n_entities=300, n_l1=20, n_l2=3, n_hyperedges=200, related_density=0.3)
    og = ref_search.OracleGraph(copy.deepcopy(nl))
    from ah_rag_b200 import csr as csr_mod
    arrays = csr_mod.csr_from_networkx(og.G)
    g = engine.GraphCSR(arrays)
    rng = np.random.default_rng(5)
    B, T = 130, 6
    idx_nodes = np.array([i for i, t in enumerate(arrays["node_type"]) if t in (0, 1)], dtype=np.int32)
    seeds = torch.from_numpy(np.stack([rng.choice(idx_nodes, 5, replace=False) for _ in range(B)])).cuda()
    scores = torch.from_numpy(np.sort(rng.uniform(0.1, 0.9, (B, 5)).astype(np.float32), axis=1)[:, ::-1].copy()).cuda()
    rr = g.anchor_rerank(seeds, scores, engine.SearchParams(member_top_m=5), 5, want_cluster=False)
    for trial in range(2):
        acts = torch.from_numpy(rng.integers(0, 5, (T, B)).astype(np.int32)).cuda()
        ref_env = engine.EnvBatchState(g, B, max_steps=T)
        want = {"features": [], "rewards": [], "done": [], "mask": []}
        f, m = ref_env.reset(rr, 5)
        want["features"].append(f.clone()); want["mask"].append(m.clone())
        for t in range(T):
            f, r, d, m = ref_env.step(acts[t])
            want["features"].append(f.clone()); want["mask"].append(m.clone())
            want["rewards"].append(r.clone()); want["done"].append(d.clone())
        if trial == 0:
            env = engine.EnvBatchState(g, B, max_steps=T)
            buf = acts.clone()
            replay, out = env.capture_rollout(rr, 5, buf)
        else:
            buf.copy_(acts)  # a policy writing new actions into the captured buffer
        replay()
        torch.cuda.synchronize()
        assert torch.equal(out["features"].view(torch.int32), torch.stack(want["features"]).view(torch.int32))
        assert torch.equal(out["mask"], torch.stack(want["mask"]))
        assert torch.equal(out["rewards"], torch.stack(want["rewards"])) and torch.equal(out["done"], torch.stack(want["done"]))
        env.check_overflow()
