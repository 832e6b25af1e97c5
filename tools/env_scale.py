"""Gym rollout (reset + 6 steps) and hop 2 (expand, limit 10) at 1024 / 65536 environments on the c5 graph."""
import json
import sys
import torch
sys.path.insert(0, ".")
import bench_workloads as bw
from ah_rag_b200 import engine
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
a = bw.make_c5_graph(dev, float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
g = bw.DeviceGraph(a, dev)
prm = engine.SearchParams(member_top_m=5)
res = {}
for n_env in (1024, 65536):
    seeds, scores, ops = bw.c5_inputs(a, n_env, 5, dev, seed=23)
    rr5 = g.anchor_rerank(seeds, scores, prm, 5, want_cluster=False)
    gen = torch.Generator(device=dev).manual_seed(5)
    acts = [torch.randint(0, 5, (n_env,), generator=gen, device=dev).to(torch.int32) for _ in range(6)]
    envb = engine.EnvBatchState(g, n_env, max_steps=6)

    def rollout():
        envb.reset(rr5, 5)
        for t in range(6):
            envb.step(acts[t])

    ms = bw.timed(rollout, steps=20, warmup=3)
    envb.check_overflow()
    take = torch.where(ops == 2, 1, 2)
    offs = torch.zeros(n_env + 1, dtype=torch.int32, device=dev)
    offs[1:] = torch.cumsum(take, 0).to(torch.int32)
    keep = torch.arange(2, device=dev)[None, :] < take[:, None]
    flat = rr5["nodes"][:, :2][keep].to(torch.int32).contiguous()
    ms2 = bw.timed(lambda: g.expand(ops, flat, offs, 10), steps=20, warmup=3)
    res[n_env] = {"rollout_ms": ms, "env_steps_per_s": n_env * 6 / ms * 1e3, "hop2_ms": ms2, "hop2_envs_per_s": n_env / ms2 * 1e3}
print(json.dumps(res))
