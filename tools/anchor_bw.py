"""Hop 1 at a machine-filling batch on the c5 graph (65536 queries x 50 seeds): time it, or profile it under ncu
(-k regex:anchor_rerank_kernel -s 3 -c 1)."""
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
nq, ns, k = 65536, 50, 50
seeds, scores, _ = bw.c5_inputs(a, nq, ns, dev, seed=17)
prm = engine.SearchParams(member_top_m=5)
ms = bw.timed(lambda: g.anchor_rerank(seeds, scores, prm, k, want_cluster=False), steps=10, warmup=3)
ms2 = bw.timed(lambda: g.anchor_rerank(seeds, scores, prm, k, want_cluster=True), steps=10, warmup=3)
print(json.dumps({"hop1_bw_ms": ms, "seeds_per_s": nq * ns / ms * 1e3, "with_cluster_outputs_ms": ms2}))
