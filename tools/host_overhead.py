"""Where the ~30 us of a graph call through Python go (host side, one B200 box): run by hand through gpurun."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ah_rag_b200 import engine, synth, csr as csr_mod  # noqa: E402
from ah_rag_b200._lib import check  # noqa: E402


def t(fn, n=5000):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return round(dt, 2)


dev = torch.device("cuda", 0)
import copy
from oracle import ref_search
nl = synth.make_graph(seed=5, n_entities=2000, n_l1=50, n_l2=5, n_hyperedges=400)
og = ref_search.OracleGraph(copy.deepcopy(nl))
g = engine.GraphCSR(csr_mod.csr_from_networkx(og.G))
nq, ns, k = 1024, 50, 50
seeds = torch.randint(0, g.n_nodes, (nq, ns), device=dev).to(torch.int32)
scores = torch.rand((nq, ns), device=dev).sort(dim=1, descending=True).values.contiguous()
prm = engine.SearchParams(member_top_m=5)
ops = torch.randint(0, 3, (nq,), device=dev).to(torch.int32)
ids = torch.randint(0, g.n_nodes, (nq,), device=dev).to(torch.int32)
offs = torch.arange(nq + 1, device=dev).to(torch.int32)
res = {
    "anchor_rerank(want_cluster=False)": t(lambda: g.anchor_rerank(seeds, scores, prm, k, want_cluster=False)),
    "expand": t(lambda: g.expand(ops, ids, offs, 10)),
    "torch.empty x1": t(lambda: torch.empty((nq, k), dtype=torch.int32, device=dev)),
    "params.to_c()": t(lambda: prm.to_c()),
    "with torch.cuda.device": t(lambda: torch.cuda.device(dev).__enter__()),
    "_stream_ptr": t(lambda: engine._stream_ptr(dev)),
    "data_ptr": t(lambda: seeds.data_ptr()),
    "ctypes call (ahr_launch_count)": t(lambda: g.lib.ahr_launch_count()),
    "_require_cuda": t(lambda: engine._require_cuda(seeds, "x")),
}
for k_, v in res.items():
    print(f"{k_:40s} {v:8.2f} us")
