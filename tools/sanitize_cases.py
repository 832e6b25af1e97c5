#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (one tool per gpurun call, B200_PROFILING.md):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py graph select
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py graph select
    compute-sanitizer --tool synccheck python tools/sanitize_cases.py graph select
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py k1            # tcgen05 / TMA kernels, own call

Every case also checks its result against the oracle, so a sanitizer-clean run is a correct run."""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ah_rag_b200 import csr as csr_mod  # noqa: E402
from ah_rag_b200 import engine, synth  # noqa: E402
from oracle import array_graph, exact, ref_search  # noqa: E402


def case_graph():
    nl = synth.make_graph(seed=5, n_entities=200, n_l1=10, n_l2=2, n_hyperedges=80)
    og = ref_search.OracleGraph(copy.deepcopy(nl))
    a = csr_mod.csr_from_networkx(og.G)
    g = engine.GraphCSR(a)
    rng = np.random.default_rng(0)
    n = a["n_nodes"]
    sn = np.stack([rng.permutation(n)[:8] for _ in range(24)]).astype(np.int32)
    ss = np.sort(rng.uniform(0.1, 0.9, (24, 8)).astype(np.float32), axis=1)[:, ::-1].copy()
    for m, top_k in ((5, 8), (200, 40)):  # shared-memory map, then the global-workspace variant (cap = 8 * 201 > 1024)
        out = g.anchor_rerank(torch.from_numpy(sn).cuda(), torch.from_numpy(ss).cuda(), engine.SearchParams(member_top_m=m), top_k)
        o = {k_: v.cpu().numpy() for k_, v in out.items()}
        for q in range(24):
            want = array_graph.anchor_rerank(a, sn[q], ss[q], top_k, m)
            got = [(int(o["nodes"][q, j]), float(o["semantic"][q, j]), float(o["score"][q, j])) for j in range(o["count"][q])]
            assert got == want, ("anchor_rerank", m, q)
    ops = torch.tensor([0, 1, 2] * 16, dtype=torch.int32).cuda()
    ids = torch.from_numpy(rng.integers(0, n, 96).astype(np.int32)).cuda()
    offs = torch.arange(0, 97, 2, dtype=torch.int32).cuda()
    for limit in (10, 200):  # warp walker, hash-set kernel
        oi, oc = g.expand(ops, ids, offs, limit)
        oi, oc, idh = oi.cpu().numpy(), oc.cpu().numpy(), ids.cpu().numpy()
        for e in range(48):
            assert list(oi[e, :oc[e]]) == array_graph.expand(a, e % 3, idh[2 * e:2 * e + 2], limit), ("expand", limit, e)
    rr = g.anchor_rerank(torch.from_numpy(sn[:, :5].copy()).cuda(), torch.from_numpy(ss[:, :5].copy()).cuda(),
                         engine.SearchParams(member_top_m=5), 5, want_cluster=False)
    envb = engine.EnvBatchState(g, 24, max_steps=6)
    envb.reset(rr, 5)
    for t in range(6):
        envb.step(torch.from_numpy(rng.integers(0, 6, 24).astype(np.int32)).cuda())
    envb.check_overflow()
    sums = [i for i, t in enumerate(a["node_type"]) if t == 1]
    g.lca_closure(sums[:2])
    world = 3
    b = [n * r // world for r in range(world)] + [n]
    shards = [engine.GraphShard(a, b[r], b[r + 1]) for r in range(world)]
    db = torch.tensor(b, dtype=torch.int64).cuda()
    g1 = torch.stack([s.stage1(ops, ids, offs, 2, 10) for s in shards])
    g2 = torch.stack([s.stage2(world, db, ops, ids, offs, 2, 10, g1) for s in shards])
    oi2, oc2 = shards[0].merge(world, db, ops, ids, offs, 2, 10, g1, g2)
    wi, wc = g.expand(ops, ids, offs, 10)
    assert torch.equal(oi2, wi) and torch.equal(oc2, wc), "sharded expand"
    torch.cuda.synchronize()
    print("graph cases ok")


def case_select():
    rng = np.random.default_rng(1)
    n, dim, nq, k = 3000, 64, 40, 20
    Xn, Qn = synth.unit_rows(n, dim, 1), synth.unit_rows(nq, dim, 2)
    X, Q = torch.from_numpy(Xn).cuda(), torch.from_numpy(Qn).cuda()
    idx = engine.VectorIndex(X)  # n <= 4096: the fp64 brute-force scan + bitonic selection
    s, i = idx.topk(Q, k, mode="exact")
    rs, ri = exact.exact_topk(Xn, Qn, k)
    assert np.array_equal(i.cpu().numpy(), ri)
    parts = []
    for lo, hi in ((0, 1000), (1000, 1700), (1700, 3000)):
        parts.append(engine.VectorIndex(X[lo:hi].contiguous(), id_base=lo).topk(Q, k, mode="exact"))
    ms, mi = engine.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))  # rank-counting merge
    assert np.array_equal(mi.cpu().numpy(), ri)
    big = [engine.VectorIndex(X[lo:lo + 150].contiguous(), id_base=lo).topk(Q, k, mode="exact") for lo in range(0, 3000, 150)]
    ms, mi = engine.merge_topk(torch.stack([p[0] for p in big]), torch.stack([p[1] for p in big]))  # 20 lists: bitonic selection
    assert np.array_equal(mi.cpu().numpy(), ri)
    ex = engine.PeerExchange(torch.device("cuda", 0), 0, 1, nq * k * 12, lambda h: [h])
    s2, i2 = ex.topk_sharded(idx, Q, k, "exact")
    assert np.array_equal(i2.cpu().numpy(), ri)
    torch.cuda.synchronize()
    print("select cases ok")


def case_k1():
    n, dim, k = 9000, 128, 50
    Xn = synth.unit_rows(n, dim, 3, clustered=16)
    X = torch.from_numpy(Xn).cuda().to(torch.bfloat16).contiguous()
    bits = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    idx = engine.VectorIndex(X)
    for nq in (64, 300):  # one CTA per tile, CTA pairs
        Q = torch.from_numpy(synth.unit_rows(nq, dim, 4, clustered=16)).cuda().to(torch.bfloat16).contiguous()
        rs, ri = exact.exact_topk(bits(X), bits(Q), k)
        s, i = idx.topk(Q, k, mode="bf16")
        assert np.mean([len(set(a) & set(b)) / k for a, b in zip(i.cpu().numpy(), ri)]) >= 0.999
        s, i = idx.topk(Q, k, mode="exact")
        assert np.array_equal(i.cpu().numpy(), ri)
    torch.cuda.synchronize()
    print("k1 cases ok")


if __name__ == "__main__":
    for w in sys.argv[1:] or ["graph", "select"]:
        {"graph": case_graph, "select": case_select, "k1": case_k1}[w]()
