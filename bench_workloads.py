#!/usr/bin/env python
"""bench_workloads.py -- the BASELINE.json configurations other than the headline (which is bench.py's job).

    python bench_workloads.py [c3x] [c4] [c5] [c2] [qsweep] [shard8] [adv]     (default: c3x c4 c5 c2; one GPU)
    torchrun --nproc-per-node N bench_workloads.py c4dist c5dist            (the sharded configurations, N = 2/4/8)

  c3x  c3 in fp32-EXACT mode: bit-exact ids at tensor-core speed (candidates + canonical rescoring + certification)
  c4   one shard of the 100M x 768 corpus at 8 GPUs (12.5M rows), 256-query batch: the HBM-bound regime of K1
  c5   synthetic hierarchical graph, 20M nodes / ~500M edges: anchor expansion + rerank for 1024 environments x 50
       seeds (hop 1) and one frontier expansion per environment (hop 2), checked on a sample against the array oracle
  c2   HotpotQA-scale synthetic graph (the shipped graph is empty): fp32 index, all 7405 dev-question-sized batch,
       exact top-5 + anchor expansion/rerank per query (the naive_rag / semantic_anchor retrieval step)

Each line printed is one JSON object.  Peaks from MEASURED_PEAKS.json.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ah_rag_b200 import engine  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


def unit_bf16(n, dim, seed, dev, chunk=500_000):
    out = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, n, chunk):
        g = torch.Generator(device=dev).manual_seed(seed + c0 // chunk)
        x = torch.randn((min(chunk, n - c0), dim), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        out[c0:c0 + x.shape[0]] = x.to(torch.bfloat16)
    return out


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def timed_graph(fn, steps=20, warmup=3, inner=8):
    """Device time of one call of ``fn``: ``inner`` calls captured into a CUDA graph (the library only enqueues on the
    caller's stream), the graph replayed ``steps`` times between two events.  Takes the Python / ctypes launch cost
    (~30 us per call here, more than the 20-30 us kernels being measured) out of the figure."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(inner):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (steps * inner)


def run_c3x(dev):
    n, dim, nq, k = 10_000_000, 1024, 4096, 50
    X, Q = unit_bf16(n, dim, 1234, dev), unit_bf16(nq, dim, 5678, dev)
    idx = engine.VectorIndex(X)
    ms = timed(lambda: idx.topk(Q, k, mode="exact"), steps=3, warmup=2)
    cert, fb = idx.exact_stats()
    s, i = idx.topk(Q, k, mode="exact")
    # spot-check 8 queries bit-exactly against the C oracle (full scan of the 10M rows on the host)
    from oracle import exact
    sel = [0, 1, 777, 2048, 4095]
    bits = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    rs, ri = exact.exact_topk(bits(X), bits(Q[sel]), k)
    ok = bool(np.array_equal(i[sel].cpu().numpy(), ri) and np.array_equal(s[sel].cpu().numpy().view(np.uint32), rs.view(np.uint32)))
    print(json.dumps({"workload": "c3x: 10M x 1024 bf16, Q=4096, k=50, fp32-exact mode", "ms": ms, "queries_per_s": nq / ms * 1e3,
                      "certified": cert, "fallback": fb, "bit_exact_vs_oracle_on_sample": ok, "sample_queries": sel}))
    del idx, X, Q


def run_c3c(dev):
    """SURVEY 8(d)'s CLUSTERED variant of c3 (seed 91011): 4096 centroids, rows = centroid + 0.3 * noise, queries near random
    rows -- every query has ~2400 cluster mates with scores a few 1e-5 apart around rank 50, so recall@k of the bf16 mode
    against the fp32-exact answer is not trivially 1 (on the N(0,1) corpus it is).  The exact mode itself is checked bit for
    bit against the C oracle (full host scan) on sampled queries."""
    n, dim, nq, k, n_cent = 10_000_000, 1024, 4096, 50, 4096
    g = torch.Generator(device=dev).manual_seed(91011)
    cent = torch.randn((n_cent, dim), generator=g, device=dev)
    X = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, n, 500_000):
        m = min(500_000, n - c0)
        x = cent[torch.randint(0, n_cent, (m,), generator=g, device=dev)] + 0.3 * torch.randn((m, dim), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        X[c0:c0 + m] = x.to(torch.bfloat16)
    pick = torch.randint(0, n, (nq,), generator=g, device=dev)
    q = X[pick].float() + 0.1 * torch.randn((nq, dim), generator=g, device=dev) / dim ** 0.5
    Q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()
    idx = engine.VectorIndex(X)
    ms_b = timed(lambda: idx.topk(Q, k, mode="bf16"), steps=5, warmup=2)
    ms_x = timed(lambda: idx.topk(Q, k, mode="exact"), steps=3, warmup=1)
    cert, fb = idx.exact_stats()
    sb, ib = idx.topk(Q, k, mode="bf16")
    sx, ix = idx.topk(Q, k, mode="exact")
    hit = (ib[:, :, None] == ix[:, None, :]).any(dim=2).float().sum(dim=1)  # per query: |bf16 top-k AND exact top-k|
    rel = ((sb - sx).abs() / sx.abs().clamp_min(1e-30)).max().item()  # rank-wise score error (same rank, maybe another row)
    from oracle import exact
    sel = [0, 1, 2048, 4095]
    bits = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    rs, ri = exact.exact_topk_rowpar(bits(X), bits(Q[sel]), k)
    ok = bool(np.array_equal(ix[sel].cpu().numpy(), ri) and np.array_equal(sx[sel].cpu().numpy().view(np.uint32), rs.view(np.uint32)))
    gap = (sx[:, k - 2] - sx[:, k - 1]).float()
    print(json.dumps({"workload": "c3 clustered: 10M x 1024 bf16 around 4096 centroids (noise 0.3), 4096 queries near random rows, k=50",
                      "bf16_ms": ms_b, "bf16_queries_per_s": nq / ms_b * 1e3, "exact_ms": ms_x, "exact_queries_per_s": nq / ms_x * 1e3,
                      "exact_certified_tier1": cert, "exact_fallback": fb, "exact_bit_exact_vs_oracle_on_sample": ok,
                      "bf16_recall_at_k_vs_exact": float(hit.mean().item()) / k, "bf16_queries_with_full_recall": int((hit == k).sum().item()),
                      "bf16_worst_query_recall": float(hit.min().item()) / k, "bf16_max_rel_score_err_rankwise": rel,
                      "median_score_gap_rank49_50": float(gap.median().item())}), flush=True)
    del idx, X, Q


def run_c4(dev):
    n, dim, nq, k = 12_500_000, 768, 256, 50
    hbm, tf, src = peaks()
    X, Q = unit_bf16(n, dim, 1234, dev), unit_bf16(nq, dim, 5678, dev)
    idx = engine.VectorIndex(X)
    idx.set_profiling(True)
    ms = timed(lambda: idx.topk(Q, k, mode="bf16"), steps=10, warmup=3)
    prof = idx.get_profile()
    k1 = prof["k1_ms_total"] / max(1, prof["k1_launches"])
    gbs = 2.0 * n * dim / (k1 * 1e-3) / 1e9
    tfl = 2.0 * nq * n * dim / (k1 * 1e-3) / 1e12
    print(json.dumps({"workload": "c4 shard: 12.5M x 768 bf16 (1/8 of 100M), Q=256, k=50, bf16 mode", "ms": ms, "k1_ms": k1,
                      "queries_per_s_per_gpu": nq / ms * 1e3, "n_slabs": prof["n_slabs"],
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                   "tensor_tflops": tfl, "tensor_frac": tfl / tf, "peak_source": src,
                                   "t_mem_ms": 2.0 * n * dim / (hbm * 1e9) * 1e3,
                                   "t_mma_ms": 2.0 * nq * n * dim / (tf * 1e12) * 1e3}}))
    del idx, X, Q


def run_qsweep(dev):
    """K1 over one c4 shard (12.5M x 768 bf16) for query batches from 1 to 1024: which roofline binds where."""
    n, dim, k = 12_500_000, 768, 50
    hbm, tf, src = peaks()
    X = unit_bf16(n, dim, 1234, dev)
    Qall = unit_bf16(1024, dim, 5678, dev)
    for variant in os.environ.get("AHR_SWEEP_VARIANTS", "2,1").split(","):
        os.environ["AHR_TC_VARIANT"] = variant
        idx = engine.VectorIndex(X)
        for nq in [int(x) for x in os.environ.get("AHR_SWEEP_NQ", "1,8,32,64,128,256,512,1024").split(",")]:
            Q = Qall[:nq].contiguous()
            idx.set_profiling(True)
            ms = timed(lambda: idx.topk(Q, k, mode="bf16"), steps=5, warmup=2)
            prof = idx.get_profile()
            idx.set_profiling(False)
            k1 = prof["k1_ms_total"] / max(1, prof["k1_launches"])
            print(json.dumps({"workload": "qsweep 12.5M x 768 bf16, k=50", "variant": int(variant), "nq": nq, "ms": round(ms, 4),
                              "k1_ms": round(k1, 4), "n_slabs": prof["n_slabs"], "hbm_gbs": round(2.0 * n * dim / (k1 * 1e-3) / 1e9, 1),
                              "hbm_frac": round(2.0 * n * dim / (k1 * 1e-3) / 1e9 / hbm, 3),
                              "tflops": round(2.0 * nq * n * dim / (k1 * 1e-3) / 1e12, 1), "qps": round(nq / ms * 1e3, 1)}), flush=True)
        del idx
    os.environ.pop("AHR_TC_VARIANT", None)
    del X


def run_shard8(dev):
    """The per-rank work of c3 on 8 GPUs: 1.25M x 1024 rows, the full 4096-query batch (strong scaling)."""
    n, dim, nq, k = 1_250_000, 1024, 4096, 50
    _, tf, _ = peaks()
    X, Q = unit_bf16(n, dim, 1234, dev), unit_bf16(nq, dim, 5678, dev)
    idx = engine.VectorIndex(X)
    idx.set_profiling(True)
    ms = timed(lambda: idx.topk(Q, k, mode="bf16"), steps=10, warmup=3)
    prof = idx.get_profile()
    k1 = prof["k1_ms_total"] / max(1, prof["k1_launches"])
    print(json.dumps({"workload": "c3 shard at 8 GPUs: 1.25M x 1024 bf16, Q=4096, k=50", "ms": round(ms, 4), "k1_ms": round(k1, 4),
                      "n_slabs": prof["n_slabs"], "tflops": round(2.0 * nq * n * dim / (k1 * 1e-3) / 1e12, 1),
                      "frac_of_sustained_peak": round(2.0 * nq * n * dim / (k1 * 1e-3) / 1e12 / tf, 3)}), flush=True)
    del idx, X, Q


def run_adv(dev):
    """Adversarial row order: 1.25M x 1024 rows sorted by ascending similarity to the centre of a tight query cluster, so
    for every query almost every row beats everything seen before it (worst case for the epilogue's lists)."""
    n, dim, nq, k = 1_250_000, 1024, 4096, 50
    X = unit_bf16(n, dim, 1234, dev)
    g = torch.Generator(device=dev).manual_seed(77)
    c = torch.randn((dim,), generator=g, device=dev)
    c /= c.norm()
    Q = c[None, :] + 0.02 * torch.randn((nq, dim), generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()
    order = torch.argsort(X.float() @ c)
    Xs = X[order].contiguous()
    del X
    for stride in ("0", "64"):
        os.environ["AHR_TC_SEED_STRIDE"] = stride
        idx = engine.VectorIndex(Xs)
        ms = timed(lambda: idx.topk(Q, k, mode="bf16"), steps=5, warmup=2)
        print(json.dumps({"workload": "adversarial order: 1.25M x 1024 bf16 sorted by similarity to the query cluster, Q=4096, k=50",
                          "seed_stride": int(stride), "ms": round(ms, 3)}), flush=True)
        del idx
    os.environ.pop("AHR_TC_SEED_STRIDE", None)


def make_c5_graph(dev, scale=1.0, seed=42):
    """CSR arrays on the device: [entities | L1 | L2 | hyperedges]; ~25 directed edges per node at scale 1."""
    g = torch.Generator(device=dev).manual_seed(seed)
    n_ent, n_l1, n_l2, n_h = int(16_000_000 * scale), int(1_800_000 * scale), int(200_000 * scale), int(2_000_000 * scale)
    e0, s1, s2, h0 = 0, n_ent, n_ent + n_l1, n_ent + n_l1 + n_l2
    n = h0 + n_h
    rel_deg = 260  # ~468M related_to + ~33M belongs_to/participates_in edges = ~500M edges (1.0e9 CSR entries)

    def csr_from_counts(counts_by_node, cols):
        rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(counts_by_node, 0)
        return rp, cols.to(torch.int32)

    def transpose(src, dst):
        """edges src->dst (link order = src-major) -> CSR keyed by dst, neighbours in link order (stable)."""
        order = torch.sort(dst, stable=True).indices
        counts = torch.bincount(dst, minlength=n)
        return csr_from_counts(counts, src[order])

    zeros = torch.zeros(n, dtype=torch.int64, device=dev)
    # belongs_to: entity -> 1..2 L1 parents ; L1 -> one L2
    cnt_e = torch.randint(1, 3, (n_ent,), generator=g, device=dev)
    src_b = torch.repeat_interleave(torch.arange(n_ent, device=dev), cnt_e)
    dst_b = s1 + torch.randint(0, n_l1, (src_b.numel(),), generator=g, device=dev)
    src_b2 = s1 + torch.arange(n_l1, device=dev)
    dst_b2 = s2 + torch.randint(0, n_l2, (n_l1,), generator=g, device=dev)
    cb = zeros.clone(); cb[:n_ent] = cnt_e; cb[s1:s2] = 1
    bel_out = csr_from_counts(cb, torch.cat([dst_b, dst_b2]))
    bel_in = transpose(torch.cat([src_b, src_b2]), torch.cat([dst_b, dst_b2]))
    # participates_in: hyperedge h has 2..5 entity participants (in-CSR built directly), out-CSR by transpose
    ar = torch.randint(2, 6, (n_h,), generator=g, device=dev)
    hs = h0 + torch.repeat_interleave(torch.arange(n_h, device=dev), ar)
    es = torch.randint(0, n_ent, (hs.numel(),), generator=g, device=dev)
    cp = zeros.clone(); cp[h0:] = ar
    part_in = csr_from_counts(cp, es)
    part_out = transpose(hs, es)  # keyed by entity, values = hyperedges
    # related_to among L1 summaries: rel_deg out-neighbours each
    src_r = s1 + torch.repeat_interleave(torch.arange(n_l1, device=dev), rel_deg)
    dst_r = s1 + torch.randint(0, n_l1, (src_r.numel(),), generator=g, device=dev)
    cr = zeros.clone(); cr[s1:s2] = rel_deg
    rel_out = csr_from_counts(cr, dst_r)
    rel_in = transpose(src_r, dst_r)
    node_type = torch.zeros(n, dtype=torch.uint8, device=dev)
    node_type[s1:h0] = 1
    node_type[h0:] = 2
    level = torch.zeros(n, dtype=torch.uint8, device=dev)
    level[s2:h0] = 2
    judge = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    conf = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    ns = n_l1 + n_l2
    judge[s1:h0] = torch.round(torch.rand(ns, generator=g, device=dev, dtype=torch.float64) * 100) / 10
    conf[s1:h0] = torch.round(torch.rand(ns, generator=g, device=dev, dtype=torch.float64) * 1000) / 100
    jt = torch.zeros(n, dtype=torch.float64, device=dev)
    # judge_term evaluated by the HOST with numpy's scalar exp, as the reference does (hierarchical_graph.py:766)
    jh = judge[s1:h0].cpu().numpy()
    jt[s1:h0] = torch.from_numpy(np.array([1.0 / (1.0 + np.exp(-((float(x) or 0.0) / 10.0))) for x in jh])).to(dev)
    slots = [part_out, part_in, bel_out, bel_in, rel_out, rel_in]
    n_edges = int(bel_out[1].numel() + part_in[1].numel() + rel_out[1].numel())
    return {"n_nodes": n, "row_ptr": [s[0] for s in slots], "col_idx": [s[1] for s in slots], "node_type": node_type,
            "level": level, "judge": judge, "judge_term": jt, "conf": conf, "n_edges": n_edges,
            "ranges": {"entity": (e0, s1), "l1": (s1, s2), "l2": (s2, h0), "hyperedge": (h0, n)}}


class DeviceGraph(engine.GraphCSR):
    """GraphCSR over tensors that already live on the device (no host round trip for the 5 GB CSR)."""

    def __init__(self, a, dev):
        import ctypes
        from ah_rag_b200 import _lib
        self.lib = _lib.load()
        self.device = dev
        self.n_nodes = int(a["n_nodes"])
        self.row_ptr, self.col_idx = a["row_ptr"], a["col_idx"]
        self.node_type, self.level, self.judge, self.judge_term, self.conf = (a["node_type"], a["level"], a["judge"],
                                                                             a["judge_term"], a["conf"])
        rp = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.row_ptr])
        ci = (ctypes.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in self.col_idx])
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.ahr_graph_create(ctypes.byref(self._h), dev.index or 0, self.n_nodes, rp, ci,
                                             self.node_type.data_ptr(), self.level.data_ptr(), self.judge.data_ptr(),
                                             self.judge_term.data_ptr(), self.conf.data_ptr()), "ahr_graph_create")


def run_c5(dev, scale=1.0):
    from oracle import array_graph
    hbm, _, src = peaks()
    t0 = time.time()
    a = make_c5_graph(dev, scale)
    torch.cuda.synchronize()
    g = DeviceGraph(a, dev)
    n = a["n_nodes"]
    n_env, n_seeds, top_k, m, limit = 1024, 50, 50, 5, 10
    gen = torch.Generator(device=dev).manual_seed(7)
    e_lo, e_hi = a["ranges"]["entity"]
    l_lo, l_hi = a["ranges"]["l1"]
    # seeds: 80 % entities, 20 % L1 summaries (the index holds entities + summaries), distinct per env, scores descending
    seeds = torch.where(torch.rand((n_env, n_seeds), generator=gen, device=dev) < 0.8,
                        torch.randint(e_lo, e_hi, (n_env, n_seeds), generator=gen, device=dev),
                        torch.randint(l_lo, l_hi, (n_env, n_seeds), generator=gen, device=dev)).to(torch.int32)
    scores = torch.sort(torch.rand((n_env, n_seeds), generator=gen, device=dev) * 0.9 + 0.05, dim=1, descending=True).values
    scores = scores.to(torch.float32).contiguous()
    prm = engine.SearchParams(member_top_m=m)
    out = g.anchor_rerank(seeds, scores, prm, top_k)
    ms_anchor = timed(lambda: g.anchor_rerank(seeds, scores, prm, top_k, want_cluster=False), steps=20)
    ms_anchor_dev = timed_graph(lambda: g.anchor_rerank(seeds, scores, prm, top_k, want_cluster=False))
    # hop 2: one expand per env on the first ids of the reranked list, ops uniform over {parents, children, related}
    ops = torch.randint(0, 3, (n_env,), generator=gen, device=dev).to(torch.int32)
    take = torch.where(ops == 2, 1, 2)
    sel = out["nodes"][:, :2].contiguous()
    offs = torch.zeros(n_env + 1, dtype=torch.int32, device=dev)
    offs[1:] = torch.cumsum(take, 0).to(torch.int32)
    flat = torch.cat([sel[e, :int(t)] for e, t in enumerate(take.tolist())]).to(torch.int32).contiguous()
    oi, oc = g.expand(ops, flat, offs, limit)
    ms_expand = timed(lambda: g.expand(ops, flat, offs, limit), steps=50)
    ms_expand_dev = timed_graph(lambda: g.expand(ops, flat, offs, limit))
    # uncapped-ish variant: limit 2048 lets the walk read whole neighbour lists (bandwidth view)
    oi2, oc2 = g.expand(ops, flat, offs, 2048)
    ms_expand_big = timed(lambda: g.expand(ops, flat, offs, 2048), steps=20)
    # parity on a sample of environments against the array oracle (host copy of the CSR)
    host = {"row_ptr": [t.cpu().numpy() for t in a["row_ptr"]], "col_idx": [t.cpu().numpy() for t in a["col_idx"]],
            "node_type": a["node_type"].cpu().numpy(), "judge": a["judge"].cpu().numpy(),
            "judge_term": a["judge_term"].cpu().numpy(), "conf": a["conf"].cpu().numpy()}
    sample = list(range(0, n_env, 16))
    outh = {k_: v.cpu().numpy() for k_, v in out.items()}
    sh, sch = seeds.cpu().numpy(), scores.cpu().numpy()
    ok_anchor = all(array_graph.anchor_rerank(host, sh[e], sch[e], top_k, m) ==
                    [(int(outh["nodes"][e, j]), float(outh["semantic"][e, j]), float(outh["score"][e, j]))
                     for j in range(outh["count"][e])] for e in sample)
    oih, och, oi2h, oc2h = oi.cpu().numpy(), oc.cpu().numpy(), oi2.cpu().numpy(), oc2.cpu().numpy()
    offh, flath, opsh = offs.cpu().numpy(), flat.cpu().numpy(), ops.cpu().numpy()
    ok_expand = all(array_graph.expand(host, int(opsh[e]), flath[offh[e]:offh[e + 1]], limit) == list(oih[e, :och[e]]) and
                    array_graph.expand(host, int(opsh[e]), flath[offh[e]:offh[e + 1]], 2048) == list(oi2h[e, :oc2h[e]])
                    for e in sample)
    emitted = int(oc2.sum().item())
    # bandwidth view of the hop kernel: 65536 frontier nodes (L1 summaries, 260 out + ~260 in related_to neighbours each),
    # limit 512 so the walk reads both neighbour lists in full
    nb = 65536
    big_in = torch.randint(l_lo, l_hi, (nb,), generator=gen, device=dev).to(torch.int32)
    big_ops = torch.full((nb,), 2, dtype=torch.int32, device=dev)
    big_off = torch.arange(nb + 1, dtype=torch.int32, device=dev)
    g.expand(big_ops, big_in, big_off, 512)
    ms_bw = timed(lambda: g.expand(big_ops, big_in, big_off, 512), steps=10)
    rp_o, rp_i = a["row_ptr"][4], a["row_ptr"][5]
    bi = big_in.long()
    deg = (rp_o[bi + 1] - rp_o[bi]) + (rp_i[bi + 1] - rp_i[bi])
    bw_bytes = int(deg.sum().item()) * 4 + nb * (2 * 16 + 4) + nb * 512 * 4  # col_idx + row_ptr pairs + ids in + ids out
    # device-resident gym rollout: 1024 environments, reset from a top-5 anchor, 6 random steps (one launch per step)
    envb = engine.EnvBatchState(g, n_env, max_steps=6)
    rr5 = g.anchor_rerank(seeds[:, :5].contiguous(), scores[:, :5].contiguous(), prm, 5, want_cluster=False)
    acts = [torch.randint(0, 5, (n_env,), generator=gen, device=dev).to(torch.int32) for _ in range(6)]

    def rollout():
        envb.reset(rr5, 5)
        for t in range(6):
            envb.step(acts[t])

    ms_roll = timed(rollout, steps=20)
    envb.check_overflow()
    print(json.dumps({"workload": f"c5: synthetic hierarchical graph {n} nodes / {a['n_edges']} edges (scale {scale}), 1024 envs x 50 seeds",
                      "graph_bytes": g.nbytes(), "build_s": round(time.time() - t0, 1),
                      "hop1_anchor_rerank_ms": ms_anchor, "hop1_envs_per_s": n_env / ms_anchor * 1e3,
                      "hop1_device_ms_cuda_graph": ms_anchor_dev, "hop2_device_ms_cuda_graph": ms_expand_dev,
                      "hop2_expand_limit10_ms": ms_expand, "hop2_envs_per_s": n_env / ms_expand * 1e3,
                      "hop2_expand_limit2048_ms": ms_expand_big, "hop2_nodes_emitted": emitted,
                      "hop_bw_65536_nodes_ms": ms_bw, "hop_bw_algorithmic_bytes": bw_bytes,
                      "hop_bw_gbs": bw_bytes / (ms_bw * 1e-3) / 1e9, "hop_bw_frac_of_hbm": bw_bytes / (ms_bw * 1e-3) / 1e9 / hbm,
                      "gym_rollout_reset_plus_6_steps_ms": ms_roll, "gym_env_steps_per_s": n_env * 6 / ms_roll * 1e3,
                      "parity_sample_envs": len(sample), "anchor_rerank_bit_exact": bool(ok_anchor),
                      "expand_bit_exact": bool(ok_expand), "hbm_peak_gbs": hbm, "peak_source": src}))


def run_c2(dev):
    """HotpotQA-scale substitute: 500k nodes, D=384 fp32 index (entities + summaries), 7405 queries, top_k=5."""
    a = make_c5_graph(dev, scale=0.025, seed=3)  # 400k entities, 45k L1, 5k L2, 50k hyperedges
    g = DeviceGraph(a, dev)
    e_lo, e_hi = a["ranges"]["entity"]
    s_lo, s_hi = a["ranges"]["l1"][0], a["ranges"]["l2"][1]
    n_index = s_hi  # entities + all summaries are indexed (hierarchical_graph.py:629-649), rows in node order
    dim, nq, k = 384, 7405, 5
    gen = torch.Generator(device=dev).manual_seed(11)
    X = torch.randn((n_index, dim), generator=gen, device=dev)
    X /= X.norm(dim=1, keepdim=True)
    Q = torch.randn((nq, dim), generator=gen, device=dev)
    Q /= Q.norm(dim=1, keepdim=True)
    idx = engine.VectorIndex(X.contiguous())
    prm = engine.SearchParams(member_top_m=5)

    def step():
        s, i = idx.topk(Q, k, mode="exact")
        return g.anchor_rerank(i.to(torch.int32), s, prm, k, want_cluster=False)

    ms = timed(step, steps=3, warmup=1)
    Xb, Qb = X.to(torch.bfloat16).contiguous(), Q.to(torch.bfloat16).contiguous()
    idxb = engine.VectorIndex(Xb)

    def step_b():
        s, i = idxb.topk(Qb, k, mode="bf16")
        return g.anchor_rerank(i.to(torch.int32), s, prm, k, want_cluster=False)

    msb = timed(step_b, steps=10, warmup=3)
    print(json.dumps({"workload": f"c2 substitute: synthetic graph {a['n_nodes']} nodes, index {n_index} x {dim}, {nq} queries, top_k={k} "
                                  "(the shipped hotpotqa_distractor graph is empty)",
                      "fp32_exact_ms": ms, "fp32_exact_queries_per_s": nq / ms * 1e3,
                      "bf16_ms": msb, "bf16_queries_per_s": nq / msb * 1e3}))


def _dist_setup():
    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, dev


def _timed_dist(dist, world, dev, fn, steps, warmup=3, after_warmup=None):
    """bench.py's timing: barrier + synchronize on both sides, CUDA events, max over ranks."""
    for _ in range(warmup):
        fn()
    if after_warmup is not None:
        after_warmup()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def run_c4dist():
    """configs[3]: 100M x 768 bf16 sharded by node range over the ranks of `torchrun`, 256-query batch, k = 50
    (strong scaling on the fixed corpus; run with 2, 4 or 8 ranks -- one GPU cannot hold the 153.6 GB)."""
    from ah_rag_b200.dist import ShardedIndex, shard_bounds
    dist, world, rank, dev = _dist_setup()
    n_total, dim, nq, k = 100_000_000, 768, 256, 50
    hbm, tf, src = peaks()
    lo, hi = shard_bounds(n_total, world, rank)
    assert (hi - lo) * dim * 2 < 150e9, "shard does not fit one GPU: use more ranks"
    X = unit_bf16(hi - lo, dim, 1234 + 1000 * rank, dev)
    Q = unit_bf16(nq, dim, 5678, dev)
    sh = ShardedIndex.from_rows(X, id_base=lo)
    peer = sh.enable_peer_exchange(nq, k) if world > 1 else False
    sh.index.set_profiling(True)
    ms = _timed_dist(dist, world, dev, lambda: sh.topk(Q, k, mode="bf16"), steps=20)
    prof = sh.index.get_profile()
    k1 = prof["k1_ms_total"] / max(1, prof["k1_launches"])
    t_mem = 2.0 * (hi - lo) * dim / (hbm * 1e9) * 1e3
    t_mma = 2.0 * nq * (hi - lo) * dim / (tf * 1e12) * 1e3
    if rank == 0:
        print(json.dumps({"workload": f"c4: 100M x 768 bf16 over {world} GPUs (shards of {hi - lo} rows), Q=256, k=50, bf16 mode",
                          "n_gpus": world, "ms_per_batch": round(ms, 4), "queries_per_s": round(nq / ms * 1e3, 1),
                          "k1_ms_rank0": round(k1, 4), "exchange": "peer-memory stores" if peer else ("nccl all-gather" if world > 1 else "none"),
                          "roofline_rank0": {"t_mem_ms": round(t_mem, 3), "t_mma_ms": round(t_mma, 3),
                                             "binding": "tensor" if t_mma > t_mem else "hbm",
                                             "frac": round(max(t_mem, t_mma) / k1, 3),
                                             "hbm_gbs": round(2.0 * (hi - lo) * dim / (k1 * 1e-3) / 1e9, 1), "peak_source": src}}), flush=True)
    del sh, X


def run_c5dist():
    """configs[4] at N GPUs: the graph is replicated on every GPU, the 1024 environments are sharded across the ranks
    (no exchange on the hop path); per batch: anchor expansion + rerank (hop 1, 50 seeds) and one frontier expansion
    (hop 2) per environment.  Value = environments per second over all ranks."""
    dist, world, rank, dev = _dist_setup()
    a = make_c5_graph(dev, float(os.environ.get("AHR_C5_SCALE", "1.0")))
    g = DeviceGraph(a, dev)
    n_env_total, n_seeds, top_k, m = 1024, 50, 50, 5
    n_env = n_env_total // world
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    e_lo, e_hi = a["ranges"]["entity"]
    l_lo, l_hi = a["ranges"]["l1"]
    seeds = torch.where(torch.rand((n_env, n_seeds), generator=gen, device=dev) < 0.8,
                        torch.randint(e_lo, e_hi, (n_env, n_seeds), generator=gen, device=dev),
                        torch.randint(l_lo, l_hi, (n_env, n_seeds), generator=gen, device=dev)).to(torch.int32)
    scores = torch.sort(torch.rand((n_env, n_seeds), generator=gen, device=dev) * 0.9 + 0.05, dim=1, descending=True).values
    scores = scores.to(torch.float32).contiguous()
    prm = engine.SearchParams(member_top_m=m)
    ops = torch.randint(0, 3, (n_env,), generator=gen, device=dev).to(torch.int32)
    offs = (torch.arange(n_env + 1, device=dev) * 2).to(torch.int32)

    def step():
        out = g.anchor_rerank(seeds, scores, prm, top_k, want_cluster=False)
        return g.expand(ops, out["nodes"][:, :2].contiguous().view(-1), offs, 10)

    ms = _timed_dist(dist, world, dev, step, steps=50)
    if rank == 0:
        print(json.dumps({"workload": f"c5: 20M nodes / {a['n_edges']} edges replicated on {world} GPU(s), {n_env_total} environments "
                                      f"sharded {n_env} per GPU, hop 1 (50 seeds, rerank) + hop 2 (expand, limit 10)",
                          "n_gpus": world, "ms_per_batch": round(ms, 4), "envs_per_s": round(n_env_total / ms * 1e3, 1),
                          "scaling": "environments sharded, graph replicated, no collective"}), flush=True)


# ---------------------------------------------------------------------------------------------
# The "extra" block of bench.py: driver-timed figures for configs[3] (c4) and configs[4] (c5), with the graph-half
# CPU baseline.  Called by bench.py after its c3 loops, in the same process / process group.
# ---------------------------------------------------------------------------------------------
def extra_c4(dist, world, rank, dev, steps=10):
    """c4 = 100M x 768 bf16, 256-query batch, k = 50 (the HBM / balance-point regime of K1).
    world > 1: the real configuration -- 100M rows sharded by node range over the ranks, candidate exchange + merge.
    world = 1: one GPU cannot hold the 153.6 GB, so the per-rank work of the 8-GPU configuration (12.5M rows) is timed,
    at Q = 256 (CTA pairs) and Q = 64 (one CTA per tile: the pure HBM stream)."""
    from ah_rag_b200.dist import ShardedIndex, shard_bounds
    n_total, dim, k = 100_000_000, 768, 50
    hbm, tf, src = peaks()
    lo, hi = shard_bounds(n_total, max(world, 8) if world == 1 else world, rank)
    n = hi - lo
    X = unit_bf16(n, dim, 1234 + 1000 * rank, dev)
    Qall = unit_bf16(256, dim, 5678, dev)
    sh = ShardedIndex.from_rows(X, id_base=lo)
    peer = sh.enable_peer_exchange(256, k) if world > 1 else False
    out = {"workload": (f"c4: 100M x 768 bf16 over {world} GPUs (shards of {n} rows), k=50, bf16 mode" if world > 1 else
                        f"c4 shard: {n} x 768 bf16 = the per-GPU work of c4 on 8 GPUs (one GPU cannot hold 153.6 GB), k=50, bf16 mode"),
           "n_gpus": world, "exchange": "peer-memory stores" if peer else ("nccl all-gather" if world > 1 else "none"),
           "hbm_peak_gbs": hbm, "tensor_peak_tflops": tf, "peak_source": src, "batches": []}
    for nq in (256, 64):
        Q = Qall[:nq].contiguous()
        sh.index.set_profiling(True)
        # per-launch K1 events are kept for the timed steps only: a kernel variant's first launch includes its lazy module
        # load on the host between the two event records (seen once as 11 ms "per launch" at Q = 64)
        ms = _timed_dist(dist, world, dev, lambda: sh.topk(Q, k, mode="bf16"), steps=steps, after_warmup=sh.index.get_profile)
        prof = sh.index.get_profile()
        sh.index.set_profiling(False)
        k1 = prof["k1_ms_total"] / max(1, prof["k1_launches"])
        k1_all = [k1]
        if world > 1:  # the step ends when the slowest rank's list arrives: every rank's K1 time goes into the line
            t_k1 = torch.tensor([k1], dtype=torch.float64, device=dev)
            t_all = [torch.zeros_like(t_k1) for _ in range(world)]
            dist.all_gather(t_all, t_k1)
            k1_all = [float(x.item()) for x in t_all]
        t_mem = 2.0 * n * dim / (hbm * 1e9) * 1e3
        t_mma = 2.0 * max(nq, 128) * n * dim / (tf * 1e12) * 1e3  # the tensor core works on whole 128-row query tiles
        out["batches"].append({"nq": nq, "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3, "k1_ms_rank0": k1,
                               "k1_ms_per_rank": k1_all, "slowest_rank_k1_share_of_batch": max(k1_all) / ms,
                               "hbm_gbs": 2.0 * n * dim / (k1 * 1e-3) / 1e9, "hbm_frac": t_mem / k1,
                               "t_mem_ms": t_mem, "t_mma_ms": t_mma, "binding": "tensor" if t_mma > t_mem else "hbm",
                               "frac_of_binding_roofline": max(t_mem, t_mma) / k1, "n_slabs": prof["n_slabs"]})
    del sh, X
    torch.cuda.empty_cache()
    return out


def extra_c1(dist, world, rank, dev):
    """configs[0]: the samples/hotpotqa_mini.txt graph (413 nodes; fixture tests/golden/g_c1_hotpotqa_mini.*, every expected
    value written by the unmodified reference) on the test_graph_search path, THROUGH THE DROP-IN Python API: the 64 queries
    of the fixture as one `search_batch` (B = 64) + the 2-hop step (`expand_related` on each query's best node), and
    `search("Tim Burton", top_k=3)` alone (B = 1) -- host in, Python dicts out.  Beside it the same calls through the oracle's
    restatement of the reference (NetworkX walks + per-query loop, one host core).  One replica per GPU; rank 0 reports."""
    import copy
    import shutil
    import tempfile
    from ah_rag_b200 import synth
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
    from tests import golden_util as gu
    fx = gu.load("g_c1_hotpotqa_mini")
    tmp = tempfile.mkdtemp(prefix=f"ahr_c1_r{rank}_")
    try:
        gdir, db = os.path.join(tmp, "graph"), os.path.join(tmp, "vector_db")
        nl = copy.deepcopy(fx["node_link"])
        nl["graph"]["vector_index"]["db_path"] = db
        synth.write_graph_dir(gdir, nl)
        hg = HierarchicalGraph.load(gdir)
        hg.device = str(dev)
        qmap = {q_: fx["query_vecs"][i] for i, q_ in enumerate(fx["queries"])}
        hg.set_embedder(lambda texts: np.stack([qmap[t] for t in texts]))
        hg.build_vector_index(db, layers={0, 1, 2}, reset=True, embeddings={nid: fx["index_rows"][r] for r, nid in enumerate(fx["index_ids"])})
        dg = hg._device_graph()
        dense, names, csr = dg["arrays"]["dense"], dg["arrays"]["node_ids"], dg["csr"]
        nq = len(fx["queries"])
        ops = torch.full((nq,), 2, dtype=torch.int32, device=dev)
        offs = torch.arange(nq + 1, dtype=torch.int32, device=dev)

        def two_hops():
            outs = hg.search_batch(fx["query_vecs"], top_k=5)
            first = torch.tensor([dense[o[0]["node_id"]] if o else -1 for o in outs], dtype=torch.int32, device=dev)
            oi, oc = csr.expand(ops, first, offs, 10)
            oi, oc = oi.tolist(), oc.tolist()
            return outs, [[names[v] for v in oi[b][:oc[b]]] for b in range(nq)]

        def wall(fn, reps):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        outs, hop2 = two_hops()
        t_b64 = wall(two_hops, 20)
        t_b1 = wall(lambda: hg.search("Tim Burton", top_k=3), 50)
        res = {"workload": f"c1: samples/hotpotqa_mini.txt graph ({len(names)} nodes, {len(fx['index_ids'])} index rows x {fx['index_rows'].shape[1]} fp32), "
                           f"{nq} queries, top_k=5 + expand_related of the best node, through HierarchicalGraph.search_batch (fp32-exact mode)",
               "n_gpus": world, "b64_two_hops_ms": t_b64 * 1e3, "b64_queries_per_s": nq / t_b64,
               "b1_search_tim_burton_top3_ms": t_b1 * 1e3, "timing": "wall clock around the Python call (host vectors in, Python dicts out)"}
        if rank == 0:
            from oracle import ref_search  # checker + reported CPU baseline, never the measured path
            want = {c["query"]: c["out"]["reranked"] for c in fx["search"] if c["kwargs"] == {"top_k": 5}}
            ok1 = all(gu.canon(outs[q_]) == gu.canon(w) for q_, w in want.items())
            og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
            env = ref_search.OracleEnv(og, fx["index_ids"], fx["index_rows"])

            def cpu_pass():
                out = []
                for b in range(nq):
                    rr = env.search(fx["query_vecs"][b], 5)["reranked"]
                    out.append(ref_search.expand(og, "related", [rr[0]["node_id"]] if rr else [], 10))
                return out

            cpu_hop2 = cpu_pass()
            t0 = time.perf_counter()
            for _ in range(3):
                cpu_pass()
            t_cpu = (time.perf_counter() - t0) / 3
            res["parity_check"] = {"search_equals_reference_goldens": bool(ok1), "golden_cases": len(want),
                                   "expand_related_equals_oracle": bool(cpu_hop2 == hop2)}
            res["cpu_baseline"] = {"value": nq / t_cpu, "unit": "queries/s (search top_k=5 + expand_related)", "cores": 1, "kind": "port",
                                   "sample": f"oracle/ref_search.py (the reference's algorithm over NetworkX, brute-force cosine in numpy) "
                                             f"sequential loop over the {nq} queries, {t_cpu * 1e3:.1f} ms per pass"}
        del hg
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def extra_c2(dist, world, rank, dev):
    """configs[1] substitute (the shipped hotpotqa_distractor graph is empty): a HotpotQA-scale synthetic graph
    (500k nodes), a 450k x 384 fp32 index (MiniLM-shaped rows, entities + summaries), all 7405 dev questions in one batch,
    top_k = 5: NaiveRAG's retrieval (fp32-exact top-k) + HierarchicalGraph.search's anchor expansion and rerank.  One GPU
    per replica ("1 B200" in the config): every rank runs it, rank 0 checks a sample against the oracle on the host."""
    a = make_c5_graph(dev, scale=0.025, seed=3)  # 400k entities, 45k L1, 5k L2, 50k hyperedges
    g = DeviceGraph(a, dev)
    n_index = a["ranges"]["l2"][1]  # entities + all summaries are indexed (hierarchical_graph.py:629-649), rows in node order
    dim, nq, k = 384, 7405, 5
    gen = torch.Generator(device=dev).manual_seed(11)
    X = torch.randn((n_index, dim), generator=gen, device=dev)
    X /= X.norm(dim=1, keepdim=True)
    Q = torch.randn((nq, dim), generator=gen, device=dev)
    Q /= Q.norm(dim=1, keepdim=True)
    X, Q = X.contiguous(), Q.contiguous()
    idx = engine.VectorIndex(X)
    prm = engine.SearchParams(member_top_m=5)

    def step():
        s, i = idx.topk(Q, k, mode="exact")
        return s, i, g.anchor_rerank(i.to(torch.int32), s, prm, k, want_cluster=False)

    s, i, out = step()
    ms = _timed_dist(dist, world, dev, step, steps=5, warmup=2)
    res = {"workload": f"c2 substitute: synthetic graph {a['n_nodes']} nodes / {a['n_edges']} edges, index {n_index} x {dim} fp32, "
                       f"{nq} queries in one batch, top_k={k}, fp32-exact top-k + anchor expansion + rerank; one replica per GPU",
           "n_gpus": world, "ms_per_batch": ms, "queries_per_s_per_gpu": nq / ms * 1e3, "queries_per_s": world * nq / ms * 1e3}
    if rank == 0:
        from oracle import array_graph, exact  # the checker, never the measured path
        sample = [0, 1, nq // 2, nq - 1]
        rs, ri = exact.exact_topk_rowpar(X.cpu().numpy(), Q[sample].cpu().numpy(), k)
        ih, sh = i.cpu().numpy(), s.cpu().numpy()
        ok_ids = bool(np.array_equal(ih[sample], ri)) and bool(np.array_equal(sh[sample].view(np.uint32), rs.view(np.uint32)))
        host = {"row_ptr": [t.cpu().numpy() for t in a["row_ptr"]], "col_idx": [t.cpu().numpy() for t in a["col_idx"]],
                "node_type": a["node_type"].cpu().numpy(), "judge": a["judge"].cpu().numpy(),
                "judge_term": a["judge_term"].cpu().numpy(), "conf": a["conf"].cpu().numpy()}
        o = {k_: v.cpu().numpy() for k_, v in out.items()}
        ok_rr = all(array_graph.anchor_rerank(host, ri[j].astype(np.int32), rs[j], k, 5) ==
                    [(int(o["nodes"][q, t]), float(o["semantic"][q, t]), float(o["score"][q, t])) for t in range(o["count"][q])]
                    for j, q in enumerate(sample))
        res["parity_check"] = {"queries": sample, "exact_ids_and_score_bits_equal": ok_ids, "anchor_rerank_bit_exact": bool(ok_rr),
                               "oracle": "oracle/exact_topk.c full scan + oracle/array_graph.py on the host"}
    del idx, g, a, X, Q
    torch.cuda.empty_cache()
    return res


def extra_c5(dist, world, rank, dev, scale=1.0, cpu_budget_s=6.0):
    """c5 = synthetic hierarchical graph 20M nodes / 500M edges, 1024 environments x 50 seeds: hop 1 (anchor expansion +
    rerank), hop 2 (one expand_* per environment), the uncapped-walk bandwidth view, the device gym rollout; graph
    replicated per GPU, environments sharded over the ranks (no exchange on the hop path).  Rank 0 adds the graph-half CPU
    baseline (sequential loop, one core): oracle/array_graph.py at this graph's full size, and oracle/ref_search.py over
    NetworkX on a reduced graph."""
    hbm, _, src = peaks()
    a = make_c5_graph(dev, scale)
    g = DeviceGraph(a, dev)
    n_env_total, n_seeds, top_k, m, limit = 1024, 50, 50, 5, 10
    n_env = n_env_total // world
    seeds, scores, ops = c5_inputs(a, n_env_total, n_seeds, dev)
    sl = slice(rank * n_env, (rank + 1) * n_env)
    seeds_l, scores_l, ops_l = seeds[sl].contiguous(), scores[sl].contiguous(), ops[sl].contiguous()
    prm = engine.SearchParams(member_top_m=m)
    out1 = g.anchor_rerank(seeds_l, scores_l, prm, top_k)
    ms_h1 = _timed_dist(dist, world, dev, lambda: g.anchor_rerank(seeds_l, scores_l, prm, top_k, want_cluster=False), steps=20)
    ms_h1_dev = timed_graph(lambda: g.anchor_rerank(seeds_l, scores_l, prm, top_k, want_cluster=False))
    take = torch.where(ops_l == 2, 1, 2)
    offs = torch.zeros(n_env + 1, dtype=torch.int32, device=dev)
    offs[1:] = torch.cumsum(take, 0).to(torch.int32)
    first2 = out1["nodes"][:, :2]
    keep = torch.arange(2, device=dev)[None, :] < take[:, None]
    flat = first2[keep].to(torch.int32).contiguous()
    oi, oc = g.expand(ops_l, flat, offs, limit)
    ms_h2 = _timed_dist(dist, world, dev, lambda: g.expand(ops_l, flat, offs, limit), steps=50)
    ms_h2_dev = timed_graph(lambda: g.expand(ops_l, flat, offs, limit))
    # both hops back to back (the per-batch cost of a rollout step that re-anchors)
    ms_both = _timed_dist(dist, world, dev, lambda: (g.anchor_rerank(seeds_l, scores_l, prm, top_k, want_cluster=False),
                                                     g.expand(ops_l, flat, offs, limit)), steps=20)
    # uncapped walk = bandwidth view: 65536 L1 summaries (260 out + ~260 in related_to neighbours each), limit 512
    l_lo, l_hi = a["ranges"]["l1"]
    gen = torch.Generator(device=dev).manual_seed(70 + rank)
    nb = 65536 // world
    big_in = torch.randint(l_lo, l_hi, (nb,), generator=gen, device=dev).to(torch.int32)
    big_ops = torch.full((nb,), 2, dtype=torch.int32, device=dev)
    big_off = torch.arange(nb + 1, dtype=torch.int32, device=dev)
    g.expand(big_ops, big_in, big_off, 512)
    ms_bw = _timed_dist(dist, world, dev, lambda: g.expand(big_ops, big_in, big_off, 512), steps=10)
    rp_o, rp_i = a["row_ptr"][4], a["row_ptr"][5]
    bi = big_in.long()
    deg = (rp_o[bi + 1] - rp_o[bi]) + (rp_i[bi + 1] - rp_i[bi])
    bw_bytes = int(deg.sum().item()) * 4 + nb * (2 * 16 + 4) + nb * 512 * 4  # col_idx + row_ptr pairs + ids in + ids out
    # hop 1 at a batch that fills the machine (bandwidth view of the anchor kernel): 65536 queries x 50 seeds
    nb1 = 65536 // world
    seeds_b, scores_b, _ = c5_inputs(a, nb1, n_seeds, dev, seed=17 + rank)
    outb = g.anchor_rerank(seeds_b, scores_b, prm, top_k, want_cluster=False)
    ms_h1_bw = _timed_dist(dist, world, dev, lambda: g.anchor_rerank(seeds_b, scores_b, prm, top_k, want_cluster=False), steps=10)
    # algorithmic bytes: seed id + score in, one 64-byte line per seed, one 32-byte head per first occurrence, results out
    h1_bytes = nb1 * n_seeds * (4 + 4 + 64) + int(outb["count"].sum().item()) * 0 + nb1 * top_k * (4 + 8 + 8) + nb1 * 4
    h1_heads = int((g.anchor_rerank(seeds_b[:4096], scores_b[:4096], prm, top_k)["exp_count"]).sum().item()) * (nb1 // 4096)
    h1_bytes += h1_heads * 32
    del outb
    # device gym rollout: reset from a top-5 anchor + 6 steps
    envb = engine.EnvBatchState(g, n_env, max_steps=6)
    rr5 = g.anchor_rerank(seeds_l[:, :5].contiguous(), scores_l[:, :5].contiguous(), prm, 5, want_cluster=False)
    acts = [torch.randint(0, 5, (n_env,), generator=gen, device=dev).to(torch.int32) for _ in range(6)]

    def rollout():
        envb.reset(rr5, 5)
        for t in range(6):
            envb.step(acts[t])

    ms_roll = _timed_dist(dist, world, dev, rollout, steps=20)
    replay, _ = envb.capture_rollout(rr5, 5, torch.stack(acts))
    ms_roll_graph = _timed_dist(dist, world, dev, replay, steps=20)
    envb.check_overflow()
    out = {"workload": f"c5: synthetic hierarchical graph {a['n_nodes']} nodes / {a['n_edges']} edges, {n_env_total} environments x "
                       f"{n_seeds} seeds; graph replicated on {world} GPU(s), {n_env} environments per GPU, no collective",
           "n_gpus": world, "graph_bytes": g.nbytes(),
           "hop1_anchor_rerank_ms": ms_h1, "hop2_expand_limit10_ms": ms_h2, "two_hops_ms": ms_both,
           "hop1_device_ms_cuda_graph": ms_h1_dev, "hop2_device_ms_cuda_graph": ms_h2_dev,
           "timing_note": "the *_ms figures are per call through Python/ctypes (launch-bound at ~30 us per call); "
                          "*_device_ms_cuda_graph replay the same calls from a CUDA graph = device time",
           "envs_per_s_two_hops": n_env_total / ms_both * 1e3,
           "uncapped_walk": {"frontier_nodes_per_gpu": nb, "ms": ms_bw, "algorithmic_bytes_per_gpu": bw_bytes,
                             "useful_gbs_per_gpu": bw_bytes / (ms_bw * 1e-3) / 1e9,
                             "frac_of_hbm": bw_bytes / (ms_bw * 1e-3) / 1e9 / hbm},
           "hop1_bandwidth_view": {"queries_per_gpu": nb1, "seeds": n_seeds, "ms": ms_h1_bw, "seeds_per_s_per_gpu": nb1 * n_seeds / ms_h1_bw * 1e3,
                                   "algorithmic_bytes_per_gpu": h1_bytes, "useful_gbs_per_gpu": h1_bytes / (ms_h1_bw * 1e-3) / 1e9,
                                   "frac_of_hbm": h1_bytes / (ms_h1_bw * 1e-3) / 1e9 / hbm,
                                   "note": "random 64-byte lines + 32-byte heads: the binding roofline is the random-sector rate of HBM, not the copy bandwidth"},
           "gym_rollout_reset_plus_6_steps_ms": ms_roll, "gym_env_steps_per_s": n_env_total * 6 / ms_roll * 1e3,
           "gym_rollout_cuda_graph_ms": ms_roll_graph, "gym_env_steps_per_s_cuda_graph": n_env_total * 6 / ms_roll_graph * 1e3,
           "hbm_peak_gbs": hbm, "peak_source": src}
    if rank == 0:
        from oracle import array_graph, cpu_graph_baseline  # checker + reported CPU baseline, never the measured path
        host = {"row_ptr": [t.cpu().numpy() for t in a["row_ptr"]], "col_idx": [t.cpu().numpy() for t in a["col_idx"]],
                "node_type": a["node_type"].cpu().numpy(), "level": a["level"].cpu().numpy(), "judge": a["judge"].cpu().numpy(),
                "judge_term": a["judge_term"].cpu().numpy(), "conf": a["conf"].cpu().numpy()}
        sh, sch, opsh = seeds.cpu().numpy(), scores.cpu().numpy(), ops.cpu().numpy()
        # parity on a sample of this rank's environments (array oracle over the host copy of the CSR)
        o1 = {k_: v.cpu().numpy() for k_, v in out1.items()}
        oih, och, offh, flath = oi.cpu().numpy(), oc.cpu().numpy(), offs.cpu().numpy(), flat.cpu().numpy()
        sample = list(range(0, n_env, max(1, n_env // 32)))
        ok1 = all(array_graph.anchor_rerank(host, sh[e], sch[e], top_k, m) ==
                  [(int(o1["nodes"][e, j]), float(o1["semantic"][e, j]), float(o1["score"][e, j])) for j in range(o1["count"][e])]
                  for e in sample)
        ok2 = all(array_graph.expand(host, int(opsh[e]), flath[offh[e]:offh[e + 1]], limit) == list(oih[e, :och[e]]) for e in sample)
        out["parity_check"] = {"sample_envs": len(sample), "anchor_rerank_bit_exact": bool(ok1), "expand_bit_exact": bool(ok2),
                               "oracle": "oracle/array_graph.py over the host copy of the CSR"}
        arr = cpu_graph_baseline.array_loop(host, sh, sch, opsh, top_k, m, limit, cpu_budget_s)
        out["cpu_baseline"] = {"value": arr["envs_per_s"], "unit": "environments/s (hop 1 + hop 2)", "cores": 1, "kind": "port",
                               "sample": f"oracle/array_graph.py (numpy CSR restatement) sequential loop over {arr['envs']} of the "
                                         f"{n_env_total} environments on the full graph, {arr['seconds']:.1f} s"}
        del host
        small = make_c5_graph(dev, 0.0025 * scale, seed=43)  # ~50k nodes / ~1.25M edges: what NetworkX builds in seconds
        hs = {"row_ptr": [t.cpu().numpy() for t in small["row_ptr"]], "col_idx": [t.cpu().numpy() for t in small["col_idx"]],
              "node_type": small["node_type"].cpu().numpy(), "level": small["level"].cpu().numpy(),
              "judge": small["judge"].cpu().numpy(), "judge_term": small["judge_term"].cpu().numpy(), "conf": small["conf"].cpu().numpy()}
        s2, sc2, op2 = c5_inputs(small, 256, n_seeds, dev)
        nxr = cpu_graph_baseline.networkx_loop(hs, s2.cpu().numpy(), sc2.cpu().numpy(), op2.cpu().numpy(), top_k, m, limit, cpu_budget_s)
        out["cpu_baseline_networkx"] = {"value": nxr["envs_per_s"], "unit": "environments/s (hop 1 + hop 2)", "cores": 1, "kind": "port",
                                        "sample": f"oracle/ref_search.py (the reference's algorithm over NetworkX, as the reference holds its "
                                                  f"graph) sequential loop, {nxr['envs']} environments on a reduced graph of {nxr['nodes']} nodes / "
                                                  f"{nxr['edges']} edges, {nxr['seconds']:.1f} s (+ {nxr['graph_build_s']:.1f} s to build the DiGraph)"}
    if world > 1:
        out["sharded_csr"] = c5_sharded(dist, world, rank, dev, a, g, seeds, scores, ops, prm, top_k, limit)
    del g, a
    torch.cuda.empty_cache()
    return out


def c5_sharded(dist, world, rank, dev, a, g, seeds, scores, ops, prm, top_k, limit):
    """SURVEY 8e option B on c5: the CSR (and the node attributes) sharded by node range over the ranks, the batch of 1024
    environments replicated; hop 1 = owner-routed records + the rerank kernel (two NCCL all-gathers), hop 2 = staged expansion
    (two all-gathers).  Every rank ends with the whole batch's answer; checked here against the replicated graph `g`."""
    from ah_rag_b200.dist import ShardedGraph
    sg = ShardedGraph.from_arrays(a, dev)
    n_env = int(seeds.shape[0])
    want1 = g.anchor_rerank(seeds, scores, prm, top_k)
    got1 = sg.anchor_rerank(seeds, scores, prm, top_k)
    cnt = want1["count"]
    live = torch.arange(top_k, device=dev)[None, :] < cnt[:, None]
    ok1 = bool(torch.equal(got1["count"], cnt) and torch.equal(got1["nodes"][live], want1["nodes"][live]) and
               torch.equal(got1["score"][live].view(torch.int64), want1["score"][live].view(torch.int64)) and
               torch.equal(got1["semantic"][live].view(torch.int64), want1["semantic"][live].view(torch.int64)))
    take = torch.where(ops == 2, 1, 2)
    offs = torch.zeros(n_env + 1, dtype=torch.int32, device=dev)
    offs[1:] = torch.cumsum(take, 0).to(torch.int32)
    keep = torch.arange(2, device=dev)[None, :] < take[:, None]
    flat = want1["nodes"][:, :2][keep].to(torch.int32).contiguous()
    wi, wc = g.expand(ops, flat, offs, limit)
    oi, oc = sg.expand(ops, flat, offs, limit=limit, max_ids=2)
    ok2 = bool(torch.equal(oi, wi) and torch.equal(oc, wc))
    ms1 = _timed_dist(dist, world, dev, lambda: sg.anchor_rerank(seeds, scores, prm, top_k, want_cluster=False), steps=20)
    ms2 = _timed_dist(dist, world, dev, lambda: sg.expand(ops, flat, offs, limit=limit, max_ids=2), steps=20)
    cp = prm.to_c()
    rec = [sg.shard.anchor_record_ints(1, int(seeds.shape[1]), cp), sg.shard.anchor_record_ints(2, int(seeds.shape[1]), cp),
           sg.shard.record_ints(1, limit, 2), sg.shard.record_ints(2, limit, 2)]
    res = {"what": f"CSR + node attributes sharded by node range over {world} GPUs ({sg.shard.nbytes()} bytes on rank 0 of "
                   f"{g.nbytes()} replicated), {n_env} environments replicated, results on every rank",
           "hop1_anchor_rerank_ms": ms1, "hop2_expand_limit10_ms": ms2, "envs_per_s_two_hops": n_env / (ms1 + ms2) * 1e3,
           "allgather_bytes_per_rank": {"hop1": [4 * n_env * rec[0], 4 * n_env * rec[1]], "hop2": [4 * n_env * rec[2], 4 * n_env * rec[3]]},
           "equal_to_replicated": {"hop1": ok1, "hop2": ok2}}
    del sg
    return res


def c5_inputs(a, n_env, n_seeds, dev, seed=7):
    """Seeds (80 % entities, 20 % L1 summaries: the index holds both), descending scores, one hop-2 op per environment."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    e_lo, e_hi = a["ranges"]["entity"]
    l_lo, l_hi = a["ranges"]["l1"]
    seeds = torch.where(torch.rand((n_env, n_seeds), generator=gen, device=dev) < 0.8,
                        torch.randint(e_lo, e_hi, (n_env, n_seeds), generator=gen, device=dev),
                        torch.randint(l_lo, l_hi, (n_env, n_seeds), generator=gen, device=dev)).to(torch.int32)
    scores = torch.sort(torch.rand((n_env, n_seeds), generator=gen, device=dev) * 0.9 + 0.05, dim=1, descending=True).values
    ops = torch.randint(0, 3, (n_env,), generator=gen, device=dev).to(torch.int32)
    return seeds, scores.to(torch.float32).contiguous(), ops


def main():
    assert torch.cuda.is_available(), "no CUDA device: the engine has no CPU fallback"
    which = sys.argv[1:] or ["c3x", "c4", "c5", "c2"]
    if any(w in ("c4dist", "c5dist") for w in which):  # torchrun workloads
        for w in which:
            {"c4dist": run_c4dist, "c5dist": run_c5dist}[w]()
            torch.cuda.empty_cache()
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for w in which:
        if w == "c3x":
            run_c3x(dev)
        elif w == "c3c":
            run_c3c(dev)
        elif w == "c4":
            run_c4(dev)
        elif w == "c5":
            run_c5(dev, float(os.environ.get("AHR_C5_SCALE", "1.0")))
        elif w == "c2":
            run_c2(dev)
        elif w == "shard8":
            run_shard8(dev)
        elif w == "adv":
            run_adv(dev)
        elif w == "qsweep":
            run_qsweep(dev)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
