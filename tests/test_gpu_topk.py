"""GPU parity tests of the similarity + top-k path, through the C ABI (ctypes), against the CPU oracle."""
import numpy as np
import pytest
import torch

from ah_rag_b200 import synth
from oracle import exact

pytestmark = pytest.mark.gpu


def _bf16(x):
    return torch.from_numpy(x).cuda().to(torch.bfloat16).contiguous()


def _bits(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.fixture(scope="module")
def engine():
    from ah_rag_b200 import engine as e
    return e


@pytest.mark.parametrize("n,dim,nq", [(1000, 64, 5), (4096, 128, 128), (5000, 1024, 130), (777, 384, 1), (300, 72, 17)])
def test_tensor_core_scores_match_canonical(engine, n, dim, nq):
    """Every accumulator the tcgen05 GEMM produces (descriptor / swizzle / TMEM addressing check)."""
    X, Q = _bf16(synth.unit_rows(n, dim, 1)), _bf16(synth.unit_rows(nq, dim, 2))
    idx = engine.VectorIndex(X)
    got = idx.debug_tc_scores(Q).cpu().numpy()
    want = exact.canonical_scores(_bits(X), _bits(Q))
    assert not np.isnan(got).any()
    err = np.abs(got - want).max()
    # fp32 accumulation of exact bf16 products on zero-mean rows: far below the certification bound (api.cu: (dim/2+64) * 2^-23)
    assert err <= (dim / 8 + 32) * 2.0 ** -23 / 4, err


@pytest.mark.parametrize("n,dim,nq,k", [(1000, 64, 5, 10), (20000, 128, 300, 50), (5000, 1024, 130, 50),
                                        (40, 64, 3, 50), (257, 64, 129, 1), (9000, 384, 64, 100)])
def test_bf16_mode_topk(engine, n, dim, nq, k):
    X, Q = _bf16(synth.unit_rows(n, dim, 3, clustered=16)), _bf16(synth.unit_rows(nq, dim, 4, clustered=16))
    idx = engine.VectorIndex(X)
    s, i = idx.topk(Q, k, mode="bf16")
    s, i = s.cpu().numpy(), i.cpu().numpy()
    rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    kk = min(k, n)
    assert (i[:, kk:] == -1).all() and np.isneginf(s[:, kk:]).all()
    # scores within 1e-2 relative of the fp32 reference (north-star tolerance); recall@k >= 0.999
    assert np.allclose(s[:, :kk], rs[:, :kk], rtol=1e-2, atol=1e-4)
    hits = sum(len(set(i[q, :kk]) & set(ri[q, :kk])) for q in range(nq))
    assert hits / (nq * kk) >= 0.999
    # sorted by (score desc, id asc)
    assert (np.diff(s[:, :kk], axis=1) <= 0).all()


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("n,dim,nq,k", [(1000, 64, 5, 10), (6000, 128, 70, 50), (50, 32, 4, 64), (5000, 384, 9, 5)])
def test_exact_mode_ids_bit_exact(engine, dtype, n, dim, nq, k):
    Xn, Qn = synth.unit_rows(n, dim, 5, clustered=8), synth.unit_rows(nq, dim, 6, clustered=8)
    Xn[n // 2] = Xn[7]  # exact duplicates: tie -> lower id
    Xn[n - 1] = Xn[7]
    Qn[0] = Xn[7]
    if dtype == "bf16":
        X, Q = _bf16(Xn), _bf16(Qn)
        rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    else:
        X, Q = torch.from_numpy(Xn).cuda(), torch.from_numpy(Qn).cuda()
        rs, ri = exact.exact_topk(Xn, Qn, k)
    idx = engine.VectorIndex(X)
    s, i = idx.topk(Q, k, mode="exact")
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))


def test_exact_mode_uses_tensor_core_candidates_and_certifies(engine):
    n, dim, nq, k = 60000, 256, 200, 50
    X, Q = _bf16(synth.unit_rows(n, dim, 7)), _bf16(synth.unit_rows(nq, dim, 8))
    idx = engine.VectorIndex(X)
    s, i = idx.topk(Q, k, mode="exact")
    cert, fb = idx.exact_stats()
    assert cert + fb == nq and cert >= nq * 0.9  # random data: nearly every query certifies
    rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))


def test_exact_mode_fallback_on_near_ties(engine):
    """Rows that differ by less than the tensor-core error bound cannot be certified -> brute-force fallback."""
    n, dim, nq, k = 8192, 128, 6, 10
    Xn = synth.unit_rows(n, dim, 9)
    Xn[1000:1200] = Xn[999]  # 201 identical rows: the k-th / kc-th gap is zero
    Qn = synth.unit_rows(nq, dim, 10)
    Qn[0] = Xn[999]
    X, Q = _bf16(Xn), _bf16(Qn)
    idx = engine.VectorIndex(X)
    s, i = idx.topk(Q, k, mode="exact")
    cert, fb = idx.exact_stats()
    assert fb >= 1
    rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert list(i[0].cpu().numpy()) == list(range(999, 999 + k))


def test_exact_mode_second_tier_resolves_moderate_tie_groups(engine):
    """A group of identical rows straddling rank k defeats the first tier (6 spare candidates) but not the second
    (62 spare): certified there, no brute-force scan, ties still broken by the lower id."""
    n, dim, nq, k = 20000, 128, 5, 50
    Xn = synth.unit_rows(n, dim, 13)
    Qn = synth.unit_rows(nq, dim, 14)
    near = Qn[0] + 0.05 * synth.unit_rows(1, dim, 15)[0]
    near /= np.linalg.norm(near)
    Xn[3000:3040] = Qn[0]          # ranks 1..40 of query 0
    Xn[7000:7030] = near           # ranks 41..70: rank 50 falls inside this group of equal scores
    X, Q = _bf16(Xn), _bf16(Qn)
    idx = engine.VectorIndex(X)
    s, i = idx.topk(Q, k, mode="exact")
    cert, fb = idx.exact_stats()
    assert (cert, fb) == (nq, 0)
    rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))
    assert list(i[0, 40:].cpu().numpy()) == list(range(7000, 7010))


def test_topk_host_entry_point(engine):
    n, dim, nq, k = 3000, 128, 33, 20
    X = _bf16(synth.unit_rows(n, dim, 11))
    Qh = torch.from_numpy(synth.unit_rows(nq, dim, 12)).to(torch.bfloat16).pin_memory()
    idx = engine.VectorIndex(X, id_base=1_000_000)
    s, i = idx.topk_host(Qh, k, mode="exact")
    rs, ri = exact.exact_topk(_bits(X), _bits(Qh), k, id_base=1_000_000)
    assert np.array_equal(i.numpy(), ri) and np.array_equal(s.numpy().view(np.uint32), rs.view(np.uint32))


def test_sharded_merge_equals_single_index(engine):
    """Node-range shards + device k-way merge == one index (what the NCCL path computes), fp32-exact."""
    n, dim, nq, k = 10000, 64, 40, 50
    Xn, Qn = synth.unit_rows(n, dim, 13, clustered=4), synth.unit_rows(nq, dim, 14, clustered=4)
    X, Q = _bf16(Xn), _bf16(Qn)
    full_s, full_i = engine.VectorIndex(X).topk(Q, k, mode="exact")
    for shards in (2, 4, 8):
        bounds = [n * g // shards for g in range(shards + 1)]
        parts = [engine.VectorIndex(X[bounds[g]:bounds[g + 1]].contiguous(), id_base=bounds[g]).topk(Q, k, mode="exact")
                 for g in range(shards)]
        ms, mi = engine.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert torch.equal(mi, full_i) and torch.equal(ms, full_s)


def test_empty_and_degenerate_inputs(engine):
    dim = 64
    X = _bf16(synth.unit_rows(10, dim, 15))
    idx = engine.VectorIndex(X)
    s, i = idx.topk(_bf16(synth.unit_rows(0, dim, 1).reshape(0, dim)), 5, mode="bf16")
    assert s.shape == (0, 5)
    with pytest.raises(RuntimeError):
        idx.topk(torch.zeros((2, dim), dtype=torch.bfloat16), 5)  # host tensor: no CPU path
    with pytest.raises(RuntimeError):
        idx.topk(_bf16(synth.unit_rows(2, dim, 1)), 5000, mode="exact")  # k too large -> library error


def test_round4_device_matches_cpython(engine):
    import random
    random.seed(3)
    xs = [random.random() * 2 for _ in range(20000)]
    xs += [round(random.random(), 4) + 0.00005 for _ in range(20000)]
    xs += [0.0, 0.00035, 5e-05, 1.00005, 2.5e-5, -0.00035, 123456.78905]
    got = engine.round4_device(torch.tensor(xs, dtype=torch.float64).cuda()).cpu().numpy()
    want = np.array([round(x, 4) for x in xs])
    assert np.array_equal(got, want)


def test_packed_merge_equals_plain_merge(engine):
    """The exchange-buffer fast path of the sharded index (no re-pack copies) == the [parts, nq, k] merge."""
    n, dim, nq, k = 6000, 64, 32, 50
    X, Q = _bf16(synth.unit_rows(n, dim, 31, clustered=4)), _bf16(synth.unit_rows(nq, dim, 32, clustered=4))
    shards = 4
    bounds = [n * g // shards for g in range(shards + 1)]
    rec = nq * k * 12
    packed = torch.empty(shards * rec, dtype=torch.uint8, device="cuda")
    parts = []
    for g in range(shards):
        mine = packed[g * rec:(g + 1) * rec]
        out = (mine[: nq * k * 4].view(torch.float32).view(nq, k), mine[nq * k * 4:].view(torch.int64).view(nq, k))
        idx = engine.VectorIndex(X[bounds[g]:bounds[g + 1]].contiguous(), id_base=bounds[g])
        parts.append(idx.topk(Q, k, mode="exact", out=out))
    ps, pi = engine.merge_topk_packed(packed, shards, nq, k)
    ms, mi = engine.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(ps, ms) and torch.equal(pi, mi)
    fs, fi = engine.VectorIndex(X).topk(Q, k, mode="exact")
    assert torch.equal(pi, fi) and torch.equal(ps, fs)


@pytest.mark.parametrize("n,dim,nq", [(5000, 384, 40), (9000, 128, 257)])
def test_fp32_index_tensor_core_scores_hi_lo_split(engine, n, dim, nq):
    """fp32 rows run on the tensor cores as hi + lo bf16 halves (3 operand segments): accumulators must be within
    the certification bound of the canonical fp32 score, and exact mode must certify and stay bit-exact."""
    Xn, Qn = synth.unit_rows(n, dim, 61, clustered=8), synth.unit_rows(nq, dim, 62, clustered=8)
    X, Q = torch.from_numpy(Xn).cuda(), torch.from_numpy(Qn).cuda()
    idx = engine.VectorIndex(X)
    got = idx.debug_tc_scores(Q).cpu().numpy()
    want = exact.canonical_scores(Xn, Qn)
    bound = (3 * dim / 2 + 64) * 2.0 ** -23 + 2.0 ** -16  # api.cu's eps_rel for fp32 rows
    assert np.abs(got - want).max() <= bound / 4, np.abs(got - want).max()
    k = 10
    s, i = idx.topk(Q, k, mode="exact")
    cert, fb = idx.exact_stats()
    assert cert >= nq * 0.9 and cert + fb == nq
    rs, ri = exact.exact_topk(Xn, Qn, k)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))


@pytest.mark.parametrize("kind", ["positive", "correlated", "large_norm"])
@pytest.mark.parametrize("dim", [1024, 2048, 4096])
def test_certification_bound_holds_on_same_sign_inputs(engine, kind, dim):
    """the exact-mode certificate rests on |tensor - canonical| <= eps_rel * ||q|| * ||x||.  Zero-mean random rows
    let the tcgen05 truncation errors cancel; all-positive / strongly correlated / large-norm rows make them add up with one
    sign.  The accumulators must stay within a quarter of the bound api.cu uses, (dim/2 + 64) * 2^-23, and exact mode must
    remain bit-exact on these inputs."""
    n, nq, k = 4608, 96, 10
    rng = np.random.default_rng(dim + len(kind))
    if kind == "positive":
        Xn, Qn = np.abs(rng.standard_normal((n, dim))), np.abs(rng.standard_normal((nq, dim)))
    elif kind == "correlated":
        c = rng.standard_normal(dim)
        Xn, Qn = c + 0.05 * rng.standard_normal((n, dim)), c + 0.05 * rng.standard_normal((nq, dim))
    else:
        Xn, Qn = np.abs(rng.standard_normal((n, dim))) * 37.0, np.abs(rng.standard_normal((nq, dim))) * 0.01
    Xn, Qn = Xn.astype(np.float32), Qn.astype(np.float32)
    if kind != "large_norm":
        Xn /= np.linalg.norm(Xn, axis=1, keepdims=True)
        Qn /= np.linalg.norm(Qn, axis=1, keepdims=True)
    X, Q = _bf16(Xn), _bf16(Qn)
    idx = engine.VectorIndex(X)
    got = idx.debug_tc_scores(Q).cpu().numpy().astype(np.float64)
    want = exact.canonical_scores(_bits(X), _bits(Q)).astype(np.float64)
    xn = np.linalg.norm(X.float().cpu().numpy().astype(np.float64), axis=1)
    qn = np.linalg.norm(Q.float().cpu().numpy().astype(np.float64), axis=1)
    eps = (dim / 2 + 64) * 2.0 ** -23 * qn[:, None] * xn[None, :]
    ratio = (np.abs(got - want) / eps).max()
    assert ratio <= 0.25, ratio
    s, i = idx.topk(Q, k, mode="exact")
    rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))


@pytest.mark.parametrize("parts,k,nq", [(2, 50, 300), (8, 50, 64), (16, 100, 9), (3, 5, 1), (1, 50, 17)])
def test_merge_of_sorted_lists_equals_oracle_merge(engine, parts, k, nq):
    """ahr_merge_topk over per-shard SORTED lists takes the rank-counting merge kernel: ragged lists (empties at the end),
    exact score ties across lists (lower id first), every list empty, k larger than the number of candidates."""
    rng = np.random.default_rng(parts * 100 + k)
    scores = np.full((parts, nq, k), -np.inf, dtype=np.float32)
    ids = np.full((parts, nq, k), -1, dtype=np.int64)
    pool = rng.choice(np.round(rng.uniform(-1, 1, 64), 2).astype(np.float32), size=(parts, nq, k))  # many ties
    for p in range(parts):
        for q in range(nq):
            n = int(rng.integers(0, k + 1)) if (p + q) % 3 else k
            if q == 0:
                n = 0  # a query without any candidate
            row_ids = rng.choice(10_000, size=n, replace=False).astype(np.int64) * parts + p  # unique across lists
            order = np.lexsort((row_ids, -pool[p, q, :n].astype(np.float64)))
            scores[p, q, :n] = pool[p, q, :n][order]
            ids[p, q, :n] = row_ids[order]
    ms, mi = engine.merge_topk(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda())
    ws, wi = exact.merge_topk([scores[p] for p in range(parts)], [ids[p] for p in range(parts)], k)
    assert np.array_equal(mi.cpu().numpy(), wi)
    assert np.array_equal(ms.cpu().numpy().view(np.uint32), ws.view(np.uint32))


def test_exact_mode_in_two_halves_equals_the_one_call_form(engine):
    """ahr_topk_exact_begin enqueues tier 1 without synchronising; other work can be queued before ahr_topk_exact_finish.
    Cases: every query certified (finish only waits), near-duplicate rows (fallback tiers run inside finish), a shape the
    tensor path does not take (brute force enqueued by begin)."""
    rng = np.random.default_rng(3)
    for n, dim, nq, k, dup in ((20000, 128, 300, 50, False), (9000, 64, 70, 50, True), (3000, 64, 9, 10, False)):
        Xn, Qn = synth.unit_rows(n, dim, 71, clustered=8), synth.unit_rows(nq, dim, 72, clustered=8)
        if dup:  # 150 copies of one row around every query's rank k: the tensor-core scores tie at both candidate widths
            Xn[100:250] = Xn[100]  # (56 and 112), certification must fail twice and the brute-force scan decides
            Qn[:] = Xn[100] + 0.01 * rng.standard_normal((nq, dim)).astype(np.float32)
            Qn /= np.linalg.norm(Qn, axis=1, keepdims=True)
        X, Q = _bf16(Xn), _bf16(Qn)
        idx = engine.VectorIndex(X)
        s, i = idx.topk_exact_begin(Q, k)
        other = idx.rows.float().sum()  # unrelated work queued between the halves
        with pytest.raises(RuntimeError):
            idx.topk_exact_begin(Q, k)  # one pending exact query per index
        idx.topk_exact_finish()
        cert, fb = idx.exact_stats()
        assert cert + fb == nq and (fb > 0) == (dup or n <= 4096), (cert, fb)
        rs, ri = exact.exact_topk(_bits(X), _bits(Q), k)
        assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(s.cpu().numpy().view(np.uint32), rs.view(np.uint32))
        ws, wi = idx.topk(Q, k, mode="exact")
        assert torch.equal(wi, i) and torch.equal(ws, s) and bool(torch.isfinite(other))
