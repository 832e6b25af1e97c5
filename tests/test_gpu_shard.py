"""GPU: frontier expansion over a CSR sharded by node range (csrc/graph_shard.cu).  The ranks of a world are emulated on
one GPU as independent shard handles (the stages of different owners do not wait on one another: stage outputs are simply
stacked in rank order, which is what the all-gather produces); the real 2-GPU exchange is tests/test_gpu_multi.py."""
import numpy as np
import pytest
import torch

from oracle import array_graph, shard_stages
from tests.graph_util import random_csr
from tests.test_shard_stages import _batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 5])
@pytest.mark.parametrize("limit", [1, 10, 32])
def test_sharded_expand_equals_unsharded_and_oracle(world, limit):
    from ah_rag_b200 import engine
    rng = np.random.default_rng(1000 * world + limit)
    a = random_csr(rng, n_ent=3000, n_sum=300, n_hyp=900, hub_deg=1500)
    n = a["n_nodes"]
    bounds = [n * r // world for r in range(world)] + [n]
    max_ids = 4
    ops, ids, offs = _batch(rng, a, 600, max_ids)
    dev = torch.device("cuda", 0)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    d_ops, d_ids, d_offs, d_bounds = t(ops), t(ids), t(offs), t(np.array(bounds, np.int64))
    shards = [engine.GraphShard(a, bounds[r], bounds[r + 1], dev) for r in range(world)]
    assert sum(s.nbytes() for s in shards) < 1.2 * engine.GraphCSR(a).nbytes()  # the shards partition the graph
    g1 = torch.stack([s.stage1(d_ops, d_ids, d_offs, max_ids, limit) for s in shards])
    g2 = torch.stack([s.stage2(world, d_bounds, d_ops, d_ids, d_offs, max_ids, limit, g1) for s in shards])
    oi, oc = shards[0].merge(world, d_bounds, d_ops, d_ids, d_offs, max_ids, limit, g1, g2)
    # (1) the unsharded kernel on the whole graph
    wi, wc = engine.GraphCSR(a).expand(d_ops, d_ids, d_offs, limit)
    assert torch.equal(oc, wc) and torch.equal(oi, wi)
    # (2) the reference semantics (environment.py:393-480) through the array oracle
    oi, oc = oi.cpu().numpy(), oc.cpu().numpy()
    for e in range(0, len(ops), 3):
        assert list(oi[e, :oc[e]]) == array_graph.expand(a, int(ops[e]), ids[offs[e]:offs[e + 1]], limit), (world, limit, e)
    # (3) record for record against the numpy twin of the stages (header + the meaningful part of each record)
    cap, r1, r2 = shard_stages.geom(limit, max_ids)
    n1 = np.stack([shard_stages.stage1(a, bounds[r], bounds[r + 1], ops, ids, offs, max_ids, limit) for r in range(world)])
    n2 = np.stack([shard_stages.stage2(a, bounds[r], bounds[r + 1], bounds, ops, ids, offs, max_ids, limit, n1) for r in range(world)])
    h1 = g1.cpu().numpy().reshape(world, len(ops), max_ids, r1)
    w1 = n1.reshape(world, len(ops), max_ids, r1)
    assert np.array_equal(h1[..., :4], w1[..., :4])
    for r, e, i in zip(*np.nonzero(w1[..., 1])):
        nc, nh = w1[r, e, i, 2], w1[r, e, i, 3]
        assert np.array_equal(h1[r, e, i, 4:4 + nc], w1[r, e, i, 4:4 + nc])
        assert np.array_equal(h1[r, e, i, 4 + cap:4 + cap + nh], w1[r, e, i, 4 + cap:4 + cap + nh])
    h2 = g2.cpu().numpy().reshape(world, len(ops), max_ids * limit, r2)
    w2 = n2.reshape(world, len(ops), max_ids * limit, r2)
    assert np.array_equal(h2[..., :3], w2[..., :3])
    for r, e, i in zip(*np.nonzero(w2[..., 1])):
        assert np.array_equal(h2[r, e, i, 4:4 + w2[r, e, i, 2]], w2[r, e, i, 4:4 + w2[r, e, i, 2]])


def test_sharded_expand_rejects_large_limits():
    from ah_rag_b200 import engine
    a = random_csr(np.random.default_rng(3), n_ent=100, n_sum=10, n_hyp=20, hub_deg=20)
    sh = engine.GraphShard(a, 0, a["n_nodes"], torch.device("cuda", 0))
    with pytest.raises(ValueError):
        sh.record_ints(1, 33, 4)


@pytest.mark.parametrize("world", [1, 2, 5])
@pytest.mark.parametrize("caps", [(2, 5), (0, 3), (4, 12)])
def test_sharded_anchor_rerank_equals_unsharded(world, caps):
    """Hop 1 (hierarchical_graph.py:719-789) over the sharded graph: stage records equal the numpy twin, the rerank fed
    from them equals the unsharded kernel output for output (every tensor, bit for bit) and the array oracle."""
    from ah_rag_b200 import engine
    parent_cap, m = caps
    rng = np.random.default_rng(77 * world + m)
    a = random_csr(rng, n_ent=3000, n_sum=300, n_hyp=400, hub_deg=600)
    n = a["n_nodes"]
    a["judge"][3000:3300] = np.round(rng.uniform(0, 10, 300), 1)
    a["conf"][3000:3300] = np.round(rng.uniform(0, 10, 300), 2)
    a["judge_term"][3000:3300] = [1.0 / (1.0 + np.exp(-(x / 10.0))) for x in a["judge"][3000:3300]]
    bounds = [n * r // world for r in range(world)] + [n]
    nq, ns, top_k = 96, 20, 30
    seeds = np.stack([rng.choice(3300, ns, replace=False) for _ in range(nq)]).astype(np.int32)
    seeds[3, 4] = -1
    seeds[5, 0] = seeds[5, 7]
    scores = np.sort(rng.uniform(0.1, 0.9, (nq, ns)).astype(np.float32), axis=1)[:, ::-1].copy()
    dev = torch.device("cuda", 0)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    d_seeds, d_scores, d_bounds = t(seeds), t(scores), t(np.array(bounds, np.int64))
    prm = engine.SearchParams(member_top_m=m, parent_cap=parent_cap)
    shards = [engine.GraphShard(a, bounds[r], bounds[r + 1], dev) for r in range(world)]
    g1 = torch.stack([s.anchor_stage1(d_seeds, prm) for s in shards])
    g2 = torch.stack([s.anchor_stage2(world, d_bounds, d_seeds, prm, g1) for s in shards])
    got = shards[-1].anchor_merge(world, d_bounds, d_seeds, d_scores, prm, top_k, g1, g2)
    want = engine.GraphCSR(a).anchor_rerank(d_seeds, d_scores, prm, top_k)
    cnt, ecnt = want["count"].cpu().numpy(), want["exp_count"].cpu().numpy()
    assert torch.equal(got["count"], want["count"]) and torch.equal(got["exp_count"], want["exp_count"])
    assert torch.equal(got["seed_semantic"].view(torch.int64), want["seed_semantic"].view(torch.int64))
    for key, c in (("nodes", cnt), ("semantic", cnt), ("score", cnt), ("exp_nodes", ecnt), ("exp_semantic", ecnt), ("exp_kind", ecnt)):
        g, w = got[key].cpu().numpy(), want[key].cpu().numpy()
        for q in range(nq):
            assert np.array_equal(g[q, :c[q]].view(np.uint8), w[q, :c[q]].view(np.uint8)), (key, q)
    # the reference semantics through the array oracle
    gn, gs, gc = got["nodes"].cpu().numpy(), got["semantic"].cpu().numpy(), got["score"].cpu().numpy()
    for q in range(0, nq if parent_cap == 2 else 0, 5):  # the reference's parent cap is the literal 2 (:742)
        ref = array_graph.anchor_rerank(a, seeds[q], scores[q], top_k, m)
        assert [(int(x), float(y), float(z)) for x, y, z in zip(gn[q, :cnt[q]], gs[q, :cnt[q]], gc[q, :cnt[q]])] == ref, q
    # records against the numpy twin
    fan, r1, r2 = shard_stages.anchor_geom(ns, parent_cap, m)
    n1 = np.stack([shard_stages.anchor_stage1(a, bounds[r], bounds[r + 1], seeds, parent_cap, m) for r in range(world)])
    h1 = g1.cpu().numpy().reshape(world, nq, ns, r1)
    w1 = n1.reshape(world, nq, ns, r1)
    assert np.array_equal(h1[..., :4], w1[..., :4])
    for r, q, i in zip(*np.nonzero(w1[..., 0])):
        assert np.array_equal(h1[r, q, i, 4:10], w1[r, q, i, 4:10])
        assert np.array_equal(h1[r, q, i, 10:10 + w1[r, q, i, 2]], w1[r, q, i, 10:10 + w1[r, q, i, 2]])
    if fan:
        n2 = np.stack([shard_stages.anchor_stage2(a, bounds[r], bounds[r + 1], bounds, seeds, parent_cap, m, n1) for r in range(world)])
        h2 = g2.cpu().numpy().reshape(world, nq, ns * fan, r2)
        w2 = n2.reshape(world, nq, ns * fan, r2)
        assert np.array_equal(h2[..., 0], w2[..., 0])  # presence; absent records carry no meaning beyond it
        for r, q, i in zip(*np.nonzero(w2[..., 0])):
            assert np.array_equal(h2[r, q, i], w2[r, q, i])
