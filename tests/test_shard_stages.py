"""CPU: the staged expansion over a node-range-sharded CSR (oracle/shard_stages.py, the numpy twin of csrc/graph_shard.cu)
equals the reference's expand_* on the whole graph (oracle/array_graph.expand, environment.py:393-480) for every shard
count; and the same through dist.ShardedGraph at world_size 2 over gloo."""
import copy
import os
import socket

import numpy as np
import pytest

from oracle import array_graph, shard_stages
from tests.graph_util import random_csr


def _batch(rng, a, n_env, max_ids):
    n, s0 = a["n_nodes"], int(np.argmax(a["node_type"] == 1))
    ops, ids, offs = [], [], [0]
    for e in range(n_env):
        k = int(rng.integers(0, max_ids + 1))
        pick = [int(x) for x in rng.integers(0, n, size=k)]
        if e % 5 == 0:
            pick = ([s0, s0 + 1] + pick)[:max_ids]       # hub summaries first
        if e % 7 == 0 and pick:
            pick[-1] = -1 if e % 2 else n + 3            # ids outside the graph
        if e % 9 == 0 and len(pick) >= 2:
            pick[1] = pick[0]                            # the same node twice
        ops.append(e % 3)
        ids += pick
        offs.append(len(ids))
    return np.array(ops, np.int32), np.array(ids, np.int32), np.array(offs, np.int32)


def _run_sharded(a, bounds, ops, ids, offs, max_ids, limit):
    world = len(bounds) - 1
    g1 = np.stack([shard_stages.stage1(a, bounds[r], bounds[r + 1], ops, ids, offs, max_ids, limit) for r in range(world)])
    g2 = np.stack([shard_stages.stage2(a, bounds[r], bounds[r + 1], bounds, ops, ids, offs, max_ids, limit, g1) for r in range(world)])
    return shard_stages.merge(a["n_nodes"], bounds, ops, ids, offs, max_ids, limit, g1, g2)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("limit", [1, 3, 10, 32])
def test_staged_expansion_equals_whole_graph_expand(world, limit):
    rng = np.random.default_rng(100 * world + limit)
    a = random_csr(rng, n_ent=700, n_sum=90, n_hyp=260, hub_deg=400)
    n = a["n_nodes"]
    bounds = [n * r // world for r in range(world)] + [n]
    max_ids = 4
    ops, ids, offs = _batch(rng, a, 240, max_ids)
    oi, oc = _run_sharded(a, bounds, ops, ids, offs, max_ids, limit)
    for e in range(len(ops)):
        want = array_graph.expand(a, int(ops[e]), ids[offs[e]:offs[e + 1]], limit)
        assert list(oi[e, :oc[e]]) == want, (world, limit, e, int(ops[e]), list(ids[offs[e]:offs[e + 1]]))
        assert (oi[e, oc[e]:] == -1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ah_rag_b200.dist import ShardedGraph, shard_bounds
        rng = np.random.default_rng(7)
        a = random_csr(rng, n_ent=500, n_sum=60, n_hyp=200, hub_deg=300)
        n = a["n_nodes"]
        b = [shard_bounds(n, world, r)[0] for r in range(world)] + [n]
        lo, hi = b[rank], b[rank + 1]
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x))
        npy = lambda x: x.numpy()

        def stage1(ops, ids, offs, max_ids, limit):
            return t(shard_stages.stage1(a, lo, hi, npy(ops), npy(ids), npy(offs), max_ids, limit))

        def stage2(world_, bounds, ops, ids, offs, max_ids, limit, g1):
            return t(shard_stages.stage2(a, lo, hi, [int(x) for x in bounds], npy(ops), npy(ids), npy(offs), max_ids, limit, npy(g1)))

        def merge(world_, bounds, ops, ids, offs, max_ids, limit, g1, g2):
            oi, oc = shard_stages.merge(n, [int(x) for x in bounds], npy(ops), npy(ids), npy(offs), max_ids, limit, npy(g1), npy(g2))
            return t(oi), t(oc)

        def a_stage1(seeds, prm):
            return t(shard_stages.anchor_stage1(a, lo, hi, npy(seeds), 2, prm))

        def a_stage2(world_, bounds, seeds, prm, g1):
            return t(shard_stages.anchor_stage2(a, lo, hi, [int(x) for x in bounds], npy(seeds), 2, prm, npy(g1)))

        def a_merge(world_, bounds, seeds, scores, prm, top_k, g1, g2, want_cluster=True):
            return shard_stages.anchor_merge(n, [int(x) for x in bounds], npy(seeds), npy(scores), top_k, 2, prm, npy(g1), npy(g2))

        sg = ShardedGraph(stage1, stage2, merge, torch.tensor(b, dtype=torch.int64), None, a_stage1, a_stage2, a_merge)
        assert sg.world == world and sg.rank == rank
        ops, ids, offs = _batch(np.random.default_rng(11), a, 120, 3)
        ok = True
        for limit in (10, 3):
            oi, oc = sg.expand(t(ops), t(ids), t(offs), limit=limit, max_ids=3)
            oi, oc = oi.numpy(), oc.numpy()
            for e in range(len(ops)):
                ok &= list(oi[e, :oc[e]]) == array_graph.expand(a, int(ops[e]), ids[offs[e]:offs[e + 1]], limit)
        # hop 1 over the sharded graph (the injected `params` is just member_top_m here)
        seeds = np.stack([rng.choice(560, 8, replace=False) for _ in range(30)]).astype(np.int32)
        scores = np.sort(rng.uniform(0.1, 0.9, (30, 8)).astype(np.float32), axis=1)[:, ::-1].copy()
        got = sg.anchor_rerank(t(seeds), t(scores), 5, 12)
        for e in range(30):
            ok &= got[e] == array_graph.anchor_rerank(a, seeds[e], scores[e], 12, 5)
        q.put((rank, bool(ok)))
    except Exception as e:  # report instead of leaving the parent waiting on the queue
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_graph_world2_gloo():
    """The two all-gathers of dist.ShardedGraph over gloo, world_size 2: every rank ends with the whole-graph answer."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


@pytest.mark.parametrize("world", [1, 2, 5])
def test_staged_anchor_rerank_equals_whole_graph(world):
    """Hop 1 over the sharded graph (numpy twin of shard_anchor_stage1/2 + the rerank): equal to the reference's anchor
    expansion + rerank on the whole graph (oracle/array_graph.anchor_rerank, hierarchical_graph.py:719-789)."""
    rng = np.random.default_rng(40 + world)
    a = random_csr(rng, n_ent=600, n_sum=80, n_hyp=100, hub_deg=300)
    n = a["n_nodes"]
    a["judge"][600:680] = np.round(rng.uniform(0, 10, 80), 1)
    a["conf"][600:680] = np.round(rng.uniform(0, 10, 80), 2)
    a["judge_term"][600:680] = [1.0 / (1.0 + np.exp(-(x / 10.0))) for x in a["judge"][600:680]]
    bounds = [n * r // world for r in range(world)] + [n]
    nq, ns, top_k, m = 40, 12, 20, 5
    seeds = np.stack([rng.choice(680, ns, replace=False) for _ in range(nq)]).astype(np.int32)
    seeds[3, 4] = -1
    seeds[5, 0] = seeds[5, 7]  # the same node twice: the later seed owns the value, the first the position
    scores = np.sort(rng.uniform(0.1, 0.9, (nq, ns)).astype(np.float32), axis=1)[:, ::-1].copy()
    g1 = np.stack([shard_stages.anchor_stage1(a, bounds[r], bounds[r + 1], seeds, 2, m) for r in range(world)])
    g2 = np.stack([shard_stages.anchor_stage2(a, bounds[r], bounds[r + 1], bounds, seeds, 2, m, g1) for r in range(world)])
    got = shard_stages.anchor_merge(n, bounds, seeds, scores, top_k, 2, m, g1, g2)
    for q in range(nq):
        assert got[q] == array_graph.anchor_rerank(a, seeds[q], scores[q], top_k, m), (world, q)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_node_range_bounds_cover_and_balance(world):
    """dist.node_range_bounds: contiguous node ranges that cover [0, n), identical from numpy and torch inputs, and -- with
    balance="edges" -- of near-equal CSR volume on a graph whose adjacency sits in a narrow id range (the c5 shape)."""
    import torch
    from ah_rag_b200.dist import node_range_bounds
    rng = np.random.default_rng(world)
    a = random_csr(rng, n_ent=4000, n_sum=400, n_hyp=300, hub_deg=3000)
    n = a["n_nodes"]
    for balance in ("nodes", "edges"):
        b = node_range_bounds(a, world, balance)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        at = {"n_nodes": n, "row_ptr": [torch.from_numpy(np.asarray(x)) for x in a["row_ptr"]]}
        assert node_range_bounds(at, world, balance) == b
    b = node_range_bounds(a, world, "edges")
    vol = [sum(int(rp[b[r + 1]] - rp[b[r]]) for rp in a["row_ptr"]) + 10 * (b[r + 1] - b[r]) for r in range(world)]
    hub = max(sum(int(rp[v + 1] - rp[v]) for rp in a["row_ptr"]) for v in (4000, 4001)) + 10  # one row cannot be split
    assert max(vol) <= sum(vol) / world + hub
    with pytest.raises(ValueError):
        node_range_bounds(a, 2, "rows")
