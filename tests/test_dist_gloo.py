"""CPU, world_size 2, gloo: the node-range sharding + single all-gather + merge plumbing of ShardedIndex.
The local top-k and the merge are the oracle's (the CUDA kernels need a GPU); what is under test is the
host-side exchange: shard bounds, candidate packing, one all_gather_into_tensor, identical results on all ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, nq, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ah_rag_b200 import synth
        from ah_rag_b200.dist import ShardedIndex, shard_bounds
        from oracle import exact
        X, Q = synth.unit_rows(n, dim, 21, clustered=4), synth.unit_rows(nq, dim, 22, clustered=4)
        lo, hi = shard_bounds(n, world, rank)

        def local_topk(qt, kk, mode):
            s, i = exact.exact_topk(X[lo:hi], qt.numpy(), kk, id_base=lo)
            return torch.from_numpy(s), torch.from_numpy(i)

        def merge(all_s, all_i):
            s, i = exact.merge_topk(list(all_s.numpy()), list(all_i.numpy()), k)
            return torch.from_numpy(s), torch.from_numpy(i)

        sh = ShardedIndex(local_topk, merge)
        assert sh.world == world and sh.rank == rank
        # no GPU here: the peer-memory rendezvous must fail on every rank TOGETHER (collective-safe) and leave the
        # all-gather path in place
        assert sh.enable_peer_exchange(16, kk_max := 32) is False and sh.peer is None and kk_max
        s, i = sh.topk(torch.from_numpy(Q), k, "exact")
        ws, wi = exact.exact_topk(X, Q, k)
        ok = np.array_equal(i.numpy(), wi) and np.array_equal(s.numpy().view(np.uint32), ws.view(np.uint32))
        q.put((rank, bool(ok)))
    except Exception as e:  # report instead of leaving the parent waiting on the queue
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_topk_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 3001, 32, 9, 20, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]
