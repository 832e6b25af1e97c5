"""Node-range sharding of the embedding index across the GPUs of one box (one process per GPU).

Each rank holds rows [lo, hi) of the corpus, computes its local top-k with the fused kernel, and the
candidate lists meet through ONE exchange -- stores straight into the peers' memory over NVLink
(``engine.PeerExchange`` / csrc/exchange.cu) on a GPU box, an all-gather (NCCL, or gloo in the CPU tests) otherwise
or when ``AHR_EXCHANGE=nccl`` -- followed by a device-side k-way merge with the total order (score desc, id asc) -- so fp32-exact results are identical
for every shard count.  Queries are replicated (every rank receives the same batch).

The reference has no distributed path at all; this is the B200-native scaling of
the ``Collection.query`` seam (hierarchical_graph.py:714).
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous node-id range of ``rank``: [n*rank/world, n*(rank+1)/world)."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def node_range_bounds(arrays, world: int, balance: str = "edges") -> list:
    """First node id of every rank (+ n_nodes): contiguous ranges of equal node count or of equal CSR volume
    (entries of all six slots + 40 bytes-worth of per-node tables)."""
    n = int(arrays["n_nodes"])
    if balance == "nodes" or world == 1 or n == 0:
        return [shard_bounds(n, world, r)[0] for r in range(world)] + [n]
    if balance != "edges":
        raise ValueError(f"balance must be 'nodes' or 'edges' (got {balance!r})")
    rps = arrays["row_ptr"]
    if isinstance(rps[0], torch.Tensor):
        cum = sum(rp.to(torch.int64) for rp in rps) + 10 * torch.arange(n + 1, device=rps[0].device, dtype=torch.int64)
        targets = (cum[-1] * torch.arange(1, world, device=cum.device, dtype=torch.int64)) // world
        inner = torch.searchsorted(cum, targets).clamp_(0, n).tolist()
    else:
        import numpy as np
        cum = sum(np.asarray(rp, dtype=np.int64) for rp in rps) + 10 * np.arange(n + 1, dtype=np.int64)
        targets = (int(cum[-1]) * np.arange(1, world, dtype=np.int64)) // world
        inner = np.clip(np.searchsorted(cum, targets), 0, n).tolist()
    b = [0] + [int(x) for x in inner] + [n]
    for i in range(1, len(b)):  # monotone even for degenerate inputs
        b[i] = max(b[i], b[i - 1])
    return b


def pack_candidates(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """[nq, k] fp32 scores + [nq, k] int64 ids -> one flat byte buffer (a single collective moves both)."""
    return torch.cat([scores.contiguous().view(torch.uint8).reshape(-1), ids.contiguous().view(torch.uint8).reshape(-1)])


def unpack_candidates(buf: torch.Tensor, world: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[world, bytes] gathered buffer -> scores [world, nq, k] fp32, ids [world, nq, k] int64."""
    sb = nq * k * 4
    buf = buf.reshape(world, -1)
    scores = buf[:, :sb].contiguous().view(torch.float32).reshape(world, nq, k)
    ids = buf[:, sb:].contiguous().view(torch.int64).reshape(world, nq, k)
    return scores, ids


class ShardedIndex:
    """Top-k over a corpus sharded by node range.  ``local_topk(q, k, mode) -> (scores, ids)`` with GLOBAL ids
    and ``merge(scores[G,nq,k], ids[G,nq,k]) -> (scores, ids)`` are injectable so the plumbing can be tested
    on CPU with gloo; in production they are ``VectorIndex.topk`` and ``engine.merge_topk``."""

    def __init__(self, local_topk: Callable, merge: Callable, group: Optional[dist.ProcessGroup] = None) -> None:
        self.local_topk = local_topk
        self.merge = merge
        self.merge_packed: Optional[Callable] = None  # device fast path: merge out of the gathered buffer, no re-pack
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    @classmethod
    def from_rows(cls, rows_local: torch.Tensor, id_base: int, group: Optional[dist.ProcessGroup] = None) -> "ShardedIndex":
        from . import engine
        index = engine.VectorIndex(rows_local, id_base=id_base)
        obj = cls(index.topk, engine.merge_topk, group)
        obj.index = index
        obj.merge_packed = engine.merge_topk_packed
        return obj

    def enable_peer_exchange(self, max_queries: int, max_k: int) -> bool:
        """Switch the exchange to peer-memory stores (one-time rendezvous of CUDA IPC handles over the process group).
        Collective: every rank must call it.  Returns True on every rank or False on every rank -- a rank that cannot
        allocate, export or map a buffer (e.g. CUDA IPC not permitted in the container) still takes part in the two
        collectives below, so nobody is left waiting, and all ranks keep the all-gather.  ``AHR_EXCHANGE=nccl`` opts out."""
        import os
        from . import engine
        self.peer = None
        if self.world == 1 or os.environ.get("AHR_EXCHANGE", "peer") == "nccl":
            return False
        peer, handle, why = None, None, ""
        try:
            peer = engine.PeerExchange(self.index.device, self.rank, self.world, max_queries * max_k * 12)
            handle = peer.export_handle()
        except Exception as e:  # noqa: BLE001 -- reported below, the collectives must still run
            why = repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                peer.connect(handles)
            except Exception as e:  # noqa: BLE001
                ok, why = False, repr(e)
        flags = [None] * self.world
        dist.all_gather_object(flags, (ok, why), group=self.group)
        if all(f[0] for f in flags):
            self.peer = peer
            return True
        self.peer_failure = "; ".join(f"rank {r}: {f[1]}" for r, f in enumerate(flags) if f[1])
        if peer is not None:
            peer.close()
        return False

    def topk_async(self, q: torch.Tensor, k: int, mode: str = "bf16") -> "PendingTopk":
        """Pipelined form for back-to-back batches: the local top-k of this batch runs on the caller's stream, its exchange
        (push to the peers, wait for theirs, merge) on a side stream -- so the NEXT batch's K1 starts without waiting for
        the slowest peer of this one, and a step no longer costs the maximum over the ranks.  The two halves of the exchange
        buffer bound the run-ahead to one batch: batch i + 2 waits for the merge of batch i.  Safe across ranks because
        every rank's exchanges run in order on ONE stream: a peer pushes batch i + 2 into a half only after its merge of
        i + 1, which needed this rank's push of i + 1, which follows this rank's merge of i.
        ``.result()`` makes the caller's current stream wait and returns (scores, ids); collective like ``topk``."""
        peer = getattr(self, "peer", None)
        nq = int(q.shape[0])
        if self.world == 1 or peer is None or (nq * k) % 4 != 0 or nq * k * 12 > peer.max_record_bytes:
            return PendingTopk(*self.topk(q, k, mode), None)
        st = self.__dict__.setdefault("_async", {"side": torch.cuda.Stream(device=q.device), "done": []})
        main = torch.cuda.current_stream(q.device)
        if len(st["done"]) >= 2:
            main.wait_event(st["done"][-2])  # the half this batch writes was read by the merge two batches ago
        ps, pi = peer.begin(nq, k)
        self.index.topk_raw(q, k, mode, ps, pi)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(st["side"]):
            st["side"].wait_event(ready)
            scores, ids = peer.exchange_merge(nq, k)
            done = torch.cuda.Event()
            done.record(st["side"])
        st["done"] = st["done"][-1:] + [done]
        return PendingTopk(scores, ids, done)

    def _drain_async(self, device) -> None:
        st = self.__dict__.get("_async")
        if st and st["done"]:
            main = torch.cuda.current_stream(device)
            for ev in st["done"]:
                main.wait_event(ev)
            st["done"] = []

    def topk(self, q: torch.Tensor, k: int, mode: str = "bf16") -> Tuple[torch.Tensor, torch.Tensor]:
        if self.world == 1:
            return self.local_topk(q, k, mode)
        self._drain_async(q.device)  # exchanges still in flight on the side stream use the same buffer halves
        nq = int(q.shape[0])
        peer = getattr(self, "peer", None)
        if peer is not None and (nq * k) % 4 == 0 and nq * k * 12 <= peer.max_record_bytes:
            # one library call: local top-k, its slab reduction storing into every peer's buffer, wait, merge
            if os.environ.get("AHR_EXCHANGE_FUSED", "1") != "0":
                return peer.topk_sharded(self.index, q, k, mode)
            ps, pi = peer.begin(nq, k)
            self.index.topk_raw(q, k, mode, ps, pi)
            return peer.exchange_merge(nq, k)
        if self.merge_packed is not None and (nq * k * 4) % 8 == 0:
            # the local top-k writes straight into this rank's record of the exchange buffer
            rec = nq * k * 12
            mine = torch.empty(rec, dtype=torch.uint8, device=q.device)
            out = (mine[: nq * k * 4].view(torch.float32).view(nq, k), mine[nq * k * 4:].view(torch.int64).view(nq, k))
            self.local_topk(q, k, mode, out=out)
            gathered = torch.empty(self.world * rec, dtype=torch.uint8, device=q.device)
            dist.all_gather_into_tensor(gathered, mine, group=self.group)  # the single exchange of the path
            return self.merge_packed(gathered, self.world, nq, k)
        scores, ids = self.local_topk(q, k, mode)
        mine = pack_candidates(scores, ids)
        gathered = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)  # the single exchange of the path
        all_s, all_i = unpack_candidates(gathered, self.world, nq, k)
        return self.merge(all_s, all_i)


class PendingTopk:
    """Result of ``ShardedIndex.topk_async``: tensors that become valid on the caller's stream after ``result()``."""

    def __init__(self, scores: torch.Tensor, ids: torch.Tensor, done) -> None:
        self.scores, self.ids, self.done = scores, ids, done

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.done is not None:
            cur = torch.cuda.current_stream(self.scores.device)
            cur.wait_event(self.done)
            self.scores.record_stream(cur)  # allocated on the side stream, consumed here
            self.ids.record_stream(cur)
            self.done = None
        return self.scores, self.ids


class ShardedGraph:
    """Frontier expansion over a CSR sharded by node range (SURVEY section 8e option B; csrc/graph_shard.cu).

    Rank r owns the adjacency rows of the dense ids ``[bounds[r], bounds[r + 1])``; the batch inputs are replicated.
    ``expand`` = stage 1 on the owners of the input nodes, all-gather, stage 2 on the owners of the hyperedges (the 2-hop
    part of ``expand_related``), all-gather, merge -- two small collectives per call (the path's real exchange steps),
    results bit-identical to ``GraphCSR.expand`` on the whole graph, on every rank.

    ``stage1(ops, ids, offs, max_ids, limit) -> records``, ``stage2(world, bounds, ops, ids, offs, max_ids, limit, g1) ->
    records`` and ``merge(world, bounds, ops, ids, offs, max_ids, limit, g1, g2) -> (ids, counts)`` are injectable so the
    plumbing can be tested on CPU with gloo; in production they are the methods of ``engine.GraphShard``."""

    def __init__(self, stage1: Callable, stage2: Callable, merge: Callable, bounds, group: Optional[dist.ProcessGroup] = None,
                 anchor_stage1: Optional[Callable] = None, anchor_stage2: Optional[Callable] = None,
                 anchor_merge: Optional[Callable] = None) -> None:
        self.stage1, self.stage2, self.merge = stage1, stage2, merge
        self.anchor_stage1, self.anchor_stage2, self.anchor_merge = anchor_stage1, anchor_stage2, anchor_merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = bounds  # int64 [world + 1] on the compute device

    @classmethod
    def from_arrays(cls, arrays, device, group: Optional[dist.ProcessGroup] = None, balance: str = "edges") -> "ShardedGraph":
        """``arrays``: the full-graph dict of ``csr.csr_from_networkx`` (every rank loads the graph file, only its node
        range goes to its GPU).  ``balance``: "nodes" = equal node counts; "edges" = contiguous node ranges of equal CSR
        volume (a hierarchy keeps most of its adjacency in a narrow id range: on c5 the 1.8M L1 summaries hold 93 % of the
        entries), so that every GPU holds ~1/world of the bytes."""
        from . import engine
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        n = int(arrays["n_nodes"])
        b = node_range_bounds(arrays, world, balance)
        shard = engine.GraphShard(arrays, b[rank], b[rank + 1], device)
        obj = cls(shard.stage1, shard.stage2, shard.merge, torch.tensor(b, dtype=torch.int64, device=shard.device), group,
                  shard.anchor_stage1, shard.anchor_stage2, shard.anchor_merge)
        obj.shard = shard
        return obj

    def _gather(self, rec: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return rec.reshape(1, *rec.shape)
        flat = torch.empty(self.world * rec.numel(), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(flat, rec.contiguous().reshape(-1), group=self.group)  # rank-major concatenation
        return flat.reshape((self.world,) + tuple(rec.shape))

    def expand(self, ops: torch.Tensor, in_ids: torch.Tensor, in_offsets: torch.Tensor, limit: int = 10, max_ids: int = 4
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """ops int32 [n_env] (0 parents / 1 children / 2 related); inputs in CSR form (int32), at most ``max_ids`` per
        environment (the gym actions pass 1-3, gym_env.py:94-100).  -> (ids [n_env, limit] padded with -1, counts)."""
        g1 = self._gather(self.stage1(ops, in_ids, in_offsets, max_ids, limit))
        g2 = self._gather(self.stage2(self.world, self.bounds, ops, in_ids, in_offsets, max_ids, limit, g1))
        return self.merge(self.world, self.bounds, ops, in_ids, in_offsets, max_ids, limit, g1, g2)

    def anchor_rerank(self, seed_nodes: torch.Tensor, seed_scores: torch.Tensor, params, top_k: int, want_cluster: bool = True):
        """``GraphCSR.anchor_rerank`` (hierarchical_graph.py:719-789) over the sharded graph: seeds (dense node ids int32
        [nq, n_seeds], e.g. the merged top-k of ``ShardedIndex``, identical on every rank) replicated; stage 1 on the owners
        of the seeds, all-gather, stage 2 on the owners of the emitted neighbours, all-gather, the rerank kernel over the
        gathered records.  Same dict of tensors, bit-identical to the unsharded call, on every rank."""
        g1 = self._gather(self.anchor_stage1(seed_nodes, params))
        g2 = self._gather(self.anchor_stage2(self.world, self.bounds, seed_nodes, params, g1))
        return self.anchor_merge(self.world, self.bounds, seed_nodes, seed_scores, params, top_k, g1, g2, want_cluster)


class ShardedGraphCSR:
    """``engine.GraphCSR``'s call surface over a ``ShardedGraph`` -- what ``HierarchicalGraph`` / ``GraphEnvironment`` hold as
    their device graph once ``HierarchicalGraph.enable_sharding(shard_graph=True)`` is on.  Every method is a collective:
    all ranks call it with the same (replicated) batch and all get the whole answer."""

    def __init__(self, sg: ShardedGraph) -> None:
        self.sg = sg
        self.device = sg.shard.device
        self.n_nodes = sg.shard.n_nodes

    def nbytes(self) -> int:
        return self.sg.shard.nbytes()

    def anchor_rerank(self, seeds: torch.Tensor, seed_scores: torch.Tensor, params, top_k: int, want_cluster: bool = True,
                      row_to_node: Optional[torch.Tensor] = None):
        if row_to_node is not None:  # index rows -> dense node ids (-1: padding or a row that is no node any more)
            rows = seeds.to(torch.int64)
            ok = (rows >= 0) & (rows < row_to_node.numel())
            seeds = torch.where(ok, row_to_node[rows.clamp(0, max(0, row_to_node.numel() - 1))], torch.full_like(rows, -1, dtype=torch.int32))
        return self.sg.anchor_rerank(seeds.to(torch.int32).contiguous(), seed_scores.contiguous(), params, int(top_k), want_cluster)

    def expand(self, ops: torch.Tensor, in_ids: torch.Tensor, in_offsets: torch.Tensor, limit: int = 10):
        lens = in_offsets[1:] - in_offsets[:-1]
        max_ids = max(1, int(lens.max().item())) if lens.numel() else 1
        return self.sg.expand(ops, in_ids, in_offsets, limit=int(limit), max_ids=max_ids)

    def lca_closure(self, *a, **k):
        raise NotImplementedError("expand_to_lca runs on a replicated graph: enable_sharding(shard_graph=False)")
