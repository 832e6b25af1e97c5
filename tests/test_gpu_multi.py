"""Multi-GPU parity of the sharded index (needs >= 2 GPUs on the box; skipped otherwise): node-range shards, the
peer-memory exchange of csrc/exchange.cu (stores into the peers' buffers over NVLink + epoch flags) and the NCCL
all-gather path must both return, on every rank, exactly what one index over the whole corpus returns."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ah_rag_b200 import engine, synth
        from ah_rag_b200.dist import ShardedIndex, shard_bounds
        n, dim, k = 60_000 + 13, 128, 50
        X = torch.from_numpy(synth.unit_rows(n, dim, 31, clustered=8)).to(dev).to(torch.bfloat16).contiguous()
        lo, hi = shard_bounds(n, world, rank)
        full = engine.VectorIndex(X)
        sh = ShardedIndex.from_rows(X[lo:hi].contiguous(), id_base=lo)
        problems = []
        for use_peer in (False, "separate push launch", "push fused into the slab reduction"):
            if use_peer and getattr(sh, "peer", None) is None and not sh.enable_peer_exchange(512, k):
                problems.append("peer exchange not enabled")
                break
            os.environ["AHR_EXCHANGE_FUSED"] = "0" if use_peer == "separate push launch" else "1"
            # several exchanges back to back without host synchronisation: exercises the two buffer halves / epochs
            outs, wants = [], []
            for it, nq in enumerate((64, 300, 8, 512, 128, 4)):
                Q = torch.from_numpy(synth.unit_rows(nq, dim, 40 + it, clustered=8)).to(dev).to(torch.bfloat16).contiguous()
                for mode in ("bf16", "exact"):
                    outs.append(sh.topk(Q, k, mode))
                    wants.append(full.topk(Q, k, mode))
            torch.cuda.synchronize()
            for j, ((s, i), (ws, wi)) in enumerate(zip(outs, wants)):
                if not (torch.equal(i, wi) and torch.equal(s.view(torch.int32), ws.view(torch.int32))):
                    problems.append(f"peer={use_peer} case {j} differs")
        # pipelined form: K1 of batch i + 1 on the main stream while batch i is exchanged and merged on the side stream
        if getattr(sh, "peer", None) is not None:
            os.environ["AHR_EXCHANGE_FUSED"] = "1"
            qs = [torch.from_numpy(synth.unit_rows(nq, dim, 70 + it, clustered=8)).to(dev).to(torch.bfloat16).contiguous()
                  for it, nq in enumerate((64, 300, 300, 8, 512, 128, 128, 128, 4, 256))]
            for rep in range(2):
                pend = [sh.topk_async(Q, k, "bf16") for Q in qs]           # ten batches in flight, no host synchronisation
                tail = sh.topk(qs[1], k, "bf16")                           # the synchronous call drains them first
                got = [p_.result() for p_ in pend]
                torch.cuda.synchronize()
                for j, (Q, (s, i)) in enumerate(zip(qs, got)):
                    ws, wi = full.topk(Q, k, "bf16")
                    if not (torch.equal(i, wi) and torch.equal(s.view(torch.int32), ws.view(torch.int32))):
                        problems.append(f"topk_async rep {rep} batch {j} differs")
                ws, wi = full.topk(qs[1], k, "bf16")
                if not (torch.equal(tail[1], wi) and torch.equal(tail[0].view(torch.int32), ws.view(torch.int32))):
                    problems.append(f"topk after topk_async differs (rep {rep})")
        # the knowledge-graph CSR sharded by node range: two NCCL all-gathers per expand call (csrc/graph_shard.cu)
        from ah_rag_b200.dist import ShardedGraph
        from tests.graph_util import random_csr
        from tests.test_shard_stages import _batch
        rng = np.random.default_rng(5)
        a = random_csr(rng, n_ent=4000, n_sum=400, n_hyp=1200, hub_deg=2000)
        sg = ShardedGraph.from_arrays(a, dev)
        whole = engine.GraphCSR(a, dev)
        ops, ids, offs = _batch(np.random.default_rng(6), a, 1024, 3)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        for limit in (10, 3, 32):
            oi, oc = sg.expand(t(ops), t(ids), t(offs), limit=limit, max_ids=3)
            wi, wc = whole.expand(t(ops), t(ids), t(offs), limit)
            if not (torch.equal(oi, wi) and torch.equal(oc, wc)):
                problems.append(f"sharded graph expand differs at limit {limit}")
        # hop 1 (anchor expansion + rerank) over the same sharded graph
        ns, top_k = 16, 20
        seeds = np.stack([rng.choice(4400, ns, replace=False) for _ in range(256)]).astype(np.int32)
        scores = np.sort(rng.uniform(0.1, 0.9, (256, ns)).astype(np.float32), axis=1)[:, ::-1].copy()
        prm = engine.SearchParams(member_top_m=5)
        got = sg.anchor_rerank(t(seeds), t(scores), prm, top_k)
        want = whole.anchor_rerank(t(seeds), t(scores), prm, top_k)
        cnt = want["count"].cpu().numpy()
        if not torch.equal(got["count"], want["count"]):
            problems.append("sharded anchor rerank: counts differ")
        else:
            for key in ("nodes", "semantic", "score"):
                g_, w_ = got[key].cpu().numpy(), want[key].cpu().numpy()
                if any(not np.array_equal(g_[q_, :cnt[q_]].view(np.uint8), w_[q_, :cnt[q_]].view(np.uint8)) for q_ in range(256)):
                    problems.append(f"sharded anchor rerank: {key} differs")
        if sg.shard.nbytes() > 0.8 * whole.nbytes() and world > 1:
            problems.append("graph shard is not a fraction of the graph")
        q.put((rank, problems))
    except Exception as e:  # report instead of leaving the parent waiting
        q.put((rank, [repr(e)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_topk_peer_exchange_and_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=280) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(r, []) for r in range(world)]


def test_peer_exchange_single_rank():
    """world = 1: the exchange degenerates to a merge of this rank's own slot (API / addressing check on one GPU)."""
    from ah_rag_b200 import engine, synth
    n, dim, nq, k = 9000, 64, 96, 50
    X = torch.from_numpy(synth.unit_rows(n, dim, 51)).cuda().to(torch.bfloat16).contiguous()
    Q = torch.from_numpy(synth.unit_rows(nq, dim, 52)).cuda().to(torch.bfloat16).contiguous()
    idx = engine.VectorIndex(X, id_base=1000)
    ex = engine.PeerExchange(torch.device("cuda", 0), 0, 1, nq * k * 12, lambda h: [h])
    for _ in range(3):  # both halves
        ps, pi = ex.begin(nq, k)
        idx.topk_raw(Q, k, "bf16", ps, pi)
        s, i = ex.exchange_merge(nq, k)
        ws, wi = idx.topk(Q, k, "bf16")
        assert torch.equal(i, wi) and torch.equal(s, ws)
    for mode in ("bf16", "exact", "bf16"):  # the one-call form (ahr_topk_sharded)
        s, i = ex.topk_sharded(idx, Q, k, mode)
        ws, wi = idx.topk(Q, k, mode)
        assert torch.equal(i, wi) and torch.equal(s, ws)
    with pytest.raises(RuntimeError):
        ex.topk_sharded(idx, torch.cat([Q] * 4), k, "bf16")  # refused BEFORE the exchange counter moves
    s, i = ex.topk_sharded(idx, Q, k, "bf16")
    assert torch.equal(i, wi)
    with pytest.raises(RuntimeError):
        ex.begin(nq * 4, k)  # record larger than the buffer
    with pytest.raises(RuntimeError):
        ex.begin(3, 3)  # 16-byte store granularity


def _dropin_worker(rank, world, port, q, tmp):
    import copy
    import json
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from ah_rag_b200 import synth
        from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
        from tests import golden_util as gu
        from tests.test_gpu_dropin import _strip
        problems = []
        for name in gu.fixture_names()[:2]:
            fx = gu.load(name)
            gdir, db = os.path.join(tmp, f"r{rank}_{name}_graph"), os.path.join(tmp, f"r{rank}_{name}_db")
            nl = copy.deepcopy(fx["node_link"])
            nl["graph"]["vector_index"]["db_path"] = db
            synth.write_graph_dir(gdir, nl)
            emb = {nid: fx["index_rows"][r] for r, nid in enumerate(fx["index_ids"])}
            qmap = {qq: fx["query_vecs"][i] for i, qq in enumerate(fx["queries"])}
            for shard_graph in (False, True):
                hg = HierarchicalGraph.load(gdir)
                hg.device = f"cuda:{rank}"
                hg.set_embedder(lambda texts: np.stack([qmap[t] for t in texts]))
                hg.build_vector_index(db, layers={0, 1, 2}, reset=True, embeddings=emb)
                hg.enable_sharding(shard_graph=shard_graph)
                for kwargs in sorted({gu.canon(c["kwargs"]) for c in fx["search"]}):
                    outs = hg.search_batch(fx["query_vecs"], return_cluster=True, **json.loads(kwargs))
                    for case in (c for c in fx["search"] if gu.canon(c["kwargs"]) == kwargs):
                        if gu.canon(_strip(outs[case["query"]])) != gu.canon(case["out"]):
                            problems.append(f"{name} shard_graph={shard_graph} query {case['query']} {kwargs}")
                # GraphEnvironment verbs (semantic_anchor / expand_* / commit) over the same sharded state: reference traces
                from ah_rag_b200.agent.environment import GraphEnvironment
                for trace in fx["env"][:3]:
                    env = GraphEnvironment(graph_dir=gdir, logging_enabled=False, hg=hg)
                    for step in trace:
                        a = step["args"]
                        if step["call"] == "reset":
                            obs, info = env.reset(seed_query=fx["queries"][a["query"]], top_k=a["top_k"])
                            info = {k_: v for k_, v in info.items() if k_ != "time_s"}
                        elif step["call"] == "commit_selection":
                            obs, info = env.commit_selection(a["node_ids"])
                        else:
                            obs, info = getattr(env, step["call"])(a["node_ids"], limit=a["limit"])
                        if gu.canon(obs) != gu.canon(step["obs"]) or gu.canon(info) != gu.canon(step["info"]):
                            problems.append(f"{name} shard_graph={shard_graph} env {step['call']}")
                case = fx["search"][0]  # the single-query API, unchanged signature
                out = hg.search(fx["queries"][case["query"]], return_cluster=True, **case["kwargs"])
                if gu.canon(_strip(out)) != gu.canon(case["out"]):
                    problems.append(f"{name} shard_graph={shard_graph} search()")
        q.put((rank, problems[:5]))
    except Exception as e:  # report instead of leaving the parent waiting
        import traceback
        q.put((rank, [repr(e), traceback.format_exc()[-800:]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2])
def test_dropin_search_with_sharded_index_and_graph(world, tmp_path):
    """HierarchicalGraph.enable_sharding: the reference's search API as a collective over 2 GPUs -- index rows sharded by row
    range (K1 + candidate exchange + merge), graph replicated or sharded by node range -- returns, on every rank, exactly the
    golden outputs of the unmodified reference."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dropin_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=280) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == [(r, []) for r in range(world)]
