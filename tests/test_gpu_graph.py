"""GPU parity tests of hop expansion + rerank (K2) through the C ABI, against the oracle / reference goldens."""
import copy

import networkx as nx
import numpy as np
import pytest
import torch

from ah_rag_b200 import csr as csr_mod
from ah_rag_b200 import synth
from oracle import exact, ref_search
from tests import golden_util as gu
from tests.graph_util import random_csr as _random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from ah_rag_b200 import engine as e
    return e


def _setup(engine, node_link):
    og = ref_search.OracleGraph(copy.deepcopy(node_link))
    arrays = csr_mod.csr_from_networkx(og.G)
    return og, arrays, engine.GraphCSR(arrays)


def _params(engine, p):
    return engine.SearchParams(alpha=p["alpha"], beta=p["beta"], gamma=p["gamma"], delta=p["delta"],
                               judge_overall_min=p["judge_overall_min"], confidence_min=p["confidence_min"],
                               type_filter=p["type_filter"], member_top_m=p["member_top_m"],
                               layer_boost=p["layer_boost"])


@pytest.mark.parametrize("name", gu.fixture_names())
def test_anchor_rerank_matches_reference_golden(engine, name):
    """Seeds taken from the golden file (outputs of the real reference); expansion/rerank on the GPU."""
    fx = gu.load(name)
    og, arrays, g = _setup(engine, fx["node_link"])
    dense = arrays["dense"]
    row_of = {nid: r for r, nid in enumerate(fx["index_ids"])}
    S = exact.canonical_scores(fx["index_rows"], fx["query_vecs"])
    for case in fx["search"]:
        kw = dict(case["kwargs"])
        top_k = kw.pop("top_k")
        prm = ref_search.resolve_params(og, **{"member_top_m": 5, **kw})
        seeds = case["out"]["seeds"]
        ns = max(1, len(seeds))
        sn = np.full((1, ns), -1, np.int32)
        ss = np.zeros((1, ns), np.float32)
        for r, s in enumerate(seeds):
            sn[0, r] = dense[s["node_id"]]
            ss[0, r] = S[case["query"], row_of[s["node_id"]]]
        out = g.anchor_rerank(torch.from_numpy(sn).cuda(), torch.from_numpy(ss).cuda(), _params(engine, prm), top_k)
        out = {k: v.cpu().numpy() for k, v in out.items()}
        want = case["out"]
        # seeds: semantic = 1 - fl32(1 - score)
        assert [float(x) for x in out["seed_semantic"][0][:len(seeds)]] == [s["semantic"] for s in seeds]
        # expanded map: order, unrounded semantic, node_type as the reference reports it
        ne = int(out["exp_count"][0])
        got_exp = []
        for j in range(ne):
            nid = arrays["node_ids"][out["exp_nodes"][0, j]]
            kind = int(out["exp_kind"][0, j])
            nt = og.G.nodes[nid].get("node_type") if kind == 1 else ("summary" if kind == 2 else "entity")
            got_exp.append({"node_id": nid, "semantic": float(out["exp_semantic"][0, j]), "node_type": nt})
        assert got_exp == want["expanded"], (name, case["query"], case["kwargs"])
        nr = int(out["count"][0])
        got = [(arrays["node_ids"][out["nodes"][0, j]], float(out["semantic"][0, j]), float(out["score"][0, j]))
               for j in range(nr)]
        ref = [(r["node_id"], r["semantic"], r["score"]) for r in want["reranked"]]
        assert got == ref, (name, case["query"], case["kwargs"])


@pytest.mark.parametrize("name", gu.fixture_names())
def test_expand_ops_match_reference_golden(engine, name):
    fx = gu.load(name)
    og, arrays, g = _setup(engine, fx["node_link"])
    dense = arrays["dense"]
    opcode = {"expand_parents": 0, "expand_children": 1, "expand_related": 2}
    calls = [st for tr in fx["env"] for st in tr if st["call"] in opcode]
    for limit in (10, 3):
        sel = [c for c in calls if c["args"]["limit"] == limit]
        ops = torch.tensor([opcode[c["call"]] for c in sel], dtype=torch.int32).cuda()
        ids, offs = [], [0]
        for c in sel:
            ids += [dense.get(n, -1) for n in c["args"]["node_ids"]]
            offs.append(len(ids))
        out_ids, out_cnt = g.expand(ops, torch.tensor(ids, dtype=torch.int32).cuda(),
                                    torch.tensor(offs, dtype=torch.int32).cuda(), limit)
        out_ids, out_cnt = out_ids.cpu().numpy(), out_cnt.cpu().numpy()
        for e, c in enumerate(sel):
            got = [arrays["node_ids"][i] for i in out_ids[e, :out_cnt[e]]]
            want = [n["node_id"] for n in c["obs"]["selection"]]
            assert got == want, (name, c["call"], c["args"])
            assert (out_ids[e, out_cnt[e]:] == -1).all()


@pytest.mark.parametrize("seed", [101, 102])
def test_batched_random_graph_vs_oracle(engine, seed):
    """Bigger random graphs, many queries / envs per launch, large fan-out and limits."""
    rng = np.random.default_rng(seed)
    nl = synth.make_graph(seed=seed, n_entities=600, n_l1=12, n_l2=3, n_hyperedges=300, related_density=0.4)
    og, arrays, g = _setup(engine, nl)
    nodes = arrays["node_ids"]
    n = len(nodes)
    # anchor + rerank: 64 queries x 50 random seeds with random scores
    nq, ns, top_k, m = 64, 50, 50, 7
    sn = np.stack([rng.permutation(n)[:ns] for _ in range(nq)]).astype(np.int32)
    ss = np.sort(rng.uniform(0.05, 0.95, size=(nq, ns)).astype(np.float32), axis=1)[:, ::-1].copy()
    prm = ref_search.resolve_params(og, member_top_m=m)
    out = g.anchor_rerank(torch.from_numpy(sn).cuda(), torch.from_numpy(ss).cuda(), _params(engine, prm), top_k)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for q in range(nq):
        seeds = ref_search.seeds_from_topk([nodes[i] for i in sn[q]], ss[q])
        want = ref_search.anchor_rerank(og, seeds, top_k, prm)
        got = [(nodes[out["nodes"][q, j]], float(out["semantic"][q, j]), float(out["score"][q, j]))
               for j in range(out["count"][q])]
        assert got == [(r["node_id"], r["semantic"], r["score"]) for r in want["reranked"]]
        assert [nodes[i] for i in out["exp_nodes"][q, :out["exp_count"][q]]] == [e["node_id"] for e in want["expanded"]]
    # expansions: 1024 envs, 1-3 inputs, mixed ops, limits 10 and 200
    for limit in (10, 200):
        n_env = 1024
        ops = rng.integers(0, 3, size=n_env).astype(np.int32)
        ids, offs = [], [0]
        for e in range(n_env):
            ids += list(rng.integers(0, n, size=int(rng.integers(0, 4))))
            offs.append(len(ids))
        out_ids, out_cnt = g.expand(torch.from_numpy(ops).cuda(), torch.tensor(ids, dtype=torch.int32).cuda(),
                                    torch.tensor(offs, dtype=torch.int32).cuda(), limit)
        out_ids, out_cnt = out_ids.cpu().numpy(), out_cnt.cpu().numpy()
        names = ["parents", "children", "related"]
        for e in range(n_env):
            want = ref_search.expand(og, names[ops[e]], [nodes[i] for i in ids[offs[e]:offs[e + 1]]], limit)
            assert [nodes[i] for i in out_ids[e, :out_cnt[e]]] == want


@pytest.mark.parametrize("limit", [65, 100, 512, 1024, 2048])
def test_expand_large_limits_edge_parallel_kernel(limit):
    """Limits above 64 take the hash-set kernel (one warp per environment, batches of 32 positions): hub nodes with long
    lists, several input ids, duplicates across segments and batches, the 2-hop related walk with the excluded input node,
    ids outside the graph, empty inputs -- against the array oracle (environment.py:393-480)."""
    from ah_rag_b200 import engine
    from oracle import array_graph
    rng = np.random.default_rng(limit)
    a = _random_csr(rng, n_ent=6000, n_sum=1500, n_hyp=800, hub_deg=4500)
    g = engine.GraphCSR(a)
    n, s0 = a["n_nodes"], 6000
    ops, ids, offs = [], [], [0]
    for e in range(300):
        op = e % 3
        k = int(rng.integers(0, 5))
        pick = [int(x) for x in rng.integers(0, n, size=k)]
        if e % 7 == 0:
            pick = [s0, s0 + 1] + pick            # the hubs first
        if e % 11 == 0:
            pick = pick + [s0, -1, n + 5, s0]     # repeated hub, ids outside the graph
        ops.append(op)
        ids += pick
        offs.append(len(ids))
    oi, oc = g.expand(torch.tensor(ops, dtype=torch.int32).cuda(), torch.tensor(ids, dtype=torch.int32).cuda(),
                      torch.tensor(offs, dtype=torch.int32).cuda(), limit)
    oi, oc = oi.cpu().numpy(), oc.cpu().numpy()
    for e in range(300):
        want = array_graph.expand(a, ops[e], ids[offs[e]:offs[e + 1]], limit)
        assert list(oi[e, :oc[e]]) == want, (limit, e, ops[e])
        assert (oi[e, oc[e]:] == -1).all()


@pytest.mark.parametrize("limit", [65, 300, 2048])
def test_expand_large_limits_rows_with_repeated_neighbours(limit):
    """The C ABI accepts any CSR: rows that repeat a neighbour (impossible in a DiGraph, so never in the reference's data)
    must still yield the FIRST occurrence in concatenation order -- inside one batch of 32 positions (the lanes of one
    claim tie on the same slot), across batches, and across segments."""
    from ah_rag_b200 import engine
    from oracle import array_graph
    rng = np.random.default_rng(900 + limit)
    a = _random_csr(rng, n_ent=3000, n_sum=600, n_hyp=400, hub_deg=2000)
    n = a["n_nodes"]
    for s_ in range(6):  # rewrite every CSR with ~40 % of the entries replaced by a copy of a nearby earlier entry
        rp, ci = a["row_ptr"][s_], a["col_idx"][s_].copy()
        for v in range(n):
            p0, p1 = int(rp[v]), int(rp[v + 1])
            for i in range(p0 + 1, p1):
                if rng.random() < 0.4:
                    ci[i] = ci[int(rng.integers(max(p0, i - 40), i))]
        a["col_idx"][s_] = ci
    g = engine.GraphCSR(a)
    ops, ids, offs = [], [], [0]
    for e in range(240):
        pick = [int(x) for x in rng.integers(0, n, size=int(rng.integers(1, 5)))]
        if e % 5 == 0:
            pick = [3000, 3001] + pick  # the hubs
        ops.append(e % 3)
        ids += pick
        offs.append(len(ids))
    oi, oc = g.expand(torch.tensor(ops, dtype=torch.int32).cuda(), torch.tensor(ids, dtype=torch.int32).cuda(),
                      torch.tensor(offs, dtype=torch.int32).cuda(), limit)
    oi, oc = oi.cpu().numpy(), oc.cpu().numpy()
    for e in range(240):
        assert list(oi[e, :oc[e]]) == array_graph.expand(a, ops[e], ids[offs[e]:offs[e + 1]], limit), (limit, e, ops[e])


@pytest.mark.parametrize("alpha,delta", [(0.6, 0.1), (250.0, 0.1), (3.0e6, -4.0e5), (-0.6, 2000.0)])
def test_anchor_rerank_sort_keys_small_and_large_scores(engine, alpha, delta):
    """The stable sort of the rerank uses one 32-bit key per entry while every rounded score is below 2^20 / 1e4 in
    magnitude and the general (double, position) comparison beyond it: weights that put scores inside the fast range, right
    across its edge (alpha = 250: a few entries above 104.8576), far outside and negative -- same order as the reference's
    `sorted(..., key=score, reverse=True)` (hierarchical_graph.py:781), many ties included (expansions share their seed's
    semantic, so equal rounded scores are common)."""
    rng = np.random.default_rng(int(abs(alpha)) + 7)
    nl = synth.make_graph(seed=77, n_entities=500, n_l1=10, n_l2=3, n_hyperedges=200, related_density=0.3)
    og, arrays, g = _setup(engine, nl)
    nodes = arrays["node_ids"]
    nq, ns, top_k = 48, 40, 60
    sn = np.stack([rng.permutation(len(nodes))[:ns] for _ in range(nq)]).astype(np.int32)
    ss = np.round(np.sort(rng.uniform(0.05, 0.95, size=(nq, ns)), axis=1)[:, ::-1], 2).astype(np.float32)  # coarse: ties
    prm = ref_search.resolve_params(og, member_top_m=5, alpha=alpha, delta=delta)
    out = g.anchor_rerank(torch.from_numpy(sn).cuda(), torch.from_numpy(ss.copy()).cuda(), _params(engine, prm), top_k)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for q in range(nq):
        want = ref_search.anchor_rerank(og, ref_search.seeds_from_topk([nodes[i] for i in sn[q]], ss[q]), top_k, prm)
        got = [(nodes[out["nodes"][q, j]], float(out["semantic"][q, j]), float(out["score"][q, j])) for j in range(out["count"][q])]
        assert got == [(r["node_id"], r["semantic"], r["score"]) for r in want["reranked"]], (alpha, delta, q)
