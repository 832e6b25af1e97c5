"""CPU tests of the oracle itself: pinned against the golden vectors written by the real reference."""
import copy
import hashlib
import json
import random

import numpy as np
import pytest

from oracle import exact, ref_search
from tests import golden_util as gu


def _env(fx):
    og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
    return ref_search.OracleEnv(og, fx["index_ids"], fx["index_rows"])


@pytest.mark.parametrize("name", gu.fixture_names())
def test_search_matches_reference_golden(name):
    fx = gu.load(name)
    env = _env(fx)
    for case in fx["search"]:
        kw = dict(case["kwargs"])
        top_k = kw.pop("top_k")
        out = env.search(fx["query_vecs"][case["query"]], top_k, **kw)
        got = {"seeds": out["seeds"], "expanded": out["expanded"], "reranked": out["reranked"]}
        assert gu.canon(got) == gu.canon(case["out"]), (name, case["query"], case["kwargs"])


@pytest.mark.parametrize("name", gu.fixture_names())
def test_env_traces_match_reference_golden(name):
    fx = gu.load(name)
    for trace in fx["env"]:
        env = _env(fx)
        for step in trace:
            a = step["args"]
            if step["call"] == "reset":
                obs, info = env.reset(fx["query_vecs"][a["query"]], top_k=a["top_k"])
                info = {**info, "query": fx["queries"][a["query"]]}
            elif step["call"] == "commit_selection":
                obs, info = env.commit_selection(a["node_ids"])
            else:
                obs, info = getattr(env, step["call"])(a["node_ids"], limit=a["limit"])
            assert gu.canon(obs) == gu.canon(step["obs"]), (name, step["call"], a)
            assert gu.canon(info) == gu.canon(step["info"]), (name, step["call"], a)


@pytest.mark.parametrize("name", gu.fixture_names())
def test_gym_traces_match_reference_golden(name):
    fx = gu.load(name)
    for ep in fx["gym"]:
        gym = ref_search.OracleGymEnv(_env(fx))
        vec, _ = gym.reset(fx["query_vecs"][ep["query"]])
        assert [float(x) for x in vec] == ep["steps"][0]["vec"]
        for st in ep["steps"][1:]:
            v, r, d, _ = gym.step(st["action"])
            assert [float(x) for x in v] == st["vec"]
            assert r == st["reward"] and d == st["done"]


def test_known_answers_of_reference_tests():
    ka = gu.known_answers()
    # test/test_answer_generator.py:9 -- id scheme of hierarchical_graph.py:56-61
    assert ka["make_entity_id"]["Tim Burton"] == "ent:4983994569"
    assert "ent:" + hashlib.sha1("Tim Burton".encode()).hexdigest()[:10] == "ent:4983994569"
    # graph_datasets/hotpotqa_distractor/meta.json:17 -- sha1 of "[]" (hierarchical_graph.py:220-236)
    assert ka["empty_graph_hash"] == "97d170e1550eee4afc0af065b78cda302a97674c"
    assert hashlib.sha1(json.dumps([]).encode()).hexdigest() == ka["empty_graph_hash"]
    assert ka["default_search_params"]["layer_boost"] == {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0}


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_c_and_numpy_exact_scorers_agree_bitwise(dtype):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((777, 96)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X[100] = X[3]  # exact duplicate rows -> tie broken by lower id
    X[500] = X[3]
    Q = np.concatenate([X[3:4], rng.standard_normal((5, 96)).astype(np.float32)])
    if dtype == "bf16":
        X, Q = exact.f32_to_bf16_bits(X), exact.f32_to_bf16_bits(Q)
    a, b = exact.canonical_scores_np(X, Q), exact.canonical_scores(X, Q)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    s1, i1 = exact.exact_topk_np(X, Q, 20)
    s2, i2 = exact.exact_topk(X, Q, 20)
    assert np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    assert list(i1[0][:3]) == [3, 100, 500]
    # k > n pads with (-inf, -1); ragged / empty inputs
    s3, i3 = exact.exact_topk(X[:5], Q, 8)
    assert (i3[:, 5:] == -1).all() and np.isneginf(s3[:, 5:]).all()
    # sharded scan + merge == single scan (what the multi-GPU path relies on)
    parts = [exact.exact_topk(X[lo:hi], Q, 20, id_base=lo) for lo, hi in ((0, 300), (300, 301), (301, 777))]
    ms, mi = exact.merge_topk([p[0] for p in parts], [p[1] for p in parts], 20)
    assert np.array_equal(mi, i2) and np.array_equal(ms.view(np.uint32), s2.view(np.uint32))
    r = exact.rescore(X, Q, i2)
    assert np.array_equal(r.view(np.uint32), s2.view(np.uint32))
    # row-parallel scan (bench.py's sampled parity check over the full corpus) == the per-query scan
    big = np.concatenate([X] * 12)  # 9324 rows: several 4096-row blocks, duplicates across blocks -> id tie-breaks
    for nt in (1, 3, 8):
        s4, i4 = exact.exact_topk_rowpar(big, Q, 20, id_base=7, n_threads=nt)
        s5, i5 = exact.exact_topk(big, Q, 20, id_base=7)
        assert np.array_equal(i4, i5) and np.array_equal(s4.view(np.uint32), s5.view(np.uint32))
    s6, i6 = exact.exact_topk_rowpar(X[:5], Q, 8, n_threads=4)
    assert np.array_equal(i6, i3) and np.array_equal(s6.view(np.uint32), s3.view(np.uint32))


def test_round4_matches_cpython():
    random.seed(7)
    xs = []
    for t in range(200000):
        m = t % 4
        if m == 0:
            xs.append(random.random() * 2)
        elif m == 1:
            xs.append(round(random.random(), 4) + 0.00005 + random.uniform(-1e-15, 1e-15))
        elif m == 2:
            xs.append(random.randint(0, 30000) / 20000.0)
        else:
            xs.append(random.randint(0, 300000) / 200000.0 * random.choice([1, 10, 100, 1e-3]))
    xs += [0.0, -0.0, 0.00035, 5e-05, 1.00005, 2.5e-5, 1e-300, 123456.78905]
    arr = np.array(xs + [-x for x in xs])
    got = exact.round4_array(arr)
    want = np.array([round(float(x), 4) for x in arr])
    assert np.array_equal(got, want)


def test_bf16_bits_roundtrip_matches_torch():
    import torch
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    bits = exact.f32_to_bf16_bits(x)
    t = torch.from_numpy(x).to(torch.bfloat16)
    assert np.array_equal(t.view(torch.int16).numpy().view(np.uint16), bits)
    assert np.array_equal(exact.bf16_bits_to_f32(bits), t.float().numpy())


@pytest.mark.parametrize("name", gu.fixture_names())
def test_array_oracle_equals_networkx_oracle(name):
    """oracle/array_graph.py (used at c5 scale) == oracle/ref_search.py (pinned to the reference) on the goldens."""
    from ah_rag_b200 import csr
    from oracle import array_graph
    fx = gu.load(name)
    og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
    if og.G.graph["search_params"].get("type_filter") or og.G.graph["search_params"].get("judge_overall_min"):
        pytest.skip("array oracle restates the default parameters only")
    a = csr.csr_from_networkx(og.G)
    nodes = a["node_ids"]
    rng = np.random.default_rng(3)
    prm = ref_search.resolve_params(og, member_top_m=4)
    for _ in range(40):
        sn = rng.permutation(len(nodes))[:6]
        ss = np.sort(rng.uniform(0.1, 0.9, 6).astype(np.float32))[::-1]
        want = ref_search.anchor_rerank(og, ref_search.seeds_from_topk([nodes[i] for i in sn], ss), 7, prm)["reranked"]
        got = array_graph.anchor_rerank(a, sn, ss, 7, 4)
        assert [(nodes[n], s, sc) for n, s, sc in got] == [(r["node_id"], r["semantic"], r["score"]) for r in want]
    for t in range(90):
        ids = [int(i) for i in rng.integers(0, len(nodes), size=int(rng.integers(1, 4)))]
        lim = int(rng.choice([3, 10, 100]))
        want = ref_search.expand(og, ["parents", "children", "related"][t % 3], [nodes[i] for i in ids], lim)
        assert [nodes[i] for i in array_graph.expand(a, t % 3, ids, lim)] == want


def test_cpu_graph_baseline_loops_agree():
    """bench.py's graph-half CPU baseline: the NetworkX form rebuilt from CSR arrays gives the array form's answers
    (same environments, same hops), so the two baseline numbers time the same work."""
    from ah_rag_b200 import csr
    from oracle import array_graph, cpu_graph_baseline
    fx = gu.load("g_small_s3")
    og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
    a = csr.csr_from_networkx(og.G)
    og2 = cpu_graph_baseline.networkx_graph(a)
    assert og2.G.number_of_nodes() == og.G.number_of_nodes() and og2.G.number_of_edges() == og.G.number_of_edges()
    # in-edge order of a DiGraph is global insertion order: re-export so both forms walk the SAME adjacency order (the
    # bench graphs are generated source-major, where the conversion is order-exact)
    a = csr.csr_from_networkx(og2.G)
    assert a["node_ids"] == [str(i) for i in range(len(a["node_ids"]))]
    rng = np.random.default_rng(9)
    prm = ref_search.resolve_params(og2, member_top_m=5)
    n = len(a["node_ids"])
    for t in range(30):
        sn = rng.permutation(n)[:8]
        ss = np.sort(rng.uniform(0.1, 0.9, 8).astype(np.float32))[::-1]
        want = array_graph.anchor_rerank(a, sn, ss, 8, 5)
        got = ref_search.anchor_rerank(og2, ref_search.seeds_from_topk([str(int(v)) for v in sn], ss), 8, prm)["reranked"]
        assert [(int(r["node_id"]), r["semantic"], r["score"]) for r in got] == want
        ids = [int(i) for i in rng.integers(0, n, size=2)]
        assert [int(x) for x in ref_search.expand(og2, ["parents", "children", "related"][t % 3], [str(i) for i in ids], 10)] == \
            array_graph.expand(a, t % 3, ids, 10)
    seeds = np.stack([rng.permutation(n)[:8] for _ in range(4)])
    scores = np.sort(rng.uniform(0.1, 0.9, (4, 8)).astype(np.float32), axis=1)[:, ::-1]
    r1 = cpu_graph_baseline.array_loop(a, seeds, scores, [0, 1, 2, 0], 8, 5, 10, 5.0)
    r2 = cpu_graph_baseline.networkx_loop(a, seeds, scores, [0, 1, 2, 0], 8, 5, 10, 5.0)
    assert r1["envs"] == 4 and r2["envs"] == 4 and r1["envs_per_s"] > 0 and r2["envs_per_s"] > 0
