"""GPU parity of the drop-in Python API (HierarchicalGraph / GraphEnvironment / AHRAGGymEnv / NaiveRAG and their
batched forms) against the golden vectors written by the UNMODIFIED reference."""
import copy
import json
import os

import numpy as np
import pytest

from ah_rag_b200 import synth
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _make_hg(fx, tmp_path, index_dtype="fp32"):
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
    gdir = str(tmp_path / "graph")
    db = str(tmp_path / "vector_db")
    nl = copy.deepcopy(fx["node_link"])
    nl["graph"]["vector_index"]["db_path"] = db
    synth.write_graph_dir(gdir, nl)
    hg = HierarchicalGraph.load(gdir)
    hg.index_dtype = index_dtype
    emb = {nid: fx["index_rows"][r] for r, nid in enumerate(fx["index_ids"])}
    qmap = {q: fx["query_vecs"][i] for i, q in enumerate(fx["queries"])}
    hg.set_embedder(lambda texts: np.stack([qmap[t] for t in texts]))
    hg.build_vector_index(db, layers={0, 1, 2}, reset=True, embeddings=emb)
    assert hg._device_index(db)["store"].ids == fx["index_ids"]  # same selection + order as the reference's index
    return hg, gdir


def _strip(out):
    return {"seeds": [{"node_id": s["node_id"], "semantic": s["semantic"]} for s in out["seeds"]],
            "expanded": [{"node_id": e["node_id"], "semantic": e["semantic"], "node_type": e["node_type"]} for e in out["expanded"]],
            "reranked": out["reranked"]}


@pytest.mark.parametrize("name", gu.fixture_names())
def test_search_single_and_batched(name, tmp_path):
    fx = gu.load(name)
    hg, _ = _make_hg(fx, tmp_path)
    for case in fx["search"]:
        out = hg.search(fx["queries"][case["query"]], return_cluster=True, **case["kwargs"])
        assert gu.canon(_strip(out)) == gu.canon(case["out"]), (name, case["query"], case["kwargs"])
    # batched entry point: every query of the fixture in one launch, per case
    for kwargs in {gu.canon(c["kwargs"]) for c in fx["search"]}:
        kw = json.loads(kwargs)
        outs = hg.search_batch(fx["query_vecs"], return_cluster=True, **kw)
        for case in (c for c in fx["search"] if gu.canon(c["kwargs"]) == kwargs):
            assert gu.canon(_strip(outs[case["query"]])) == gu.canon(case["out"])
    # return_cluster=False returns the reranked list only (test/test_graph_search.py invariants)
    res = hg.search(fx["queries"][0], top_k=3)
    assert len(res) > 0
    for item in res:
        assert item["node_type"] in {"entity", "summary", "hyperedge"} and item["score"] >= 0.0


@pytest.mark.parametrize("name", gu.fixture_names())
def test_environment_traces(name, tmp_path):
    from ah_rag_b200.agent.environment import GraphEnvironment
    fx = gu.load(name)
    hg, gdir = _make_hg(fx, tmp_path)
    for trace in fx["env"]:
        env = GraphEnvironment(graph_dir=gdir, logging_enabled=False, hg=hg)
        for step in trace:
            a = step["args"]
            if step["call"] == "reset":
                obs, info = env.reset(seed_query=fx["queries"][a["query"]], top_k=a["top_k"])
                info = {k: v for k, v in info.items() if k != "time_s"}
                assert info["returned"] == len(obs["selection"]) and len(obs["seeds"]) > 0  # test/test_environment.py
            elif step["call"] == "commit_selection":
                obs, info = env.commit_selection(a["node_ids"])
            else:
                obs, info = getattr(env, step["call"])(a["node_ids"], limit=a["limit"])
            assert gu.canon(obs) == gu.canon(step["obs"]), (name, step["call"], a)
            assert gu.canon(info) == gu.canon(step["info"]), (name, step["call"], a)


@pytest.mark.parametrize("name", gu.fixture_names())
def test_gym_traces_single_and_vectorised(name, tmp_path):
    from ah_rag_b200.agent.gym_env import AHRAGGymEnv, VecAHRAGGymEnv
    fx = gu.load(name)
    hg, gdir = _make_hg(fx, tmp_path)
    for ep in fx["gym"]:
        gym = AHRAGGymEnv(graph_dir=gdir, hg=hg)
        vec, info = gym.reset(fx["queries"][ep["query"]])
        assert vec.shape == (84,) and vec.dtype == np.float32
        assert [float(x) for x in vec] == ep["steps"][0]["vec"] and info["action_mask"] == ep["steps"][0]["mask"]
        for st in ep["steps"][1:]:
            v, r, d, inf = gym.step(st["action"])
            assert [float(x) for x in v] == st["vec"] and r == st["reward"] and d == st["done"]
            assert inf["action_mask"] == st["mask"]
    # vectorised: all episodes side by side, one launch per step (episodes of different length: finished envs idle on 5)
    eps = fx["gym"]
    venv = VecAHRAGGymEnv(hg, len(eps))
    obs, _ = venv.reset(np.stack([fx["query_vecs"][ep["query"]] for ep in eps]))
    for b, ep in enumerate(eps):
        assert [float(x) for x in obs[b]] == ep["steps"][0]["vec"]
    for t in range(1, max(len(ep["steps"]) for ep in eps)):
        acts = [ep["steps"][t]["action"] if t < len(ep["steps"]) else 5 for ep in eps]
        o, r, d, _ = venv.step(acts)
        for b, ep in enumerate(eps):
            if t < len(ep["steps"]):
                st = ep["steps"][t]
                assert [float(x) for x in o[b]] == st["vec"] and r[b] == st["reward"] and bool(d[b]) == st["done"]


def test_naive_rag_retrieval_and_bf16_index(tmp_path):
    from ah_rag_b200.baselines.naive_rag import NaiveRAG
    fx = gu.load("g_dense_s11")
    hg, _ = _make_hg(fx, tmp_path)

    class Gen:
        def generate(self, query, context, cfg):
            return {"query": query, "used": context["used_nodes"], "text": context["context_text"]}

    rag = NaiveRAG(hg, Gen())
    case = next(c for c in fx["search"] if c["kwargs"] == {"top_k": 5})
    out = rag.run(fx["queries"][case["query"]], top_k=5)
    assert out["used"] == [r["node_id"] for r in case["out"]["reranked"]]
    outs = rag.run_batch(fx["queries"], fx["query_vecs"], top_k=5)
    for c in (c for c in fx["search"] if c["kwargs"] == {"top_k": 5}):
        assert outs[c["query"]]["used"] == [r["node_id"] for r in c["out"]["reranked"]]
    # bf16 (tensor-core) index: same seeds up to bf16 rounding -> recall of the reranked ids stays high
    hg.index_dtype = "bf16"
    got = hg.search_batch(fx["query_vecs"], top_k=5)
    hits = tot = 0
    for c in (c for c in fx["search"] if c["kwargs"] == {"top_k": 5}):
        want = {r["node_id"] for r in c["out"]["reranked"]}
        hits += len(want & {r["node_id"] for r in got[c["query"]]})
        tot += len(want)
    assert hits / tot >= 0.9


@pytest.mark.parametrize("name", gu.fixture_names())
def test_expand_to_lca_matches_reference_golden(name, tmp_path):
    """expand_to_lca over the exported CSR against the traces of the unmodified reference
    (environment.py:243-299), including its quirk (descendant sets) and the error for nodes outside the hierarchy."""
    from ah_rag_b200.agent.environment import GraphEnvironment
    fx = gu.load(name)
    gdir = str(tmp_path / "graph")
    synth.write_graph_dir(gdir, copy.deepcopy(fx["node_link"]))
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
    hg = HierarchicalGraph.load(gdir)
    env = GraphEnvironment(graph_dir=gdir, logging_enabled=False, hg=hg)
    assert fx["lca"], "fixture without lca cases"
    for case in fx["lca"]:
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                env.expand_to_lca(list(case["node_ids"]), max_results=case["max_results"])
            assert type(ei.value).__name__ == case["raises"], case
            continue
        obs, info = env.expand_to_lca(list(case["node_ids"]), max_results=case["max_results"])
        info = {k: v for k, v in info.items() if k != "time_s"}
        assert gu.canon(obs) == gu.canon(case["obs"]), case["node_ids"]
        assert gu.canon(info) == gu.canon(case["info"]), case["node_ids"]


def test_attribute_edits_in_place_reach_the_device_graph(tmp_path):
    """the reference reads G on every call.  Editing a node attribute through hg.G.nodes[...] (no add_* call,
    node / edge counts unchanged) must change the next search like it changes the reference's: the cached device graph is
    keyed by an attribute fingerprint for graphs of this size."""
    fx = gu.load("g_small_s3")
    hg, _ = _make_hg(fx, tmp_path)
    q = fx["queries"][0]
    before = hg.search(q, top_k=5)
    target = next(r["node_id"] for r in before if r["node_type"] == "summary")
    hg.G.nodes[target]["judge_scores"] = json.dumps({"overall": 0.0})
    hg.G.nodes[target]["confidence"] = 0.0
    after = hg.search(q, top_k=5)
    got = next((r for r in after if r["node_id"] == target), None)
    old = next(r for r in before if r["node_id"] == target)
    assert got is None or (got["judge_overall"] == 0.0 and got["score"] < old["score"])
    # a request wider than the tensor-core lists on a bf16 index takes the exact scan instead of raising
    hg.index_dtype = "bf16"
    wide = hg.search(q, top_k=200)
    assert len(wide) > 0
