"""Dev-container only: the oracle restatement against the UNMODIFIED reference on random graphs."""
import copy
import json

import numpy as np
import pytest

from ah_rag_b200 import synth
from oracle import ref_runner, ref_search, stubs

pytestmark = pytest.mark.skipif(not ref_runner.reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_random_graphs_search_env_gym(seed):
    stubs.reset_store()
    rng = np.random.default_rng(seed)
    g = synth.make_graph(seed=seed, n_entities=int(rng.integers(10, 90)), n_l1=int(rng.integers(1, 9)),
                         n_l2=int(rng.integers(0, 3)), n_hyperedges=int(rng.integers(0, 30)),
                         related_density=float(rng.uniform(0, 0.6)))
    rs = ref_runner.ReferenceSession(g)
    og = ref_search.OracleGraph(copy.deepcopy(g))
    env = ref_search.OracleEnv(og, rs.index_ids, rs.index_rows)
    for qi in range(12):
        q = f"random query {seed}-{qi}"
        kw = dict(top_k=int(rng.integers(1, 12)), member_top_m=int(rng.integers(1, 8)))
        if qi % 3 == 1:
            kw["type_filter"] = ["entity"]
        if qi % 4 == 2:
            kw["judge_overall_min"] = 5.0
        ref = ref_runner.strip_search(rs.search(q, **kw))
        tk = kw.pop("top_k")
        got = env.search(stubs.hash_embed(q), tk, **kw)
        assert json.dumps(ref, sort_keys=True) == json.dumps(ref_runner.strip_search(got), sort_keys=True)
    renv = rs.make_env()
    o1, _ = renv.reset("random query x", top_k=5)
    o2, _ = env.reset(stubs.hash_embed("random query x"), top_k=5)
    assert json.dumps(o1, sort_keys=True) == json.dumps(o2, sort_keys=True)
    nodes = list(og.G.nodes)
    for t in range(30):
        ids = [nodes[int(i)] for i in rng.integers(len(nodes), size=int(rng.integers(1, 4)))]
        op = ["expand_parents", "expand_children", "expand_related"][t % 3]
        lim = int(rng.choice([1, 3, 10, 50]))
        a, ia = getattr(renv, op)(ids, limit=lim)
        b, ib = getattr(env, op)(ids, limit=lim)
        assert json.dumps(a, sort_keys=True) == json.dumps(b, sort_keys=True)
        assert ia == ib


def test_featurizer_and_reward_equal_the_reference_functions():
    """The drop-in featurizer / reward (own structure) against the reference's functions, bit for bit, on the golden
    observations and on random observation dicts (missing keys, None values, observations without "state")."""
    import importlib.util
    import random

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    rf = load("_ref_featurizer", "/root/reference/src/ah_rag/agent/featurizer.py")
    rr = load("_ref_reward", "/root/reference/src/ah_rag/agent/reward.py")
    from ah_rag_b200.agent import featurizer as mf, reward as mr
    from tests import golden_util as gu
    rng = random.Random(1)
    observations = []
    for name in gu.fixture_names():
        fx = gu.load(name)
        observations += [st["obs"] for tr in fx["env"] for st in tr] + [c["obs"] for c in fx["lca"] if "obs" in c]

    def random_obs():
        o = {"step": rng.choice([None, 0, 3, 7.5]), "seeds": [{}] * rng.randrange(4), "selection": []}
        if rng.random() < 0.8:
            o["state"] = {"selection_ids": [str(rng.randrange(6)) for _ in range(rng.randrange(5))],
                          "frontier_ids": [str(rng.randrange(9)) for _ in range(rng.randrange(14))]}
        for _ in range(rng.randrange(13)):
            o["selection"].append({"node_id": rng.choice(["a", "b", None]),
                                   "node_type": rng.choice(["entity", "summary", "hyperedge", None]),
                                   "layer": rng.choice([0, 1, 2, None]), "score": rng.choice([None, 0.1234, 1.0, 0]),
                                   "semantic": rng.choice([None, 0.5, 0.33333]),
                                   "judge_overall": rng.choice([None, 7.5, 0.0]), "confidence": rng.choice([None, 3, 8.25])})
        return o

    observations += [random_obs() for _ in range(300)]
    for o in observations:
        a, ia = rf.featurize_observation(o)
        b, ib = mf.featurize_observation(o)
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32), b.view(np.uint32)) and ia == ib
    assert np.array_equal(mf.featurize_batch(observations[:40]),
                          np.stack([rf.featurize_observation(o)[0] for o in observations[:40]]))
    for _ in range(1000):
        p, c = rng.choice(observations + [None]), rng.choice(observations)
        assert rr.step_reward(p, c) == mr.step_reward(p, c)
    for _ in range(100):
        m = {k: rng.random() for k in rng.sample(["f1", "faithfulness", "answer_relevancy", "contextual_recall", "x"], rng.randrange(5))}
        assert rr.final_reward(m) == mr.final_reward(m)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_l1_relations_against_reference(seed, tmp_path):
    """generate_l1_relations (host logic, cosines injected as the reference's fp32 np.dot) against
    ``SemanticAggregator.generate_l1_relations`` of the unmodified reference, files included."""
    from ah_rag_b200.aggregate.l1_relations import generate_l1_relations
    from tests.test_host import _pairwise_np_dot
    for kw in (dict(), dict(min_overlap=1, min_jaccard=0.05, min_cosine=0.3), dict(top_k=5)):
        a = synth.make_l1_nodes(seed, n_topics=40, dim=12)
        b = copy.deepcopy(a)
        want = ref_runner.reference_l1_relations(a, **kw)
        got = generate_l1_relations(b, cosine_fn=_pairwise_np_dot, artifact_dir=str(tmp_path / "art"), **kw)
        assert json.dumps(want) == json.dumps(got) and json.dumps(a) == json.dumps(b)


def test_build_from_artifacts_against_reference(tmp_path):
    """build_from_artifacts (own structure) against the unmodified reference on 24 synthetic artifact sets."""
    import contextlib
    import io
    import networkx as nx
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph as Mine
    HG = stubs.import_reference(ref_runner.REF_ROOT)[0]
    for seed in range(6):
        for variant in range(4):
            d = str(tmp_path / f"a{seed}_{variant}")
            synth.write_artifacts(d, seed, variant)
            ref = HG()
            with contextlib.redirect_stdout(io.StringIO()):
                ref.build_from_artifacts(d)
            mine = Mine()
            mine.build_from_artifacts(d)
            a, b = nx.node_link_data(ref.G, edges="links"), nx.node_link_data(mine.G, edges="links")
            assert json.dumps(a, sort_keys=True, default=str) == json.dumps(b, sort_keys=True, default=str), (seed, variant)
            assert [x["id"] for x in a["nodes"]] == [x["id"] for x in b["nodes"]]
            assert [(l["source"], l["target"]) for l in a["links"]] == [(l["source"], l["target"]) for l in b["links"]]
            assert ref.name_to_entity_id == mine.name_to_entity_id and ref.topic_to_summary_id == mine.topic_to_summary_id


def test_graph_utility_methods_against_reference(tmp_path):
    """get_siblings / search_by_name_or_title / summaries_with_top_word / validate_* / stats (hierarchical_graph.py:471-553)."""
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph as Mine
    HG = stubs.import_reference(ref_runner.REF_ROOT)[0]
    for seed, cyc in ((41, False), (42, True)):
        g = synth.make_graph(seed=seed, n_entities=60, n_l1=7, n_l2=2, n_hyperedges=25, related_density=0.4)
        if cyc:  # a belongs_to cycle between two summaries, a nameless entity, a hyperedge without description
            sums = [n["id"] for n in g["nodes"] if n.get("node_type") == "summary"]
            g["links"] += [{"source": sums[0], "target": sums[1], "edge_type": "belongs_to"},
                           {"source": sums[1], "target": sums[0], "edge_type": "belongs_to"}]
            g["nodes"] += [{"id": "ent:anon", "node_type": "entity"}, {"id": "hedge:x", "node_type": "hyperedge", "relation_type": "r"},
                           {"id": "sum:none", "node_type": "summary", "top_words": "not json"}]
        d = str(tmp_path / f"g{seed}")
        synth.write_graph_dir(d, copy.deepcopy(g))
        ref, mine = HG.load(d), Mine.load(d)
        assert ref.validate_belongs_to_dag() == mine.validate_belongs_to_dag() == (not cyc)
        assert ref.validate_required_attributes() == mine.validate_required_attributes()
        assert ref.stats() == mine.stats()
        for n in list(ref.G.nodes)[::5]:
            assert set(ref.get_siblings(n)) == set(mine.get_siblings(n))
        for q, lim in (("entity", 20), ("E", 3), ("topic", 0), ("zzz", 5), ("", 2)):
            assert [x[0] for x in ref.search_by_name_or_title(q, lim)] == [x[0] for x in mine.search_by_name_or_title(q, lim)]
        words = set()
        for _, dd in ref.G.nodes(data=True):
            if dd.get("top_words") and isinstance(dd["top_words"], str) and dd["top_words"].startswith("["):
                words.update(json.loads(dd["top_words"]))
        for w in sorted(words)[:6] + ["nope"]:
            for lim in (50, 1):
                assert ref.summaries_with_top_word(w, lim) == mine.summaries_with_top_word(w, lim)
