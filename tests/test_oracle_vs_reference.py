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


@pytest.mark.parametrize("seed", [31, 32])
def test_expand_to_lca_against_reference_on_random_graphs(seed, tmp_path):
    """the CSR-based expand_to_lca of the drop-in environment against the unmodified reference
    (environment.py:243-299) on random graphs and random input sets, errors included."""
    from ah_rag_b200.agent.environment import GraphEnvironment
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
    stubs.reset_store()
    rng = np.random.default_rng(seed)
    g = synth.make_graph(seed=seed, n_entities=int(rng.integers(20, 120)), n_l1=int(rng.integers(2, 10)),
                         n_l2=int(rng.integers(0, 4)), n_hyperedges=int(rng.integers(0, 30)),
                         related_density=float(rng.uniform(0, 0.5)))
    rs = ref_runner.ReferenceSession(copy.deepcopy(g))
    renv = rs.make_env()
    gdir = str(tmp_path / "graph")
    synth.write_graph_dir(gdir, copy.deepcopy(g))
    env = GraphEnvironment(graph_dir=gdir, logging_enabled=False, hg=HierarchicalGraph.load(gdir))
    nodes = list(rs.hg.G.nodes)
    summaries = [n for n in nodes if rs.hg.G.nodes[n].get("node_type") == "summary"]
    for t in range(40):
        pool = summaries if t % 2 else nodes
        ids = [pool[int(i)] for i in rng.integers(len(pool), size=int(rng.integers(1, 4)))]
        mr = int(rng.choice([1, 3, 5]))
        try:
            want = renv.expand_to_lca(list(ids), max_results=mr)
        except Exception as e:  # noqa: BLE001 -- nx.ancestors on a node outside the belongs_to subgraph
            with pytest.raises(type(e)):
                env.expand_to_lca(list(ids), max_results=mr)
            continue
        got = env.expand_to_lca(list(ids), max_results=mr)
        strip = lambda i: {k: v for k, v in i.items() if k != "time_s"}
        assert json.dumps(want[0], sort_keys=True) == json.dumps(got[0], sort_keys=True), ids
        assert strip(want[1]) == strip(got[1]), ids
