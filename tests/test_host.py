"""CPU tests of the host-side logic: CSR export, vector store, graph persistence, sharding helpers, featurizer."""
import copy
import json
import os

import networkx as nx
import numpy as np
import pytest
import torch

from ah_rag_b200 import csr, synth
from ah_rag_b200.dist import pack_candidates, shard_bounds, unpack_candidates
from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
from ah_rag_b200.graph.vector_store import VectorStore
from oracle import ref_search
from tests import golden_util as gu

SLOTS = [("participates_in", "out"), ("participates_in", "in"), ("belongs_to", "out"), ("belongs_to", "in"),
         ("related_to", "out"), ("related_to", "in")]


@pytest.mark.parametrize("name", gu.fixture_names())
def test_csr_export_matches_loaded_adjacency_order(name):
    fx = gu.load(name)
    og = ref_search.OracleGraph(copy.deepcopy(fx["node_link"]))
    a = csr.csr_from_networkx(og.G)
    assert a["node_ids"] == list(og.G.nodes)
    for s, (t, d) in enumerate(SLOTS):
        rp, ci = a["row_ptr"][s], a["col_idx"][s]
        assert rp[0] == 0 and rp[-1] == len(ci)
        for i, nid in enumerate(a["node_ids"]):
            assert [a["node_ids"][j] for j in ci[rp[i]:rp[i + 1]]] == og.neighbors(nid, d, t)
    for i, nid in enumerate(a["node_ids"]):
        d = og.G.nodes[nid]
        j = ref_search.judge_overall(d)
        assert (np.isnan(a["judge"][i]) and j is None) or a["judge"][i] == j
        if j is not None:
            assert a["judge_term"][i] == 1.0 / (1.0 + np.exp(-((j or 0.0) / 10.0)))
        c = d.get("confidence")
        assert (np.isnan(a["conf"][i]) and c is None) or a["conf"][i] == c


def test_csr_from_edges_equals_networkx_path():
    arr = synth.make_csr_graph(300, 12, 3, 80, seed=4)
    G = nx.DiGraph()
    names = {0: "entity", 1: "summary", 2: "hyperedge"}
    for i in range(int(arr["n_nodes"])):
        G.add_node(i, node_type=names[int(arr["node_type"][i])])
    et = {0: "participates_in", 1: "belongs_to", 2: "related_to"}
    seen = set()
    keep = []
    for e, (s, d, t) in enumerate(zip(arr["src"], arr["dst"], arr["etype"])):
        if (int(s), int(d)) in seen:
            continue
        seen.add((int(s), int(d)))
        keep.append(e)
        G.add_edge(int(s), int(d), edge_type=et[int(t)])
    keep = np.asarray(keep)
    a = csr.csr_from_networkx(G)
    b = csr.csr_from_edges(int(arr["n_nodes"]), arr["src"][keep], arr["dst"][keep], arr["etype"][keep], arr["node_type"])
    for s in range(6):
        assert np.array_equal(a["row_ptr"][s], b["row_ptr"][s]) and np.array_equal(a["col_idx"][s], b["col_idx"][s])


def test_save_load_roundtrip_and_known_answers(tmp_path):
    hg = HierarchicalGraph()
    assert hg._graph_snapshot_hash() == gu.known_answers()["empty_graph_hash"]
    assert HierarchicalGraph.make_entity_id("Tim Burton") == "ent:4983994569"
    assert hg.G.graph["search_params"] == gu.known_answers()["default_search_params"]
    e = hg.add_entity("Tim Burton", "director", "PERSON")
    s = hg.add_summary(3, title="Film", confidence=7.5, judge_scores={"overall": 8.0})
    h = hg.add_hyperedge("u1", "directed", "directed", 0.9)
    hg.add_belongs_to(e, s, 0.8)
    hg.add_participation(e, h)
    hg.save(str(tmp_path / "g"))
    with open(tmp_path / "g" / "structure.json") as f:
        assert "links" in json.load(f)  # the key load() (and the reference's load) reads
    hg2 = HierarchicalGraph.load(str(tmp_path / "g"))
    assert list(hg2.G.nodes) == list(hg.G.nodes) and hg2.name_to_entity_id["Tim Burton"] == e
    assert hg2.topic_to_summary_id[3] == s and hg2.get_belongs_to(e) == [s] and hg2.get_summary_members(s) == [e]
    assert hg2._graph_snapshot_hash() == hg._graph_snapshot_hash()


def test_vector_store_upsert_semantics(tmp_path):
    db = str(tmp_path / "db")
    st = VectorStore.open(db)
    st.upsert(["a", "b"], ["da", "db"], np.eye(2, 4, dtype=np.float32), [{"i": 0}, {"i": 1}])
    st.upsert(["b", "c"], ["db2", "dc"], np.ones((2, 4), dtype=np.float32), [{"i": 11}, {"i": 2}])
    st.save()
    st2 = VectorStore.open(db)
    assert st2.ids == ["a", "b", "c"] and st2.docs[1] == "db2" and st2.metas[1] == {"i": 11}
    assert np.array_equal(st2.rows[1], np.ones(4, np.float32)) and st2.rows.shape == (3, 4)
    assert VectorStore.stamp(db) is not None
    assert VectorStore.open(db, reset=True).ids == []


def test_build_vector_index_selection_and_incremental(tmp_path):
    fx = gu.load("g_small_s3")
    gdir, db = str(tmp_path / "graph"), str(tmp_path / "db")
    synth.write_graph_dir(gdir, copy.deepcopy(fx["node_link"]))
    hg = HierarchicalGraph.load(gdir)
    calls = []

    def emb(texts):
        calls.append(len(texts))
        return np.tile(np.arange(8, dtype=np.float32), (len(texts), 1))

    hg.set_embedder(emb)
    hg.build_vector_index(db, layers={0, 1, 2}, reset=True)
    assert VectorStore.open(db).ids == fx["index_ids"]  # entities + all summaries in G.nodes order, no hyperedges
    hg.build_vector_index(db, layers={0, 1, 2})  # nothing changed -> nothing re-embedded (content-hash upsert)
    assert calls == [len(fx["index_ids"])]
    some = fx["index_ids"][0]
    hg.G.nodes[some]["description"] = "changed"
    hg.build_vector_index(db, layers={0, 1, 2})
    assert calls == [len(fx["index_ids"]), 1]


def test_shard_bounds_and_candidate_packing():
    for n, w in ((10, 3), (10_000_000, 8), (7, 8), (0, 4)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    s = torch.randn(5, 7)
    i = torch.randint(0, 1 << 40, (5, 7))
    buf = torch.stack([pack_candidates(s, i), pack_candidates(s + 1, i + 1)])
    us, ui = unpack_candidates(buf, 2, 5, 7)
    assert torch.equal(us[0], s) and torch.equal(ui[1], i + 1)


def test_featurizer_and_reward_match_oracle():
    from ah_rag_b200.agent.featurizer import featurize_observation
    from ah_rag_b200.agent.reward import step_reward
    fx = gu.load("g_small_s3")
    for trace in fx["env"]:
        prev = None
        for st in trace:
            v, aux = featurize_observation(st["obs"])
            assert v.shape == (84,) and np.array_equal(v, ref_search.featurize(st["obs"]))
            assert step_reward(prev, st["obs"]) == ref_search.step_reward(prev, st["obs"])
            prev = st["obs"]


# ---------------------------------------------------------------------------------------------
# on-disk -> device loader (vectorised node-link export + binary snapshot)
# ---------------------------------------------------------------------------------------------
def _same_arrays(a, b):
    assert a["n_nodes"] == b["n_nodes"] and a["node_ids"] == b["node_ids"]
    for s in range(6):
        assert np.array_equal(a["row_ptr"][s], b["row_ptr"][s]), s
        assert np.array_equal(a["col_idx"][s], b["col_idx"][s]), s
    for k in ("node_type", "level", "judge", "judge_term", "conf"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_csr_from_node_link_equals_networkx_walk():
    """The vectorised export of structure.json must equal the walk over the graph NetworkX loads from it, including
    what DiGraph does with repeated links (slot of the first, edge_type of the last that has the key), links without
    an edge_type, unknown edge types, links naming nodes absent from "nodes", and repeated node entries."""
    import copy
    import random
    import networkx as nx
    from ah_rag_b200 import csr, synth
    from tests import golden_util as gu
    for name in gu.fixture_names():
        nl = gu.load(name)["node_link"]
        _same_arrays(csr.csr_from_node_link(nl),
                     csr.csr_from_networkx(nx.node_link_graph(copy.deepcopy(nl), directed=True, edges="links")))
    rng = random.Random(3)
    for trial in range(12):
        nl = synth.make_graph(seed=100 + trial, n_entities=30, n_l1=4, n_l2=2, n_hyperedges=10)
        links, ids = nl["links"], [d["id"] for d in nl["nodes"]]
        extra = []
        for _ in range(40):
            l, r = dict(rng.choice(links)), rng.random()
            if r < 0.3:
                l["edge_type"] = rng.choice(["belongs_to", "related_to", "participates_in", "weird", None])
            elif r < 0.5:
                l.pop("edge_type", None)
            elif r < 0.7:
                l["source"], l["target"] = l["target"], l["source"]
            elif r < 0.85:
                l["target"] = "ghost:%d" % rng.randrange(5)
            else:
                l["source"], l["target"] = "ghost:%d" % rng.randrange(5), rng.choice(ids)
            extra.append(l)
        for p, l in zip(reversed(sorted(rng.sample(range(len(links) + 1), len(extra)))), reversed(extra)):
            links.insert(p, l)
        nl["nodes"].append({"id": ids[3], "confidence": 7.5, "level": 2})
        _same_arrays(csr.csr_from_node_link(nl),
                     csr.csr_from_networkx(nx.node_link_graph(copy.deepcopy(nl), directed=True, edges="links")))


def test_device_snapshot_round_trip_and_invalidation(tmp_path):
    import json
    from ah_rag_b200 import csr, synth
    from ah_rag_b200.graph import snapshot
    from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
    d = str(tmp_path / "graph")
    nl = synth.make_graph(seed=4, n_entities=120, n_l1=6, n_l2=2, n_hyperedges=40)
    synth.write_graph_dir(d, nl)
    hg = HierarchicalGraph.load(d)
    want = csr.csr_from_networkx(hg.G)
    _same_arrays(hg._export_arrays(), want)            # built from the document, snapshot written
    assert os.path.exists(snapshot.path_in(d))
    hg2 = HierarchicalGraph.load(d)
    hg2._source["doc"] = None                           # force the snapshot route
    _same_arrays(hg2._export_arrays(), want)
    # a changed structure.json makes the snapshot stale: ignored, rebuilt
    with open(os.path.join(d, "structure.json"), "r", encoding="utf-8") as f:
        doc = json.load(f)
    doc["links"] = doc["links"][:-3]
    with open(os.path.join(d, "structure.json"), "w", encoding="utf-8") as f:
        json.dump(doc, f)
    hg3 = HierarchicalGraph.load(d)
    assert snapshot.load(d, hg3._source["key"]) is None
    _same_arrays(hg3._export_arrays(), csr.csr_from_networkx(hg3.G))
    assert snapshot.load(d, hg3._source["key"]) is not None
    # an edited graph is no longer its document: NetworkX walk
    hg3.add_entity("Somebody New")
    assert hg3._source is None
    _same_arrays(hg3._export_arrays(), csr.csr_from_networkx(hg3.G))
    # a corrupt snapshot is ignored, not an error
    with open(snapshot.path_in(d), "wb") as f:
        f.write(b"not a zip")
    hg4 = HierarchicalGraph.load(d)
    _same_arrays(hg4._export_arrays(), csr.csr_from_networkx(hg4.G))


def _pairwise_np_dot(mat):
    """fp32 ``np.dot`` per pair, i.e. what the reference evaluates (semantic_aggregator.py:634)."""
    t = mat.shape[0]
    out = np.zeros((t, t), dtype=np.float32)
    for i in range(t):
        for j in range(t):
            out[i, j] = np.dot(mat[i], mat[j])
    return out


def test_l1_relations_host_logic_matches_reference_golden(tmp_path):
    """generate_l1_relations against the edges / neighbour lists the unmodified reference produced
    (tests/golden/l1_relations.json).  The cosine matrix is injected (fp32 np.dot per pair, the reference's own
    arithmetic), so everything else -- overlap, Jaccard, rules, rounding, ordering, top-k pruning, neighbours, the two
    artifact files -- must be identical."""
    from ah_rag_b200.aggregate.l1_relations import generate_l1_relations
    with open(os.path.join(gu.GOLDEN_DIR, "l1_relations.json"), "r", encoding="utf-8") as f:
        cases = json.load(f)
    assert sum(len(c["edges"]) for c in cases) > 1000
    for c in cases:
        nodes = copy.deepcopy(c["l1_nodes"])
        art = str(tmp_path / "artifacts")
        edges = generate_l1_relations(nodes, cosine_fn=_pairwise_np_dot, artifact_dir=art, **c["kwargs"])
        assert json.dumps(edges) == json.dumps(c["edges"]), (c["seed"], c["kwargs"])
        assert [n["neighbors"] for n in nodes] == c["neighbors"]
        with open(os.path.join(art, "l1_edges.json"), "r", encoding="utf-8") as f:
            assert json.load(f) == c["edges"]
    assert generate_l1_relations([], artifact_dir=None) == []


def test_build_from_artifacts_matches_reference_golden(tmp_path):
    """The data format in front of the path: artifacts/*.json -> graph (hierarchical_graph.py:241-421).  Node order,
    attributes, link order and the two name maps must equal what the unmodified reference built from the same files."""
    import networkx as nx
    with open(os.path.join(gu.GOLDEN_DIR, "artifacts_graph.json"), "r", encoding="utf-8") as f:
        cases = json.load(f)
    assert cases
    for c in cases:
        d = str(tmp_path / f"art_{c['seed']}_{c['variant']}")
        synth.write_artifacts(d, c["seed"], c["variant"])
        hg = HierarchicalGraph()
        hg.build_from_artifacts(d)
        got = json.loads(json.dumps(nx.node_link_data(hg.G, edges="links"), default=str))
        assert gu.canon(got) == gu.canon(c["node_link"]), (c["seed"], c["variant"])
        assert [n["id"] for n in got["nodes"]] == [n["id"] for n in c["node_link"]["nodes"]]
        assert [(l["source"], l["target"]) for l in got["links"]] == [(l["source"], l["target"]) for l in c["node_link"]["links"]]
        assert hg.name_to_entity_id == c["name_to_entity_id"]
        assert {str(k): v for k, v in hg.topic_to_summary_id.items()} == c["topic_to_summary_id"]
        assert hg.G.graph["dirty"] is True
