"""Generate tests/golden/* by running the UNMODIFIED reference (dev container only).

    python -m oracle.make_golden            # rewrites tests/golden/

Every expected value in the fixtures is an output of /root/reference's own code
(HierarchicalGraph.search, GraphEnvironment.*, AHRAGGymEnv.*) executed through ``oracle/stubs.py``;
inputs are ``ah_rag_b200.synth.make_graph`` graphs and hash embeddings.  The similarity seam inside
the reference run is the stub's exact cosine (parity there is unpinned, see oracle/__init__.py).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ah_rag_b200 import synth  # noqa: E402
from oracle import ref_runner, stubs  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

GRAPHS = [
    dict(name="g_small_s3", kw=dict(seed=3)),
    dict(name="g_dense_s11", kw=dict(seed=11, n_entities=120, n_l1=6, n_l2=3, n_hyperedges=60, related_density=0.5)),
    dict(name="g_flat_s5", kw=dict(seed=5, n_entities=40, n_l1=3, n_l2=0, n_hyperedges=5)),
    dict(name="g_filter_s7", kw=dict(seed=7, n_entities=50, n_l1=10, n_l2=2, n_hyperedges=20,
                                     search_params={"judge_overall_min": 3.0, "member_top_m": 3,
                                                    "type_filter": ["summary", "entity"],
                                                    "alpha": 0.5, "beta": 0.3, "gamma": 0.15, "delta": 0.05})),
]

SEARCH_CASES = [
    dict(top_k=5), dict(top_k=3, member_top_m=2), dict(top_k=1), dict(top_k=10, type_filter=["summary"]),
    dict(top_k=8, judge_overall_min=4.0), dict(top_k=8, confidence_min=3.0, alpha=0.3, beta=0.4),
    dict(top_k=50, member_top_m=5), dict(top_k=200, member_top_m=50),
]

GYM_ACTION_SEQS = [[1, 0, 2, 2, 3, 4, 5], [0, 0, 0, 0, 0, 0], [2, 1, 3, 3, 1, 2], [3, 4, 1, 5], [5]]


def write_fixture(name, g, queries, search_cases, extra=None, env_queries=(0, 3), env_top_k=5):
    """Run the UNMODIFIED reference over node-link graph `g` and write tests/golden/<name>.json/.npz: search results,
    environment / expand_to_lca / gym traces.  Returns the ReferenceSession."""
    stubs.reset_store()
    rs = ref_runner.ReferenceSession(g)
    qvecs = np.stack([stubs.hash_embed(q) for q in queries])
    fixture = {"name": name, **(extra or {}), "node_link": g,
               "index_ids": rs.index_ids, "queries": queries, "search": [], "env": [], "gym": [], "lca": []}
    for qi, q in enumerate(queries):
        for case in search_cases:
            out = ref_runner.strip_search(rs.search(q, **case))
            fixture["search"].append({"query": qi, "kwargs": case, "out": out})
    # environment traces: reset then every expand op on several id slices, then commit
    for qi in env_queries:
        env = rs.make_env()
        trace = []
        obs, info = env.reset(seed_query=queries[qi], top_k=env_top_k)
        info = {k: v for k, v in info.items() if k != "time_s"}
        trace.append({"call": "reset", "args": {"query": qi, "top_k": env_top_k}, "obs": obs, "info": info})
        ids = [n["node_id"] for n in obs["selection"]]
        for op in ("expand_parents", "expand_children", "expand_related"):
            for sl in (ids[:1], ids[:2], ids[1:4], ids, ["not-a-node"] + ids[:1]):
                for limit in (10, 3):
                    o, i = getattr(env, op)(list(sl), limit=limit)
                    trace.append({"call": op, "args": {"node_ids": list(sl), "limit": limit}, "obs": o, "info": i})
        o, i = env.commit_selection(ids[:3])
        trace.append({"call": "commit_selection", "args": {"node_ids": ids[:3]}, "obs": o, "info": i})
        o, i = env.expand_related(ids[:2])
        trace.append({"call": "expand_related", "args": {"node_ids": ids[:2], "limit": 10}, "obs": o, "info": i})
        fixture["env"].append(trace)
    # expand_to_lca (environment.py:243-299): summaries, entities, mixed sets, a node without belongs_to edges
    lenv = rs.make_env()
    G = rs.hg.G
    by_type = {t: [n for n, d in G.nodes(data=True) if d.get("node_type") == t] for t in ("entity", "summary", "hyperedge")}
    ents, sums = by_type["entity"], by_type["summary"]
    parent_of = lambda n: [v for _, v, d in G.out_edges(n, data=True) if d.get("edge_type") == "belongs_to"]
    cases = [[s_] for s_ in sums[:3]] + [sums[:2], sums[-2:], ents[:1], ents[:2], [ents[0]] + parent_of(ents[0])[:1],
                                         parent_of(ents[1])[:1] + [ents[1]], [], by_type["hyperedge"][:1], ["not-a-node"]]
    lrng = np.random.default_rng(len(G))
    hier = [n for n in G if any(d.get("edge_type") == "belongs_to" for _, _, d in G.out_edges(n, data=True)) or
            any(d.get("edge_type") == "belongs_to" for _, _, d in G.in_edges(n, data=True))]
    for t in range(24):  # random sets: summaries only (odd t) or any node of the hierarchy
        pool = sums if (t % 2 and sums) else hier
        cases.append([pool[int(i)] for i in lrng.integers(len(pool), size=int(lrng.integers(1, 4)))])
    for ids in cases:
        for mr in (5, 1):
            try:
                o, i = lenv.expand_to_lca(list(ids), max_results=mr)
                i = {k: v for k, v in i.items() if k != "time_s"}
                fixture["lca"].append({"node_ids": list(ids), "max_results": mr, "obs": o, "info": i})
            except Exception as e:  # nx.ancestors on a node outside the belongs_to subgraph
                fixture["lca"].append({"node_ids": list(ids), "max_results": mr, "raises": type(e).__name__})
    for si, seq in enumerate(GYM_ACTION_SEQS):
        gym = rs.make_gym(max_steps=6)
        qi = si % len(queries)
        vec, info = gym.reset(queries[qi])
        steps = [{"vec": [float(x) for x in vec], "mask": info.get("action_mask")}]
        for a in seq:
            v, r, d, inf = gym.step(a)
            steps.append({"action": a, "vec": [float(x) for x in v], "reward": r, "done": d,
                          "mask": inf.get("action_mask")})
            if d:
                break
        fixture["gym"].append({"query": qi, "steps": steps})
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with open(os.path.join(GOLDEN_DIR, name + ".json"), "w", encoding="utf-8") as f:
        json.dump(fixture, f, ensure_ascii=False)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                        index_rows=rs.index_rows.astype(np.float32), query_vecs=qvecs.astype(np.float32))
    print(name, "nodes", rs.hg.G.number_of_nodes(), "edges", rs.hg.G.number_of_edges(),
          "index", len(rs.index_ids), "search cases", len(fixture["search"]))
    return rs


def main() -> None:
    assert ref_runner.reference_available(), "needs /root/reference"
    HG_cls = stubs.import_reference(ref_runner.REF_ROOT)[0]
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for spec in GRAPHS:
        rs = write_fixture(spec["name"], synth.make_graph(**spec["kw"]), [f"golden query {i}" for i in range(6)],
                           SEARCH_CASES, extra={"make_graph_kwargs": spec["kw"]})

    # generate_l1_relations (semantic_aggregator.py:594-680): inputs + the reference's edges and neighbour lists
    import copy
    l1_cases = []
    for seed in range(5):
        for kw in (dict(), dict(min_overlap=1, min_jaccard=0.05, min_cosine=0.3), dict(top_k=7),
                   dict(min_cosine=0.9, min_overlap=99, min_jaccard=0.99)):
            nodes = synth.make_l1_nodes(seed, n_topics=30 if seed < 4 else 60, dim=16 if seed < 4 else 24)
            inp = copy.deepcopy(nodes)
            edges = ref_runner.reference_l1_relations(nodes, **kw)
            l1_cases.append({"seed": seed, "kwargs": kw, "l1_nodes": inp, "edges": edges,
                             "neighbors": [n["neighbors"] for n in nodes]})
    with open(os.path.join(GOLDEN_DIR, "l1_relations.json"), "w", encoding="utf-8") as f:
        json.dump(l1_cases, f)
    print("l1 relation cases", len(l1_cases), "edges", sum(len(c["edges"]) for c in l1_cases))

    # build_from_artifacts (hierarchical_graph.py:241-421): the reference's graph for synthetic artifact sets
    import contextlib
    import io
    import tempfile
    import networkx as nx
    art_cases = []
    for seed, variant in ((0, 0), (1, 1), (2, 2), (3, 3)):
        d = tempfile.mkdtemp(prefix="ahrag_art_")
        synth.write_artifacts(d, seed, variant)
        ref = HG_cls()
        with contextlib.redirect_stdout(io.StringIO()):
            ref.build_from_artifacts(d)
        art_cases.append({"seed": seed, "variant": variant,
                          "node_link": json.loads(json.dumps(nx.node_link_data(ref.G, edges="links"), default=str)),
                          "name_to_entity_id": ref.name_to_entity_id,
                          "topic_to_summary_id": {str(k): v for k, v in ref.topic_to_summary_id.items()}})
    with open(os.path.join(GOLDEN_DIR, "artifacts_graph.json"), "w", encoding="utf-8") as f:
        json.dump(art_cases, f)
    print("artifact graphs", [(len(c["node_link"]["nodes"]), len(c["node_link"]["links"])) for c in art_cases])

    # known answers held by the reference's own tests / shipped data
    HG = rs.HG
    known = {
        "make_entity_id": {"Tim Burton": HG.make_entity_id("Tim Burton")},  # test/test_answer_generator.py:9
        "empty_graph_hash": HG()._graph_snapshot_hash(),  # graph_datasets/hotpotqa_distractor/meta.json:17
        "default_search_params": HG().G.graph["search_params"],  # hierarchical_graph.py:37-47
    }
    with open(os.path.join(GOLDEN_DIR, "known_answers.json"), "w", encoding="utf-8") as f:
        json.dump(known, f, indent=1)
    print(known["make_entity_id"], known["empty_graph_hash"])


if __name__ == "__main__":
    main()
