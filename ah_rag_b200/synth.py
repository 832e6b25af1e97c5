"""Deterministic synthetic inputs for the retrieval hot path (tests and bench.py).

The reference ships no built graph, no vector index and no golden retrieval output
(graph_datasets/hotpotqa_distractor/structure.json has 0 nodes; graph/ and vector_db/ are git-ignored),
and its ingest pipeline needs an LLM + BERTopic.  Everything the engine is exercised on is therefore
generated here, in the on-disk format the reference's loader reads
(``HierarchicalGraph.load``, hierarchical_graph.py:191-218: node-link JSON with the ``"links"`` key)
and with the node/edge attributes its hot path consumes (node_type, judge_scores, confidence, level,
edge_type -- hierarchical_graph.py:128-167, :399-400).
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

ENTITY_TYPES = ["PERSON", "ORG", "LOCATION", "WORK", "EVENT", None]
RELATION_TYPES = ["directed", "founded", "located_in", "member_of", "released"]

DEFAULT_SEARCH_PARAMS = {  # hierarchical_graph.py:37-47
    "alpha": 0.6, "beta": 0.2, "gamma": 0.1, "delta": 0.1,
    "judge_overall_min": None, "confidence_min": None, "member_top_m": 5, "type_filter": None,
    "layer_boost": {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0},
}


def entity_id(name: str) -> str:
    """Stable entity id, same scheme as hierarchical_graph.py:56-61 ('ent:' + sha1(name)[:10])."""
    return "ent:" + hashlib.sha1(name.encode("utf-8")).hexdigest()[:10]


def make_graph(n_entities: int = 60, n_l1: int = 8, n_l2: int = 2, n_hyperedges: int = 25, seed: int = 0,
               shuffle: bool = True, related_density: float = 0.25,
               search_params: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """Small hierarchical knowledge graph as a node-link dict (``"links"`` key).

    Deliberately covers the cases the hot path's order-dependent semantics care about: entities with
    0..3 ``belongs_to`` parents (only the first two are expanded, hierarchical_graph.py:734), summaries
    with more than ``member_top_m`` members, missing ``judge_scores`` / ``confidence``, L2 summaries
    (``level = 2``), hyperedges of arity 2-5, ``related_to`` in both directions, a duplicated link
    whose later copy overwrites ``edge_type`` (DiGraph keeps one edge per ordered pair), and shuffled
    node / link order so adjacency order differs from id order.
    """
    rng = np.random.default_rng(seed)
    nodes: List[Dict[str, Any]] = []
    links: List[Dict[str, Any]] = []

    ent_ids = []
    for i in range(n_entities):
        name = f"Entity {seed}-{i}"
        nid = entity_id(name)
        ent_ids.append(nid)
        node = {"id": nid, "node_type": "entity", "name": name,
                "description": f"synthetic entity number {i}" if i % 5 else None,
                "entity_type": ENTITY_TYPES[int(rng.integers(len(ENTITY_TYPES)))],
                "embedding_ref": None, "l1_parents": None}
        if i % 11 == 3:  # a few entities carry quality attributes (exercises the generic rerank path)
            node["confidence"] = round(float(rng.uniform(1, 10)), 2)
        nodes.append(node)

    l1_ids = [f"sum:{t}" for t in range(n_l1)]
    l2_ids = [f"sum:{10000 + t}" for t in range(n_l2)]
    for t, sid in enumerate(l1_ids + l2_ids):
        is_l2 = t >= n_l1
        judge = None
        if t % 4 != 1:
            overall = round(float(rng.uniform(0, 10)), 1)
            judge = json.dumps({"coherence": 7.0, "overall": overall})
        conf = None if t % 5 == 2 else round(float(rng.uniform(0, 10)), 2)
        node = {"id": sid, "node_type": "summary", "topic_id": int(sid.split(":")[1]),
                "title": f"Topic {sid}", "summary_text": f"summary text of {sid}", "confidence": conf,
                "top_words": json.dumps([f"w{t}a", f"w{t}b"]), "members": None,
                "judge_scores": judge, "centroid": None}
        if is_l2:
            node["level"] = 2
        nodes.append(node)

    hedge_ids = [f"hedge:{seed}-{h}" for h in range(n_hyperedges)]
    for h, hid in enumerate(hedge_ids):
        nodes.append({"id": hid, "node_type": "hyperedge", "description": f"relation {h}",
                      "relation_type": RELATION_TYPES[h % len(RELATION_TYPES)],
                      "confidence_score": round(float(rng.uniform(0, 1)), 3), "source_text_ref": None})

    # belongs_to: entity -> 0..3 L1 parents
    if n_l1:
        for i, e in enumerate(ent_ids):
            n_par = int(rng.choice([0, 1, 2, 3], p=[0.08, 0.47, 0.3, 0.15]))
            for p in rng.permutation(n_l1)[: min(n_par, n_l1)]:
                links.append({"source": e, "target": l1_ids[int(p)], "edge_type": "belongs_to",
                              "prob": round(float(rng.uniform(0.1, 1)), 3)})
    # belongs_to: L1 -> L2
    if n_l2:
        for t, s in enumerate(l1_ids):
            if t % 7 != 6:
                links.append({"source": s, "target": l2_ids[int(rng.integers(n_l2))],
                              "edge_type": "belongs_to", "prob": None})
    # participates_in: entity -> hyperedge, arity 2..5
    for hid in hedge_ids:
        arity = int(rng.integers(2, 6))
        for e in rng.permutation(n_entities)[: min(arity, n_entities)]:
            links.append({"source": ent_ids[int(e)], "target": hid, "edge_type": "participates_in",
                          "role": None})
    # related_to among L1 summaries (directed; sometimes both directions)
    for a in range(n_l1):
        for b in range(n_l1):
            if a != b and rng.uniform() < related_density:
                links.append({"source": l1_ids[a], "target": l1_ids[b], "edge_type": "related_to",
                              "weight": round(float(rng.uniform()), 3), "jaccard": None, "cosine": None,
                              "overlap": None, "confidence": None})
    if n_l2 >= 2:
        links.append({"source": l2_ids[0], "target": l2_ids[1], "edge_type": "related_to", "weight": 0.5})

    if shuffle:
        nodes = [nodes[int(i)] for i in rng.permutation(len(nodes))]
        links = [links[int(i)] for i in rng.permutation(len(links))]

    # a duplicated ordered pair: the later copy overwrites edge_type but keeps the adjacency position
    bel = [l for l in links if l["edge_type"] == "belongs_to" and l["source"].startswith("sum:")]
    if bel:
        links.append({"source": bel[0]["source"], "target": bel[0]["target"], "edge_type": "related_to",
                      "weight": 0.9})

    sp = dict(DEFAULT_SEARCH_PARAMS)
    if search_params:
        sp.update(search_params)
    return {
        "directed": True, "multigraph": False,
        "graph": {"search_params": sp, "dirty": False,
                  "vector_index": {"db_path": "vector_db", "model": "all-MiniLM-L6-v2",
                                   "indexed_nodes": 0, "indexed_meta": {}}},
        "nodes": nodes, "links": links,
    }


def write_graph_dir(directory: str, node_link: Dict[str, Any]) -> str:
    """Write ``structure.json`` + ``meta.json`` the way ``HierarchicalGraph.save`` lays them out
    (hierarchical_graph.py:175-189), using the ``"links"`` key its ``load`` requires (:196)."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "structure.json"), "w", encoding="utf-8") as f:
        json.dump(node_link, f, ensure_ascii=False)
    meta = {"search_params": node_link["graph"].get("search_params", {}),
            "dirty": False, "vector_index": node_link["graph"].get("vector_index", {})}
    with open(os.path.join(directory, "meta.json"), "w", encoding="utf-8") as f:
        json.dump(meta, f, ensure_ascii=False)
    return directory


def unit_rows(n: int, dim: int, seed: int, clustered: int = 0, noise: float = 0.3) -> np.ndarray:
    """Seeded L2-normalised fp32 rows; ``clustered`` > 0 draws rows around that many centroids so
    near-duplicates and close scores occur (top-k is then non-trivial)."""
    rng = np.random.default_rng(seed)
    if clustered:
        cent = rng.standard_normal((clustered, dim)).astype(np.float32)
        x = cent[rng.integers(clustered, size=n)] + noise * rng.standard_normal((n, dim)).astype(np.float32)
    else:
        x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# Array-level generator for graphs too large for NetworkX (BASELINE config c5: 20 M nodes / 500 M edges)
# ---------------------------------------------------------------------------------------------
def make_csr_graph(n_entities: int, n_l1: int, n_l2: int, n_hyperedges: int, seed: int = 42,
                   parents_per_entity: Tuple[int, int] = (1, 2), arity: Tuple[int, int] = (2, 5),
                   related_per_l1: int = 4) -> Dict[str, np.ndarray]:
    """Edge lists (src, dst, type) of a synthetic hierarchical graph in dense ids.

    Dense id layout: [entities | L1 summaries | L2 summaries | hyperedges].  Returned ``edges`` are in
    "link order"; adjacency order per node is the order of appearance, exactly what a NetworkX load of
    the same link list would give, so ``csr.csr_from_edges`` builds the same CSR the JSON path does.
    """
    rng = np.random.default_rng(seed)
    e0, s1, s2, h0 = 0, n_entities, n_entities + n_l1, n_entities + n_l1 + n_l2
    n_nodes = h0 + n_hyperedges
    src, dst, typ = [], [], []
    # belongs_to entity -> L1
    lo, hi = parents_per_entity
    cnt = rng.integers(lo, hi + 1, size=n_entities)
    es = np.repeat(np.arange(n_entities, dtype=np.int64), cnt)
    src.append(es); dst.append(s1 + rng.integers(n_l1, size=es.size)); typ.append(np.full(es.size, 1, np.uint8))
    # belongs_to L1 -> L2
    if n_l2:
        a = np.arange(n_l1, dtype=np.int64)
        src.append(s1 + a); dst.append(s2 + rng.integers(n_l2, size=n_l1)); typ.append(np.full(n_l1, 1, np.uint8))
    # participates_in entity -> hyperedge
    lo, hi = arity
    ar = rng.integers(lo, hi + 1, size=n_hyperedges)
    hs = np.repeat(np.arange(n_hyperedges, dtype=np.int64), ar)
    src.append(rng.integers(n_entities, size=hs.size)); dst.append(h0 + hs); typ.append(np.full(hs.size, 0, np.uint8))
    # related_to among L1
    if related_per_l1 and n_l1 > 1:
        a = np.repeat(np.arange(n_l1, dtype=np.int64), related_per_l1)
        b = rng.integers(n_l1 - 1, size=a.size)
        b = b + (b >= a)
        src.append(s1 + a); dst.append(s1 + b); typ.append(np.full(a.size, 2, np.uint8))
    src = np.concatenate(src); dst = np.concatenate(dst); typ = np.concatenate(typ)
    node_type = np.empty(n_nodes, np.uint8)
    node_type[e0:s1] = 0; node_type[s1:h0] = 1; node_type[h0:] = 2
    level = np.zeros(n_nodes, np.uint8)
    level[s2:h0] = 2
    judge = np.full(n_nodes, np.nan, np.float64)
    conf = np.full(n_nodes, np.nan, np.float64)
    ns = n_l1 + n_l2
    judge[s1:h0] = np.round(rng.uniform(0, 10, ns), 1)
    conf[s1:h0] = np.round(rng.uniform(0, 10, ns), 2)
    judge[s1:h0][rng.uniform(size=ns) < 0.2] = np.nan
    return {"n_nodes": np.int64(n_nodes), "src": src, "dst": dst, "etype": typ, "node_type": node_type,
            "level": level, "judge": judge, "conf": conf}


def make_l1_nodes(seed: int, n_topics: int = 30, dim: int = 16, n_entities: int = 60) -> List[Dict[str, Any]]:
    """Synthetic L1 topic records (topic_id, members, centroid, confidence) for ``generate_l1_relations``: clustered
    centroids (so cosines spread over [-1, 1]), overlapping member sets, and the irregularities the reference tolerates
    (missing / null centroid, missing confidence, permuted and duplicated topic ids)."""
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(4, dim))
    nodes: List[Dict[str, Any]] = []
    for t in range(n_topics):
        node: Dict[str, Any] = {"topic_id": int(t * 3 % n_topics) if seed % 2 else t,
                                "members": [f"e{int(x)}" for x in rng.integers(0, n_entities, size=int(rng.integers(0, 12)))],
                                "confidence": float(np.round(rng.uniform(1, 9), 2))}
        centroid = (base[t % 4] + 0.6 * rng.normal(size=dim)).tolist() if rng.random() > 0.1 else None
        if centroid is not None or rng.random() > 0.5:
            node["centroid"] = centroid
        if rng.random() < 0.1:
            node.pop("confidence")
        nodes.append(node)
    return nodes


def write_artifacts(directory: str, seed: int, variant: int) -> None:
    """Synthetic JSON artifacts of the reference's ingest pipeline, as ``HierarchicalGraph.build_from_artifacts`` reads
    them (hierarchical_graph.py:241-421).  ``variant`` 0: topics.json + extractions.json only; 1: + l1_nodes.json and
    l1_edges.json; 2: + the L2 level and judge files; 3: as 2 without topics.json.  Irregularities on purpose: parents
    naming a topic that does not exist, entities only seen in hyperedges, entity records without a name, hyperedges
    without id / relation_type, edges to unknown topics, an L2 topic id that clashes with an L1 one."""
    rng = np.random.default_rng(seed)
    names = [f"Entity {i}" for i in range(25)]
    n_topics = 6
    e2p: Dict[str, Any] = {}
    for n in names[:20]:
        picks = rng.choice(n_topics + 1, size=int(rng.integers(0, 3)), replace=False)
        e2p[n] = [{"topic_id": int(t), "prob": float(np.round(rng.uniform(0.1, 1), 3))} for t in picks]
    l1 = [{"topic_id": t, "title": f"T{t}" if t % 2 else None, "summary": f"about {t}" if t % 3 else None,
           "summary_text": f"alt {t}", "confidence": float(t) + 0.5 if t != 2 else None,
           "top_words": ["a", "b"] if t % 2 else [], "members": names[t:t + 4] if t else None,
           "centroid": rng.normal(size=4).tolist() if t != 1 else None} for t in range(n_topics)]
    edges = [{"source": 0, "target": 1, "weight": 0.5, "jaccard": 0.2, "cosine": 0.8, "overlap": 2, "confidence": 5.0},
             {"source": 2, "target": 9, "weight": 0.1}, {"source": 3, "target": 4}]
    hyper = []
    for i in range(12):
        ents: List[Dict[str, Any]] = []
        for _ in range(int(rng.integers(1, 4))):
            ents.append({"name": str(rng.choice(names)), "type": [None, "PERSON", "ORG"][int(rng.integers(3))],
                         "description": [None, "", "desc A", "desc B"][int(rng.integers(4))],
                         "role": [None, "subj", "obj"][int(rng.integers(3))]})
        if i == 3:
            ents += [{"name": None}, {"type": "X"}]
        h = {"hyperedge": f"fact {i}", "relation_type": ["rel", "other"][int(rng.integers(2))],
             "confidence_score": float(np.round(rng.uniform(1, 10), 1)), "entities": ents}
        if i % 4 == 0:
            h["id"] = f"u{i}"
        if i == 5:
            h.pop("relation_type")
        hyper.append(h)
    os.makedirs(directory, exist_ok=True)

    def dump(name: str, obj: Any) -> None:
        with open(os.path.join(directory, name), "w", encoding="utf-8") as f:
            json.dump(obj, f)

    if variant != 3:
        dump("topics.json", {"entity_to_parents": e2p, "l1_nodes": l1[:3]})
    if variant >= 1:
        dump("l1_nodes.json", l1)
        dump("l1_edges.json", edges)
    dump("extractions.json", hyper)
    if variant >= 2:
        dump("l2_nodes.json", [{"topic_id": 100, "title": "L2a", "summary": "s", "confidence": 7.0, "top_words": ["x"],
                                "members": ["1", "2"], "centroid": [0.1, 0.2]},
                               {"topic_id": 3, "title": "clash with L1 3", "members": []}])
        dump("l1_to_l2.json", {"0": 100, "1": 100, "x": 100, "4": 3, "5": 777})
        dump("l1_judge_nodes.json", [{"id": 0, "overall": 8.0, "comments": "ok"}, {"id": 100, "overall": 6.5},
                                     {"id": 55, "overall": 1}])
        dump("l1_judge_edges.json", [{"source": 0, "target": 1, "overall": 5}])
