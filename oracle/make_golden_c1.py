"""BASELINE config c1: the ``samples/hotpotqa_mini.txt`` graph and the reference's own test path over it
(dev container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden_c1          # writes tests/golden/g_c1_hotpotqa_mini.{json,npz}

The reference ships the sample as raw text, not as a graph, and its ingest needs an LLM
endpoint plus BERTopic/UMAP/HDBSCAN.  The graph is therefore built with the pieces of the reference that do run here:

  passages    ``### title`` blocks of /root/reference/samples/hotpotqa_mini.txt
  extraction  the reference's own no-LLM path, ``HypergraphExtractor._fallback_extract``
              (src/ah_rag/extract/hypergraph_extractor.py:214-264), unmodified
  clustering  a deterministic stand-in for the BERTopic stage (OURS, labelled as such): an entity's L1 topic is a
              SHA-1 bucket of its name, an entity seen in several passages gets a second parent, topics that share a
              hyperedge are ``related_to``, two L2 topics, judge scores / confidences derived from the bucket number
  graph       the reference's own ``HierarchicalGraph.build_from_artifacts`` (hierarchical_graph.py:241-421) + ``save``
  embeddings  a deterministic bag-of-words hash embedding (embedding-model inference is out of scope), so that the
              query "Tim Burton" of test/test_graph_search.py:10 does land on the Tim Burton nodes

Expected values are then produced by the UNMODIFIED reference exactly as for the synthetic fixtures
(``make_golden.write_fixture``): ``search("Tim Burton", top_k=3)`` and 63 more queries (test/test_graph_search.py:7-15),
``GraphEnvironment.reset(seed_query="Tim Burton", top_k=3)`` + expand traces (test/test_environment.py:7-12), gym and
expand_to_lca traces.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden, ref_runner, stubs  # noqa: E402

SAMPLE = os.path.join(ref_runner.REF_ROOT, "samples", "hotpotqa_mini.txt")
NAME = "g_c1_hotpotqa_mini"
N_L1, N_QUERIES = 8, 64

C1_SEARCH_CASES = [dict(top_k=3), dict(top_k=5), dict(top_k=3, member_top_m=2), dict(top_k=10, type_filter=["summary"]),
                   dict(top_k=8, judge_overall_min=4.0), dict(top_k=50, member_top_m=5)]


def _bucket(text: str, mod: int) -> int:
    return int.from_bytes(hashlib.sha1(text.encode("utf-8")).digest()[:4], "little") % mod


def lexical_embed(text: str, dim: int = stubs.EMBED_DIM) -> np.ndarray:
    """Unit vector = normalised sum of one hash vector per lower-cased word (deterministic, order-free)."""
    words = re.findall(r"[a-z0-9]+", text.lower()) or ["<empty>"]
    acc = np.zeros(dim, dtype=np.float64)
    for w in words:
        seed = int.from_bytes(hashlib.sha1(("w:" + w).encode("utf-8")).digest()[:8], "little")
        acc += np.random.default_rng(seed).standard_normal(dim)
    acc /= np.linalg.norm(acc)
    v = acc.astype(np.float32)
    return (v / np.float32(np.linalg.norm(v.astype(np.float64)))).astype(np.float32)


def passages():
    with open(SAMPLE, "r", encoding="utf-8") as f:
        txt = f.read()
    seen, out = set(), []
    for block in re.split(r"^### ", txt, flags=re.M):
        block = block.strip()
        if not block:
            continue
        title = block.split("\n", 1)[0].strip()
        if title in seen:  # the sample repeats passages shared by several questions
            continue
        seen.add(title)
        out.append((title, block))
    return out


def build_artifacts(directory: str):
    """extractions.json from the reference's fallback extractor; topics / L1 / L2 / judge files from the stand-in."""
    src = os.path.join(ref_runner.REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    stubs.install()
    from ah_rag.extract.hypergraph_extractor import HypergraphExtractor  # the reference's own class
    ex = HypergraphExtractor()
    hyperedges, seen_in = [], {}
    for pi, (title, block) in enumerate(passages()):
        with contextlib.redirect_stdout(io.StringIO()):
            items = ex._fallback_extract(block)
        for it in items:
            h = it.model_dump()
            hyperedges.append(h)
            for e in h["entities"]:
                seen_in.setdefault(e["name"], set()).add(pi)
    names = list(seen_in)  # first-appearance order
    e2p, members = {}, {t: [] for t in range(N_L1)}
    for nm in names:
        t1 = _bucket(nm, N_L1)
        parents = [{"topic_id": t1, "prob": round(0.5 + _bucket(nm + "#p", 50) / 100.0, 2)}]
        if len(seen_in[nm]) > 1:
            t2 = _bucket(nm + "#2", N_L1)
            if t2 != t1:
                parents.append({"topic_id": t2, "prob": round(0.1 + _bucket(nm + "#q", 40) / 100.0, 2)})
        e2p[nm] = parents
        for p in parents:
            members[p["topic_id"]].append(nm)
    l1_nodes = []
    for t in range(N_L1):
        ms = members[t]
        words = [w.lower() for m in ms[:6] for w in m.split()][:8]
        l1_nodes.append({"topic_id": t, "title": "Topic %d: %s" % (t, ", ".join(ms[:3])),
                         "summary": "Entities grouped in bucket %d: %s." % (t, "; ".join(ms[:12])),
                         "confidence": float(4 + (t * 3) % 6) + 0.5 * (t % 2), "top_words": words, "members": ms,
                         "centroid": None})
    pair_count = {}
    for h in hyperedges:
        ts = sorted({p["topic_id"] for e in h["entities"] for p in e2p[e["name"]]})
        for i, a in enumerate(ts):
            for b in ts[i + 1:]:
                pair_count[(a, b)] = pair_count.get((a, b), 0) + 1
    l1_edges = [{"source": a, "target": b, "weight": float(c), "overlap": c, "jaccard": round(c / (c + 10.0), 4),
                 "cosine": None, "confidence": None}
                for (a, b), c in sorted(pair_count.items(), key=lambda kv: (-kv[1], kv[0])) if c >= 3]
    l2_nodes = [{"topic_id": 100 + j, "title": "Theme %d" % j, "summary": "Theme %d over the L1 buckets %s" %
                 (j, [t for t in range(N_L1) if t % 2 == j]), "confidence": 6.0 + j, "top_words": ["theme", str(j)],
                 "members": ["sum:%d" % t for t in range(N_L1) if t % 2 == j], "centroid": None} for j in range(2)]
    l1_to_l2 = {str(t): 100 + t % 2 for t in range(N_L1)}
    judge_nodes = [{"id": t, "overall": float(2 + (t * 5) % 8), "coherence": float(1 + t % 9)} for t in range(N_L1)
                   if t != 5]  # bucket 5 stays unjudged (judge is None there)
    judge_nodes.append({"id": 100, "overall": 7.5})
    files = {"extractions.json": hyperedges, "topics.json": {"entity_to_parents": e2p, "l1_nodes": l1_nodes},
             "l1_nodes.json": l1_nodes, "l1_edges.json": l1_edges, "l2_nodes.json": l2_nodes, "l1_to_l2.json": l1_to_l2,
             "l1_judge_nodes.json": judge_nodes}
    os.makedirs(directory, exist_ok=True)
    for fn, obj in files.items():
        with open(os.path.join(directory, fn), "w", encoding="utf-8") as f:
            json.dump(obj, f, ensure_ascii=False)
    return names


def build_node_link():
    """The mini graph as the reference's own build_from_artifacts + save produce it (node-link, "links" key)."""
    import networkx as nx
    HG = stubs.import_reference(ref_runner.REF_ROOT)[0]
    art = tempfile.mkdtemp(prefix="ahrag_c1_art_")
    names = build_artifacts(art)
    hg = HG()
    with contextlib.redirect_stdout(io.StringIO()):
        hg.build_from_artifacts(art)
    nl = json.loads(json.dumps(nx.node_link_data(hg.G, edges="links"), default=str))
    return nl, names


def main() -> None:
    assert ref_runner.reference_available(), "needs /root/reference"
    nl, names = build_node_link()
    titles = [t for t, _ in passages()]
    queries = ["Tim Burton"]  # test/test_graph_search.py:10, test/test_environment.py:9
    for q in titles + names:
        if q not in queries and len(queries) < N_QUERIES:
            queries.append(q)
    plain = stubs.hash_embed
    stubs.hash_embed = lexical_embed  # the stand-in embedder of THIS fixture (rows and query vectors are committed)
    try:
        make_golden.write_fixture(NAME, nl, queries, C1_SEARCH_CASES, env_queries=(0, 1), env_top_k=3,
                                  extra={"source": "samples/hotpotqa_mini.txt via the reference's _fallback_extract + "
                                                   "build_from_artifacts; clustering stand-in is ours (oracle/make_golden_c1.py)"})
    finally:
        stubs.hash_embed = plain


if __name__ == "__main__":
    main()
