"""Drop-in for ``ah_rag.graph.hierarchical_graph.HierarchicalGraph`` on the retrieval hot path.

Same public surface the reference's callers use (hierarchical_graph.py: ``load`` :191, ``save`` :175,
``build_vector_index`` :602, ``search`` :675, ``.G``, ``name_to_entity_id``, ``topic_to_summary_id``, the id
helpers :56-69 and the edge-type aware neighbour helpers :441-470), with the two things underneath
replaced by the B200 engine:

  * the Chroma / HNSW collection  -> ``engine.VectorIndex`` (device-resident rows, batched fused top-k)
  * the NetworkX adjacency walks + Python rerank loop -> ``engine.GraphCSR.anchor_rerank`` (one kernel)

``search`` keeps its signature and result schema; ``search_batch`` is the additive batched entry point the
single-query call delegates to.  The graph itself stays a ``networkx.DiGraph`` on the host (it is the
on-disk / in-memory contract other reference modules read directly); the CSR is exported from it in the
LOADED adjacency order.  Ingest (``build_from_artifacts``: LLM extraction + BERTopic) and embedding-model
inference are out of scope: embeddings come from a pluggable ``embedder`` callable.

There is no CPU fallback: ``search`` raises if CUDA or the compiled library is unavailable.
"""
from __future__ import annotations

import datetime
import hashlib
import json
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import networkx as nx
import numpy as np

ARTIFACT_DIR = "artifacts"
GRAPH_DIR = "graph"

_DEFAULT_SEARCH_PARAMS = {
    "alpha": 0.6, "beta": 0.2, "gamma": 0.1, "delta": 0.1,
    "judge_overall_min": None, "confidence_min": None, "member_top_m": 5, "type_filter": None,
    "layer_boost": {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0},
}

Embedder = Callable[[List[str]], np.ndarray]


class HierarchicalGraph:
    def __init__(self) -> None:
        self.G: nx.DiGraph = nx.DiGraph()
        self.name_to_entity_id: Dict[str, str] = {}
        self.topic_to_summary_id: Dict[int, str] = {}
        self.G.graph.setdefault("search_params", json.loads(json.dumps(_DEFAULT_SEARCH_PARAMS)))
        self.G.graph.setdefault("dirty", False)
        self.G.graph.setdefault("vector_index", {"db_path": "vector_db", "model": "all-MiniLM-L6-v2",
                                                 "indexed_nodes": 0, "indexed_meta": {}})
        # engine-side state
        self.embedder: Optional[Embedder] = None  # texts -> [n, d] float32, L2-normalised
        self.index_dtype: str = "fp32"            # "fp32": exact mode (faithful) | "bf16": tensor-core mode
        self.device: str = "cuda"
        self._dev_graph: Optional[Dict[str, Any]] = None
        self._dev_index: Optional[Dict[str, Any]] = None
        self._source: Optional[Dict[str, Any]] = None  # set by load(): lets the device export skip the NetworkX walk
        self._host_arrays: Optional[Tuple[Any, Dict[str, Any]]] = None
        self._st_model = None
        self._st_name: Optional[str] = None
        self._shard: Optional[Dict[str, Any]] = None  # set by enable_sharding(): multi-GPU form of the same API

    # ------------------------------------------------------------------ ids (hierarchical_graph.py:56-69)
    @staticmethod
    def _hash(text: str, length: int = 10) -> str:
        return hashlib.sha1(text.encode("utf-8")).hexdigest()[:length]

    @staticmethod
    def make_entity_id(name: str) -> str:
        return "ent:" + HierarchicalGraph._hash(name)

    @staticmethod
    def make_hyperedge_id(uid: str) -> str:
        return f"hedge:{uid}"

    @staticmethod
    def make_summary_id(topic_id: int) -> str:
        return f"sum:{int(topic_id)}"

    # ------------------------------------------------------------------ construction (:74-168, host only)
    def _touch(self) -> None:
        self.G.graph["dirty"] = True
        self._dev_graph = None
        self._host_arrays = None
        self._source = None  # the graph no longer is the document it was loaded from

    def add_entity(self, name: str, description: Optional[str] = None, entity_type: Optional[str] = None,
                   embedding_ref: Optional[Tuple[str, int]] = None,
                   l1_parents: Optional[Dict[str, float]] = None) -> str:
        nid = self.name_to_entity_id.get(name)
        if nid is None:
            nid = self.make_entity_id(name)
            self.G.add_node(nid, node_type="entity", name=name, description=description, entity_type=entity_type,
                            embedding_ref=embedding_ref, l1_parents=json.dumps(l1_parents) if l1_parents else None)
            self.name_to_entity_id[name] = nid
        else:
            d = self.G.nodes[nid]
            if description is not None and not d.get("description"):
                d["description"] = description
            if entity_type is not None and not d.get("entity_type"):
                d["entity_type"] = entity_type
            if embedding_ref is not None:
                d["embedding_ref"] = embedding_ref
            if l1_parents is not None:
                d["l1_parents"] = json.dumps(l1_parents)
        self._touch()
        return nid

    def add_hyperedge(self, uid: str, description: str, relation_type: str,
                      confidence_score: Optional[float] = None, source_text_ref: Optional[str] = None) -> str:
        nid = self.make_hyperedge_id(uid)
        if not self.G.has_node(nid):
            self.G.add_node(nid, node_type="hyperedge", description=description, relation_type=relation_type,
                            confidence_score=confidence_score, source_text_ref=source_text_ref)
        self._touch()
        return nid

    def add_summary(self, topic_id: int, title: Optional[str] = None, summary_text: Optional[str] = None,
                    confidence: Optional[float] = None, top_words: Optional[List[str]] = None,
                    members: Optional[List[str]] = None, judge_scores: Optional[Dict[str, Any]] = None,
                    centroid: Optional[List[float]] = None) -> str:
        fields = {"title": title, "summary_text": summary_text, "confidence": confidence,
                  "top_words": json.dumps(top_words) if top_words else None,
                  "members": json.dumps(members) if members else None,
                  "judge_scores": json.dumps(judge_scores) if judge_scores else None,
                  "centroid": json.dumps(centroid) if centroid is not None else None}
        nid = self.topic_to_summary_id.get(int(topic_id))
        if nid is None:
            nid = self.make_summary_id(topic_id)
            self.G.add_node(nid, node_type="summary", topic_id=int(topic_id), **fields)
            self.topic_to_summary_id[int(topic_id)] = nid
        else:
            d = self.G.nodes[nid]
            given = {"title": title, "summary_text": summary_text, "confidence": confidence, "top_words": top_words,
                     "members": members, "judge_scores": judge_scores, "centroid": centroid}
            for key, val in given.items():
                if val is not None:  # an update serialises whatever it is given, even an empty list (:145-152)
                    d[key] = json.dumps(val) if key in ("top_words", "members", "judge_scores", "centroid") else val
        self._touch()
        return nid

    def add_participation(self, entity_id: str, hyperedge_id: str, role: Optional[str] = None) -> None:
        self.G.add_edge(entity_id, hyperedge_id, edge_type="participates_in", role=role)
        self._touch()

    def add_belongs_to(self, entity_id: str, summary_id: str, prob: Optional[float] = None) -> None:
        self.G.add_edge(entity_id, summary_id, edge_type="belongs_to", prob=prob)
        self._touch()

    def add_related(self, summary_a: str, summary_b: str, weight: Optional[float] = None,
                    jaccard: Optional[float] = None, cosine: Optional[float] = None, overlap: Optional[int] = None,
                    confidence: Optional[float] = None) -> None:
        self.G.add_edge(summary_a, summary_b, edge_type="related_to", weight=weight, jaccard=jaccard, cosine=cosine,
                        overlap=overlap, confidence=confidence)
        self._touch()

    def build_from_artifacts(self, artifacts_dir: str = ARTIFACT_DIR) -> None:
        """Assemble the graph from the JSON artifacts of the reference's ingest pipeline
        (hierarchical_graph.py:241-421; the data format in front of the retrieval path):

          topics.json            {"entity_to_parents": {name: [{"topic_id", "prob"}]}, "l1_nodes": [...]}
          l1_nodes.json          L1 topic records (replaces topics.json's list when present)
          l1_edges.json          related_summary edges (``aggregate.l1_relations.generate_l1_relations``)
          extractions.json       hyperedges with their participating entities
          l2_nodes.json, l1_to_l2.json      second summary level and the L1 -> L2 assignment
          l1_judge_nodes.json, l1_judge_edges.json   judge scores (per summary node / kept at graph level)

        Insertion order is the reference's (it fixes node and adjacency order, hence the CSR): entities named by
        ``entity_to_parents``, entities only seen in hyperedges, L1 summaries, belongs_to, related_to, hyperedges with
        their participation edges, L2 summaries and L1 -> L2 belongs_to, judge scores.  Missing files are skipped."""
        def read(name: str, default: Any) -> Any:
            path = os.path.join(artifacts_dir, name)
            if not os.path.exists(path):
                return default
            with open(path, "r", encoding="utf-8") as f:
                return json.load(f)

        topics = read("topics.json", {})
        entity_to_parents: Dict[str, List[Dict[str, Any]]] = topics.get("entity_to_parents", {}) if topics else {}
        l1_nodes = read("l1_nodes.json", None)
        if l1_nodes is None:
            l1_nodes = topics.get("l1_nodes", []) if topics else []
        l1_edges = read("l1_edges.json", [])
        hyperedges = read("extractions.json", [])
        l2_nodes = read("l2_nodes.json", [])
        l1_to_l2 = read("l1_to_l2.json", {})
        judge_nodes = read("l1_judge_nodes.json", [])
        judge_edges = read("l1_judge_edges.json", [])

        # what the hyperedges say about each entity: type of its first mention, every distinct non-empty description
        seen: Dict[str, Dict[str, Any]] = {}
        for h in hyperedges:
            for ent in h.get("entities", []):
                name = ent.get("name")
                if not name:
                    continue
                rec = seen.setdefault(name, {"entity_type": ent.get("type"), "descriptions": []})
                text = ent.get("description")
                if text and text not in rec["descriptions"]:
                    rec["descriptions"].append(text)

        def described(name: str) -> Tuple[Optional[str], Optional[str]]:
            rec = seen.get(name, {})
            texts = rec.get("descriptions", [])
            return ("; ".join(texts) if texts else None), rec.get("entity_type")

        # the reference points every entity at the module-level artifacts directory, whatever directory it was given
        embedding_ref = (os.path.join(ARTIFACT_DIR, "embeddings.npy"), -1)
        for name, parents in entity_to_parents.items():
            text, etype = described(name)
            self.add_entity(name=name, description=text, entity_type=etype, embedding_ref=embedding_ref,
                            l1_parents={str(p.get("topic_id")): p.get("prob") for p in parents})
        for name in seen:
            if name not in entity_to_parents:
                text, etype = described(name)
                self.add_entity(name=name, description=text, entity_type=etype)

        for node in l1_nodes:
            self.add_summary(topic_id=int(node.get("topic_id")), title=node.get("title"),
                             summary_text=node.get("summary") or node.get("summary_text"),
                             confidence=node.get("confidence"), top_words=node.get("top_words") or [],
                             members=node.get("members") or [], centroid=node.get("centroid"))

        for name, parents in entity_to_parents.items():
            eid = self.name_to_entity_id.get(name)
            if not eid:
                continue
            for p in parents:
                sid = self.topic_to_summary_id.get(int(p.get("topic_id")))
                if sid:
                    self.add_belongs_to(eid, sid, prob=p.get("prob"))

        for e in l1_edges:
            a = self.topic_to_summary_id.get(int(e.get("source")))
            b = self.topic_to_summary_id.get(int(e.get("target")))
            if a and b:
                self.add_related(a, b, weight=e.get("weight"), jaccard=e.get("jaccard"), cosine=e.get("cosine"),
                                 overlap=e.get("overlap"), confidence=e.get("confidence"))

        for pos, h in enumerate(hyperedges):
            uid = h.get("id")
            if uid is None:
                uid = f"hedge_{pos}_{h.get('relation_type', 'unknown')}"
            hid = self.add_hyperedge(uid=str(uid), description=h.get("hyperedge"), relation_type=h.get("relation_type"),
                                     confidence_score=h.get("confidence_score"))
            for ent in h.get("entities", []):
                eid = self.name_to_entity_id.get(ent.get("name")) if ent.get("name") else None
                if eid:
                    self.add_participation(eid, hid, role=ent.get("role") if isinstance(ent, dict) else None)

        if l2_nodes:
            for n in l2_nodes:
                sid = self.add_summary(topic_id=int(n.get("topic_id")), title=n.get("title"), summary_text=n.get("summary"),
                                       confidence=n.get("confidence"), top_words=n.get("top_words"),
                                       members=n.get("members"), centroid=n.get("centroid"))
                self.G.nodes[sid]["level"] = 2
            for l1_key, l2_tid in l1_to_l2.items():
                try:
                    l1_tid = int(l1_key)
                except Exception:
                    continue
                lower, upper = self.topic_to_summary_id.get(l1_tid), self.topic_to_summary_id.get(int(l2_tid))
                if lower and upper:
                    self.add_belongs_to(lower, upper, prob=1.0)

        by_topic = {int(d.get("topic_id")): n for n, d in self.G.nodes(data=True) if d.get("node_type") == "summary"}
        for score in judge_nodes:
            nid = by_topic.get(int(score.get("id")))
            if nid:
                self.G.nodes[nid]["judge_scores"] = json.dumps(score)
        if judge_edges:
            self.G.graph["judge_edges"] = judge_edges
        self._touch()

    # ------------------------------------------------------------------ persistence (:175-236)
    def save(self, directory: str = GRAPH_DIR, meta: Optional[Dict[str, Any]] = None) -> None:
        """Node-link JSON with the ``"links"`` key ``load`` reads (the reference's own save() writes ``"edges"``
        under networkx >= 3.6 and can no longer be loaded -- SURVEY.md section 0-7)."""
        os.makedirs(directory, exist_ok=True)
        data = nx.node_link_data(self.G, edges="links")
        with open(os.path.join(directory, "structure.json"), "w", encoding="utf-8") as f:
            json.dump(data, f, ensure_ascii=False, indent=2)
        merged = dict(meta) if meta else {}
        merged["search_params"] = self.G.graph.get("search_params", {})
        merged["graph_hash"] = self._graph_snapshot_hash()
        merged["dirty"] = self.G.graph.get("dirty", False)
        merged["vector_index"] = self.G.graph.get("vector_index", {})
        with open(os.path.join(directory, "meta.json"), "w", encoding="utf-8") as f:
            json.dump(merged, f, ensure_ascii=False, indent=2)

    @classmethod
    def load(cls, directory: str = GRAPH_DIR) -> "HierarchicalGraph":
        with open(os.path.join(directory, "structure.json"), "rb") as f:
            raw = f.read()
        data = json.loads(raw.decode("utf-8"))
        hg = cls()
        key = "links" if "links" in data else "edges"
        from . import snapshot
        # what the device export needs later (csr.csr_from_node_link, or the binary snapshot keyed by these bytes);
        # node_link_graph copies the attribute dicts it reads and leaves the document untouched
        hg._source = {"dir": directory, "key": snapshot.key_of(raw), "doc": data, "edges": key}
        hg.G = nx.node_link_graph(data, directed=True, edges=key)
        hg._source["sig"] = hg._graph_sig()[1:]  # (nodes, edges[, attribute fingerprint]) of the document as loaded
        for nid, attr in hg.G.nodes(data=True):
            if attr.get("node_type") == "entity" and attr.get("name"):
                hg.name_to_entity_id[attr["name"]] = nid
            if attr.get("node_type") == "summary" and attr.get("topic_id") is not None:
                hg.topic_to_summary_id[int(attr["topic_id"])] = nid
        meta_path = os.path.join(directory, "meta.json")
        if os.path.exists(meta_path):
            try:
                with open(meta_path, "r", encoding="utf-8") as f:
                    meta = json.load(f)
                if isinstance(meta.get("search_params"), dict):
                    hg.G.graph["search_params"] = {**hg.G.graph.get("search_params", {}), **meta["search_params"]}
                if isinstance(meta.get("vector_index"), dict):
                    hg.G.graph["vector_index"] = meta["vector_index"]
                if isinstance(meta.get("dirty"), bool):
                    hg.G.graph["dirty"] = meta["dirty"]
            except Exception:
                pass
        return hg

    def _graph_snapshot_hash(self) -> str:
        items = [(nid, d.get("node_type"), d.get("name"), d.get("title"), d.get("summary_text"), d.get("description"))
                 for nid, d in self.G.nodes(data=True)]
        items.sort(key=lambda x: x[0])
        return hashlib.sha1(json.dumps(items, ensure_ascii=False).encode("utf-8")).hexdigest()

    # ------------------------------------------------------------------ host neighbour helpers (:426-470)
    def get_parents(self, node_id: str) -> List[str]:
        return list(self.G.successors(node_id))

    def get_children(self, node_id: str) -> List[str]:
        return list(self.G.predecessors(node_id))

    def find_entity(self, name: str) -> Optional[str]:
        return self.name_to_entity_id.get(name)

    def find_summary(self, topic_id: int) -> Optional[str]:
        return self.topic_to_summary_id.get(int(topic_id))

    def _neighbors_by_edge_type(self, node_id: str, direction: str, edge_type: str) -> List[str]:
        if direction == "out":
            return [v for _, v, d in self.G.out_edges(node_id, data=True) if d.get("edge_type") == edge_type]
        return [u for u, _, d in self.G.in_edges(node_id, data=True) if d.get("edge_type") == edge_type]

    def get_belongs_to(self, node_id: str) -> List[str]:
        return self._neighbors_by_edge_type(node_id, "out", "belongs_to")

    def get_summary_members(self, summary_id: str) -> List[str]:
        return self._neighbors_by_edge_type(summary_id, "in", "belongs_to")

    def get_hyperedge_participants(self, hyperedge_id: str) -> List[str]:
        return self._neighbors_by_edge_type(hyperedge_id, "in", "participates_in")

    def get_entity_hyperedges(self, entity_id: str) -> List[str]:
        return self._neighbors_by_edge_type(entity_id, "out", "participates_in")

    # ------------------------------------------------------------------ small host utilities (:471-553)
    def get_siblings(self, node_id: str) -> List[str]:
        """Other members of the summaries ``node_id`` belongs to (unordered, as the reference's ``list(set)``)."""
        found = {m for parent in set(self.get_belongs_to(node_id)) for m in self.get_summary_members(parent)}
        found.discard(node_id)
        return list(found)

    def search_by_name_or_title(self, q: str, limit: int = 20) -> List[Tuple[str, Dict[str, Any]]]:
        """Case-insensitive substring match on ``name`` / ``title`` in node order; the cut is checked after every node,
        so ``limit <= 0`` still returns the first match (reference behaviour)."""
        needle, hits = q.lower(), []
        for nid, d in self.G.nodes(data=True):
            if needle in str(d.get("name") or "").lower() or needle in str(d.get("title") or "").lower():
                hits.append((nid, d))
            if len(hits) >= limit:
                break
        return hits

    def summaries_with_top_word(self, word: str, limit: int = 50) -> List[str]:
        needle, hits = word.lower(), []
        for nid, d in self.G.nodes(data=True):
            if d.get("node_type") != "summary":
                continue
            words = d.get("top_words")
            if words:
                if isinstance(words, str):
                    try:
                        words = json.loads(words)
                    except Exception:
                        words = []
                if any(needle in str(w).lower() for w in words):
                    hits.append(nid)
            if len(hits) >= limit:
                break
        return hits

    def validate_belongs_to_dag(self) -> bool:
        """Is the belongs_to subgraph acyclic?  Kahn's algorithm on the exported CSR (no DiGraph copy)."""
        from ..agent.environment import GraphEnvironment
        return GraphEnvironment._belongs_to_is_dag(self._graph_arrays())

    def validate_required_attributes(self) -> Dict[str, List[str]]:
        required = {"entity": ("name",), "hyperedge": ("description", "relation_type")}
        problems: Dict[str, List[str]] = {"entity": [], "hyperedge": [], "summary": []}
        for nid, d in self.G.nodes(data=True):
            kind = d.get("node_type")
            if kind in required:
                if any(not d.get(key) for key in required[kind]):
                    problems[kind].append(nid)
            elif kind == "summary" and d.get("topic_id") is None:
                problems["summary"].append(nid)
        return problems

    def stats(self) -> Dict[str, Any]:
        from collections import Counter
        node_kinds = Counter(d.get("node_type") for _, d in self.G.nodes(data=True))
        edge_kinds = Counter(d.get("edge_type") for _, _, d in self.G.edges(data=True))
        return {"nodes": {k: node_kinds.get(k, 0) for k in ("entity", "hyperedge", "summary")},
                "edges": {k: edge_kinds.get(k, 0) for k in ("participates_in", "belongs_to", "related_to")},
                "n_nodes": self.G.number_of_nodes(), "n_edges": self.G.number_of_edges()}

    # ------------------------------------------------------------------ embeddings (out of scope: pluggable)
    def set_embedder(self, fn: Optional[Embedder]) -> None:
        self.embedder = fn

    def _embed(self, texts: List[str], model_name: str) -> np.ndarray:
        if self.embedder is not None:
            out = np.asarray(self.embedder(list(texts)), dtype=np.float32)
        else:
            try:
                from sentence_transformers import SentenceTransformer  # the reference's embedder (:8, :655, :709)
            except Exception as e:
                raise RuntimeError("embedding-model inference is out of scope of the B200 engine and "
                                   "sentence_transformers is not installed: set HierarchicalGraph.embedder to a "
                                   "callable texts -> [n, d] float32 (L2-normalised)") from e
            if self._st_model is None or self._st_name != model_name:
                self._st_model, self._st_name = SentenceTransformer(model_name), model_name
            out = np.asarray(self._st_model.encode(list(texts), normalize_embeddings=True), dtype=np.float32)
        return out.reshape(len(texts), -1)

    def _embedding_text(self, node_id: str) -> Tuple[str, Dict[str, Any]]:
        """Text + metadata a node is indexed under (hierarchical_graph.py:558-600)."""
        from ..csr import judge_overall
        d = self.G.nodes[node_id]
        nt = d.get("node_type")
        if nt == "entity":
            text = f"Entity: {d.get('name') or ''}. {d.get('description') or ''}"
        elif nt == "summary":
            words: List[str] = []
            if d.get("top_words"):
                try:
                    words = json.loads(d["top_words"]) if isinstance(d["top_words"], str) else d["top_words"]
                except Exception:
                    words = []
            kw = ", ".join(words[:10]) if words else ""
            text = f"Summary: {d.get('title') or ''}. {d.get('summary_text') or d.get('summary') or ''}. Keywords: {kw}"
        else:
            text = f"Relation: {d.get('relation_type') or ''}. {d.get('description') or ''}"
        meta = {"node_id": node_id, "node_type": nt, "layer": 1 if nt == "summary" else 0,
                "judge_overall": judge_overall(d), "confidence": d.get("confidence"),
                "updated_at": datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat() + "Z"}
        return text, meta

    # ------------------------------------------------------------------ vector index (:602-673)
    @staticmethod
    def _index_key(d: Dict[str, Any]) -> str:
        blob = "|".join([str(d.get("node_type")), str(d.get("name") or d.get("title") or ""),
                         str(d.get("summary_text") or ""), str(d.get("description") or "")])
        return hashlib.sha1(blob.encode("utf-8")).hexdigest()

    def build_vector_index(self, db_path: str = "vector_db", layers: set = {0, 1}, include_hyperedges: bool = False,
                           model_name: str = "all-MiniLM-L6-v2", upsert_only: bool = True, reset: bool = False,
                           embeddings: Optional[Dict[str, np.ndarray]] = None) -> None:
        """Same selection / incremental-upsert rules as the reference; the store is ``<db_path>/nodes.npz``
        (ids + fp32 rows + metadata) instead of a Chroma directory.  ``embeddings`` may supply precomputed
        rows by node id (synthetic inputs), bypassing the embedder."""
        from . import vector_store
        os.makedirs(db_path, exist_ok=True)
        store = vector_store.VectorStore.open(db_path, reset=reset)
        vi = self.G.graph.get("vector_index", {})
        prev = vi.get("indexed_meta") if isinstance(vi.get("indexed_meta"), dict) else {}
        ids, docs, metas = [], [], []
        for nid, d in self.G.nodes(data=True):
            nt = d.get("node_type")
            wanted = (nt == "entity" and 0 in layers) or (nt == "summary" and 1 in layers) or \
                     (nt == "hyperedge" and include_hyperedges)
            if not wanted:
                continue
            if upsert_only and prev.get(nid) == self._index_key(d):
                continue
            text, meta = self._embedding_text(nid)
            ids.append(nid); docs.append(text); metas.append(meta)
        if not ids:
            return
        if embeddings is not None:
            rows = np.stack([np.asarray(embeddings[n], dtype=np.float32) for n in ids])
        else:
            rows = self._embed(docs, model_name)
        store.upsert(ids, docs, rows, metas)
        store.save()
        new_meta = dict(prev)
        for nid in ids:
            new_meta[nid] = self._index_key(self.G.nodes[nid])
        self.G.graph["vector_index"] = {"db_path": db_path, "model": model_name, "indexed_nodes": len(new_meta),
                                        "indexed_meta": new_meta}
        self.G.graph["dirty"] = False
        self._dev_index = None

    # ------------------------------------------------------------------ device residency
    _ATTR_FINGERPRINT_MAX_NODES = 2000

    def _graph_sig(self):
        """What the cached device graph is valid for.  ``add_*`` and ``load`` reset the caches themselves; this catches a
        replaced graph object, nodes / edges added behind the class's back and -- for graphs up to
        ``_ATTR_FINGERPRINT_MAX_NODES`` nodes (~0.25 us per node, once per call) -- attributes edited in place through
        ``hg.G.nodes[...]`` (the reference reads G on every call).  Larger graphs: call ``invalidate_device_cache()`` after
        editing node attributes directly."""
        n = len(self.G._node)
        # (DiGraph.number_of_edges() walks a degree view: 50 us on 400 nodes; the successor dicts give the same count)
        sig = (id(self.G), n, sum(map(len, self.G._succ.values())))
        if n <= self._ATTR_FINGERPRINT_MAX_NODES:
            # in-process hash of the four attributes the kernels read (strings cache their hash: ~0.25 us per node; the
            # repr + crc32 form this replaces cost 1.3 us per node and ran four times per search -- 1 ms of a B = 1 call)
            sig += (hash(tuple((d.get("node_type"), d.get("level"), d.get("judge_scores"), d.get("confidence"))
                               for d in self.G._node.values())),)
        return sig

    def _graph_arrays(self) -> Dict[str, Any]:
        """Host copy of the CSR + attribute arrays of the current graph (cached until the graph changes)."""
        sig = self._graph_sig()
        if self._host_arrays is None or self._host_arrays[0] != sig:
            self._host_arrays = (sig, self._export_arrays())
        return self._host_arrays[1]

    def enable_sharding(self, group=None, shard_graph: bool = False, peer_exchange: bool = True) -> None:
        """Multi-GPU form of the SAME API (one process per GPU under ``torchrun``, ``torch.distributed`` initialised, every
        rank holding the same graph directory and vector index): the embedding index is sharded by row range over the ranks
        (K1 on the local rows, one candidate exchange, device merge -- ``dist.ShardedIndex``), and with ``shard_graph`` the
        CSR and the node attributes by node range as well (``dist.ShardedGraph``; otherwise the graph stays replicated).
        ``search`` / ``search_batch`` / ``GraphEnvironment`` / ``NaiveRAG`` then are collectives: every rank calls them with
        the same arguments and gets the same, unsharded-identical result."""
        self._shard = {"group": group, "graph": bool(shard_graph), "peer": bool(peer_exchange)}
        self._dev_graph = None
        self._dev_index = None

    @staticmethod
    def _world(shard: Dict[str, Any]) -> int:
        import torch.distributed as tdist
        if not tdist.is_initialized():
            raise RuntimeError("enable_sharding needs torch.distributed to be initialised (one process per GPU, torchrun)")
        return tdist.get_world_size(shard["group"])

    def _device_graph(self) -> Dict[str, Any]:
        sig = self._graph_sig()
        if self._dev_graph is None or self._dev_graph["sig"] != sig:
            from .. import engine
            arrays = self._graph_arrays()
            shard = getattr(self, "_shard", None)
            if shard is not None and shard["graph"]:
                from .. import dist as dist_mod
                import torch
                dev = torch.device(self.device) if self.device != "cuda" else torch.device("cuda", torch.cuda.current_device())
                csr = dist_mod.ShardedGraphCSR(dist_mod.ShardedGraph.from_arrays(arrays, dev, shard["group"]))
            else:
                csr = engine.GraphCSR(arrays, self.device)
            self._dev_graph = {"sig": sig, "arrays": arrays, "csr": csr}
            self._dev_index = None  # row -> node map depends on dense ids
        return self._dev_graph

    def _export_arrays(self) -> Dict[str, Any]:
        """CSR + attribute arrays of the current graph.  A graph that still is the document ``load`` read comes from
        the binary snapshot beside it (written on first use) or, failing that, from the vectorised node-link export;
        anything built or edited in memory is walked through NetworkX."""
        from .. import csr as csr_mod
        from . import snapshot
        src = self._source
        if src is not None and src.get("sig") == self._graph_sig()[1:]:  # still the document that was loaded
            arrays = snapshot.load(src["dir"], src["key"])
            if arrays is None and src.get("doc") is not None:
                arrays = csr_mod.csr_from_node_link(src["doc"], edges=src["edges"])
                if os.environ.get("AHR_SNAPSHOT", "1") != "0" and snapshot.save(src["dir"], src["key"], arrays):
                    src["doc"] = None  # the snapshot replaces the parsed document from here on
            if arrays is not None and arrays["node_ids"] == list(self.G.nodes):
                return arrays
        return csr_mod.csr_from_networkx(self.G)

    def _device_index(self, db_path: str, dg: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        import torch
        from . import vector_store
        from .. import engine
        dg = dg if dg is not None else self._device_graph()
        stamp = vector_store.VectorStore.stamp(db_path)
        cur = self._dev_index
        if cur is None or cur["db_path"] != db_path or cur["stamp"] != stamp or cur["dtype"] != self.index_dtype:
            store = vector_store.VectorStore.open(db_path)
            rows = torch.from_numpy(np.ascontiguousarray(store.rows, dtype=np.float32))
            dev = dg["csr"].device
            dense = dg["arrays"]["dense"]
            row_to_node = torch.tensor([dense.get(n, -1) for n in store.ids], dtype=torch.int32, device=dev)
            shard = getattr(self, "_shard", None)
            sharded = shard is not None and len(store.ids) >= 2 * self._world(shard)  # (a handful of rows stays replicated)
            lo, hi = 0, len(store.ids)
            if sharded:
                # this rank uploads rows [lo, hi) only; ids stay global (id_base), the merge restores the unsharded order
                from .. import dist as dist_mod
                import torch.distributed as tdist
                lo, hi = dist_mod.shard_bounds(len(store.ids), self._world(shard), tdist.get_rank(shard["group"]))
            rows_dev = rows[lo:hi].to(dev)
            rows_dev = (rows_dev.to(torch.bfloat16) if self.index_dtype == "bf16" else rows_dev).contiguous()
            if sharded:
                index = dist_mod.ShardedIndex.from_rows(rows_dev, id_base=lo, group=shard["group"])
                if shard["peer"]:
                    index.enable_peer_exchange(4096, 128)
            else:
                index = engine.VectorIndex(rows_dev) if len(store.ids) else None
            cur = {"db_path": db_path, "stamp": stamp, "dtype": self.index_dtype, "store": store, "index": index,
                   "row_to_node": row_to_node}
            self._dev_index = cur
        return cur

    def _stage(self, key: str, t):
        """Asynchronous device -> host copy of ``t`` into a pinned buffer kept per output name (grown when needed); valid
        after the caller synchronises the stream, overwritten by the next call."""
        import torch
        pool = self.__dict__.setdefault("_pinned", {})
        buf = pool.get(key)
        n = t.numel()
        if buf is None or buf.dtype != t.dtype or buf.numel() < n:
            buf = pool[key] = torch.empty((max(n, 64),), dtype=t.dtype).pin_memory()
        view = buf[:n].view(t.shape)
        view.copy_(t, non_blocking=True)
        return view

    def invalidate_device_cache(self) -> None:
        """Call after editing nodes / edges / attributes behind the class's back (``hg.G...``): the next search exports the
        graph again by walking NetworkX (the on-disk snapshot describes the document as loaded, not the edited graph)."""
        self._dev_graph = None
        self._host_arrays = None
        self._dev_index = None
        self._source = None

    # ------------------------------------------------------------------ search (:675-789)
    def _resolve(self, member_top_m, alpha, beta, gamma, delta, judge_overall_min, confidence_min, type_filter):
        from ..engine import SearchParams
        sp = self.G.graph.get("search_params", {})
        if type_filter is None:
            tf = sp.get("type_filter")
            type_filter = list(tf) if isinstance(tf, (set, tuple)) else tf
        return SearchParams(
            alpha=alpha if alpha is not None else sp.get("alpha", 0.6),
            beta=beta if beta is not None else sp.get("beta", 0.2),
            gamma=gamma if gamma is not None else sp.get("gamma", 0.1),
            delta=delta if delta is not None else sp.get("delta", 0.1),
            judge_overall_min=judge_overall_min if judge_overall_min is not None else sp.get("judge_overall_min"),
            confidence_min=confidence_min if confidence_min is not None else sp.get("confidence_min"),
            type_filter=type_filter,
            member_top_m=member_top_m if member_top_m is not None else sp.get("member_top_m", 5),
            layer_boost=sp.get("layer_boost", {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0}))

    def search(self, query: str, top_k: int = 5, member_top_m: int = 5,
               alpha: Optional[float] = None, beta: Optional[float] = None, gamma: Optional[float] = None,
               delta: Optional[float] = None, judge_overall_min: Optional[float] = None,
               confidence_min: Optional[float] = None, type_filter: Optional[List[str]] = None,
               return_cluster: bool = False, db_path: Optional[str] = None,
               model_name: Optional[str] = None) -> Union[List[Dict[str, Any]], Dict[str, Any]]:
        """Single-query API of the reference; one-row batch through :meth:`search_batch`."""
        db_path, model_name = self._resolve_index(db_path, model_name)
        q = self._embed([query], model_name)
        return self.search_batch(q, top_k=top_k, member_top_m=member_top_m, alpha=alpha, beta=beta, gamma=gamma,
                                 delta=delta, judge_overall_min=judge_overall_min, confidence_min=confidence_min,
                                 type_filter=type_filter, return_cluster=return_cluster, db_path=db_path)[0]

    def _resolve_index(self, db_path: Optional[str], model_name: Optional[str]) -> Tuple[str, str]:
        if not db_path or not model_name:
            vi = self.G.graph.get("vector_index", {}) if isinstance(self.G.graph.get("vector_index"), dict) else {}
            db_path = db_path or vi.get("db_path", "vector_db")
            model_name = model_name or vi.get("model", "all-MiniLM-L6-v2")
        return db_path, model_name

    def _device_search(self, query_embeddings, top_k: int, params, want_cluster: bool, db_path: str,
                       dg: Optional[Dict[str, Any]] = None, di: Optional[Dict[str, Any]] = None):
        """Top-k + anchor expansion/rerank for a batch, results left on the device (no host synchronisation).
        -> (dict of device tensors from ``GraphCSR.anchor_rerank``, index rows [B, k_eff] of the seeds)."""
        import torch
        dg = dg if dg is not None else self._device_graph()
        di = di if di is not None else self._device_index(db_path, dg)
        csr = dg["csr"]
        dev = csr.device
        q = query_embeddings if isinstance(query_embeddings, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(query_embeddings, dtype=np.float32))
        k_eff = min(int(top_k), len(di["store"].ids))
        if k_eff > 1024:
            raise ValueError(f"top_k = {top_k} over an index of {len(di['store'].ids)} rows: the device top-k returns at most 1024 seeds per query")
        if self.index_dtype == "bf16":
            qd = q.to(dev).to(torch.bfloat16).contiguous()
            # the tensor-core lists hold up to 120 entries; wider requests take the exact scan of the same bf16 rows
            scores, rows = di["index"].topk(qd, k_eff, mode="bf16" if k_eff <= 120 else "exact")
        else:
            qd = q.to(dev).to(torch.float32).contiguous()
            scores, rows = di["index"].topk(qd, k_eff, mode="exact")
        # the index rows go straight into the expansion kernel: the row -> node lookup happens there
        out = csr.anchor_rerank(rows, scores, params, int(top_k), want_cluster=want_cluster, row_to_node=di["row_to_node"])
        return out, rows

    def search_batch(self, query_embeddings, top_k: int = 5, member_top_m: int = 5,
                     alpha: Optional[float] = None, beta: Optional[float] = None, gamma: Optional[float] = None,
                     delta: Optional[float] = None, judge_overall_min: Optional[float] = None,
                     confidence_min: Optional[float] = None, type_filter: Optional[List[str]] = None,
                     return_cluster: bool = False, db_path: Optional[str] = None) -> List[Any]:
        """Batched search: ``query_embeddings`` [B, d] (numpy or torch, host or device).  Per item the result
        equals ``search`` on that query.  One fused top-k launch + one expansion/rerank launch for the batch."""
        params = self._resolve(member_top_m, alpha, beta, gamma, delta, judge_overall_min, confidence_min, type_filter)
        db_path, _ = self._resolve_index(db_path, "unused")
        dg = self._device_graph()  # (the graph signature is evaluated once per call)
        di = self._device_index(db_path, dg)
        arrays, store = dg["arrays"], di["store"]
        B = len(query_embeddings)
        n_index = len(store.ids)
        if B == 0:
            return []
        if n_index == 0 or top_k <= 0:
            empty = {"seeds": [], "expanded": [], "reranked": []}
            return [dict(empty) if return_cluster else [] for _ in range(B)]
        k_eff = min(int(top_k), n_index)  # Chroma returns at most the collection size
        out, rows = self._device_search(query_embeddings, int(top_k), params, return_cluster, db_path, dg, di)
        # every output into its pinned staging buffer without waiting, ONE synchronisation for the lot (a `.cpu()` per
        # tensor is a 16 us round trip each), then plain Python lists (numpy scalar indexing would dominate the loop)
        import torch
        staged = {k: self._stage(k, v) for k, v in out.items()}
        rows_st = self._stage("rows", rows) if return_cluster else None
        torch.cuda.current_stream(rows.device).synchronize()
        host = {k: v.numpy().tolist() for k, v in staged.items()}
        rows_h = rows_st.numpy().tolist() if return_cluster else None
        node_ids, nodes_view = arrays["node_ids"], self.G.nodes
        judge_arr = arrays["judge"]  # float(js["overall"]) per node, NaN = None: the value `search` re-parses per result
        results: List[Any] = []
        for b in range(B):
            reranked = []
            h_nodes, h_sem, h_score = host["nodes"][b], host["semantic"][b], host["score"][b]
            for j in range(host["count"][b]):
                v = h_nodes[j]
                nid = node_ids[v]
                d = nodes_view.get(nid, {})
                nt = d.get("node_type")
                ju = judge_arr[v]
                reranked.append({"node_id": nid, "node_type": nt, "layer": 1 if nt == "summary" else 0,
                                 "semantic": h_sem[j], "judge_overall": None if ju != ju else float(ju),
                                 "confidence": d.get("confidence"), "score": h_score[j],
                                 "name": d.get("name"), "title": d.get("title")})
            if not return_cluster:
                results.append(reranked)
                continue
            seeds, seed_meta = [], {}
            seed_sem = host["seed_semantic"][b]
            for r in range(k_eff):
                row = rows_h[b][r]
                if row < 0:
                    continue
                nid = store.ids[row]
                seeds.append({"node_id": nid, "semantic": seed_sem[r], "meta": store.metas[row]})
                seed_meta[nid] = store.metas[row]
            expanded = []
            e_nodes, e_sem, e_kind = host["exp_nodes"][b], host["exp_semantic"][b], host["exp_kind"][b]
            for j in range(host["exp_count"][b]):
                nid = node_ids[e_nodes[j]]
                kind = e_kind[j]
                if kind == 1:
                    expanded.append({"node_id": nid, "semantic": e_sem[j], "meta": seed_meta.get(nid),
                                     "node_type": nodes_view.get(nid, {}).get("node_type")})
                else:
                    expanded.append({"node_id": nid, "semantic": e_sem[j], "meta": {"expanded": True},
                                     "node_type": "summary" if kind == 2 else "entity"})
            results.append({"seeds": seeds, "expanded": expanded, "reranked": reranked})
        return results
