"""sys.modules stand-ins that let the UNMODIFIED reference run in this container (TEST INFRASTRUCTURE).

``import ah_rag.graph.hierarchical_graph`` needs ``sentence_transformers`` (hierarchical_graph.py:8)
and ``chromadb`` / ``chromadb.config`` (:12-13), neither of which is installed or installable here.
``install()`` registers minimal replacements exposing exactly the calls the hot path makes:

  chromadb.PersistentClient(path, settings)              hierarchical_graph.py:605, :706
  client.delete_collection(name)                         :608
  client.get_or_create_collection(name, metadata)        :611, :707
  Collection.upsert(ids, documents, embeddings, metadatas)   :660
  Collection.query(query_embeddings, n_results, include) :714
  SentenceTransformer(name).encode(texts, normalize_embeddings=True, show_progress_bar=...)  :655-656, :712

``Collection.query`` is the exact-cosine contract of ``oracle/exact.py`` (chromadb's own HNSW search is
approximate and un-pinned: parity at that seam is UNPINNED).  The embedder is a deterministic
text -> unit-vector hash (embedding-model inference is out of scope); ``EMBED_OVERRIDE`` lets a test
pin the vector a given text maps to.
"""
from __future__ import annotations

import hashlib
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np

from . import exact

EMBED_DIM = 384
EMBED_OVERRIDE: Dict[str, np.ndarray] = {}
_STORE: Dict[str, Dict[str, "Collection"]] = {}


def hash_embed(text: str, dim: int = EMBED_DIM) -> np.ndarray:
    """Deterministic unit vector for a text (stands in for all-MiniLM-L6-v2, out of scope)."""
    if text in EMBED_OVERRIDE:
        return np.asarray(EMBED_OVERRIDE[text], dtype=np.float32)
    seed = int.from_bytes(hashlib.sha1(text.encode("utf-8")).digest()[:8], "little")
    v = np.random.default_rng(seed).standard_normal(dim).astype(np.float32)
    v /= np.float32(np.linalg.norm(v.astype(np.float64)))
    return v.astype(np.float32)


class SentenceTransformer:
    def __init__(self, model_name: str = "all-MiniLM-L6-v2", *a: Any, **k: Any) -> None:
        self.model_name = model_name

    def encode(self, texts, show_progress_bar: bool = False, normalize_embeddings: bool = False, **k: Any):
        if isinstance(texts, str):
            return hash_embed(texts)
        return np.stack([hash_embed(t) for t in texts]) if len(texts) else np.zeros((0, EMBED_DIM), np.float32)


class Collection:
    def __init__(self, name: str, metadata: Optional[dict]) -> None:
        self.name = name
        self.metadata = metadata or {}
        self._rows: Dict[str, int] = {}
        self._ids: List[str] = []
        self._emb: List[np.ndarray] = []
        self._docs: List[Any] = []
        self._metas: List[Any] = []

    def count(self) -> int:
        return len(self._ids)

    def upsert(self, ids, documents=None, embeddings=None, metadatas=None) -> None:
        for j, nid in enumerate(ids):
            e = np.asarray(embeddings[j], dtype=np.float32)
            doc = documents[j] if documents is not None else None
            meta = metadatas[j] if metadatas is not None else None
            if nid in self._rows:
                r = self._rows[nid]
                self._emb[r], self._docs[r], self._metas[r] = e, doc, meta
            else:
                self._rows[nid] = len(self._ids)
                self._ids.append(nid)
                self._emb.append(e)
                self._docs.append(doc)
                self._metas.append(meta)

    def matrix(self) -> np.ndarray:
        return np.stack(self._emb).astype(np.float32) if self._emb else np.zeros((0, EMBED_DIM), np.float32)

    def query(self, query_embeddings, n_results: int = 10, include=None) -> Dict[str, Any]:
        out: Dict[str, Any] = {"ids": [], "distances": [], "metadatas": [], "documents": []}
        X = self.matrix()
        for q in query_embeddings:
            if X.shape[0] == 0:
                for key in out:
                    out[key].append([])
                continue
            qv = np.asarray(q, dtype=np.float32)[None, :]
            k = min(int(n_results), X.shape[0])
            scores, rows = exact.exact_topk_np(X, qv, k)
            dist = (np.float32(1.0) - scores[0]).astype(np.float32)  # hnswlib cosine space: 1.0f - <q, x>
            out["ids"].append([self._ids[r] for r in rows[0]])
            out["distances"].append([float(d) for d in dist])
            out["metadatas"].append([self._metas[r] for r in rows[0]])
            out["documents"].append([self._docs[r] for r in rows[0]])
        return out


class PersistentClient:
    def __init__(self, path: str = "vector_db", settings: Any = None, **k: Any) -> None:
        self.path = str(path)
        _STORE.setdefault(self.path, {})

    def delete_collection(self, name: str) -> None:
        if name not in _STORE[self.path]:
            raise ValueError(f"Collection {name} does not exist.")
        del _STORE[self.path][name]

    def get_or_create_collection(self, name: str, metadata: Optional[dict] = None, **k: Any) -> Collection:
        colls = _STORE[self.path]
        if name not in colls:
            colls[name] = Collection(name, metadata)
        return colls[name]


class Settings:
    def __init__(self, *a: Any, **k: Any) -> None:
        self.kwargs = k


def reset_store() -> None:
    _STORE.clear()
    EMBED_OVERRIDE.clear()


def install() -> None:
    """Register the stand-ins (idempotent).  Never overrides a real installation."""
    if "sentence_transformers" not in sys.modules:
        m = types.ModuleType("sentence_transformers")
        m.SentenceTransformer = SentenceTransformer
        sys.modules["sentence_transformers"] = m
    if "chromadb" not in sys.modules:
        c = types.ModuleType("chromadb")
        c.PersistentClient = PersistentClient
        c.Collection = Collection
        cfg = types.ModuleType("chromadb.config")
        cfg.Settings = Settings
        c.config = cfg
        sys.modules["chromadb"] = c
        sys.modules["chromadb.config"] = cfg


def import_reference(ref_root: str = "/root/reference"):
    """Import the reference's hot-path modules unmodified; returns (HierarchicalGraph, GraphEnvironment,
    AHRAGGymEnv, NaiveRAG).  Only usable where /root/reference exists (the dev container)."""
    import importlib
    import os

    install()
    src = os.path.join(ref_root, "src")
    for p in (src, ref_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    hg = importlib.import_module("ah_rag.graph.hierarchical_graph")
    env = importlib.import_module("ah_rag.agent.environment")
    gym = importlib.import_module("ah_rag.agent.gym_env")
    nr = importlib.import_module("baselines.naive_rag")
    return hg.HierarchicalGraph, env.GraphEnvironment, gym.AHRAGGymEnv, nr.NaiveRAG


def import_reference_aggregator(ref_root: str = "/root/reference"):
    """``SemanticAggregator`` of the unmodified reference (semantic_aggregator.py) for the one method that is pure
    numpy / Python, ``generate_l1_relations`` (:594-680).  The clustering libraries the module imports at the top
    (bertopic, umap, hdbscan) are absent here and not needed by that method: empty stand-ins."""
    import importlib
    import os

    install()
    for name, attr in (("bertopic", "BERTopic"), ("umap", "UMAP"), ("hdbscan", "HDBSCAN")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, type(attr, (), {}))
            sys.modules[name] = m
    src = os.path.join(ref_root, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    return importlib.import_module("ah_rag.aggregate.semantic_aggregator").SemanticAggregator
