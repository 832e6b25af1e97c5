"""Runs the UNMODIFIED reference hot path in the dev container (TEST INFRASTRUCTURE).

Needs /root/reference, so it is used only by ``oracle/make_golden.py`` (to write ``tests/golden/``) and
by ``tests/test_oracle_vs_reference.py`` (skipped where the reference tree is absent, e.g. the GPU box).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import tempfile
from typing import Any, Dict, List, Sequence

import numpy as np

from . import stubs

REF_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "ah_rag"))


class ReferenceSession:
    """A reference HierarchicalGraph + vector index built from a node-link dict."""

    def __init__(self, node_link: Dict[str, Any], workdir: str | None = None) -> None:
        self.HG, self.Env, self.Gym, self.Naive = stubs.import_reference(REF_ROOT)
        self.tmp = workdir or tempfile.mkdtemp(prefix="ahrag_ref_")
        self.graph_dir = os.path.join(self.tmp, "graph")
        self.db_path = os.path.join(self.tmp, "vector_db")
        os.makedirs(self.graph_dir, exist_ok=True)
        nl = json.loads(json.dumps(node_link))
        nl["graph"].setdefault("vector_index", {})
        nl["graph"]["vector_index"].update({"db_path": self.db_path, "model": "all-MiniLM-L6-v2"})
        with open(os.path.join(self.graph_dir, "structure.json"), "w", encoding="utf-8") as f:
            json.dump(nl, f, ensure_ascii=False)
        meta = {"search_params": nl["graph"].get("search_params", {}), "dirty": False,
                "vector_index": {"db_path": self.db_path, "model": "all-MiniLM-L6-v2",
                                 "indexed_nodes": 0, "indexed_meta": {}}}
        with open(os.path.join(self.graph_dir, "meta.json"), "w", encoding="utf-8") as f:
            json.dump(meta, f)
        self.hg = self.HG.load(self.graph_dir)
        with contextlib.redirect_stdout(io.StringIO()):
            # same call as scripts/demo_cli.py:115
            self.hg.build_vector_index(self.db_path, layers={0, 1, 2}, reset=True)
        coll = stubs._STORE[self.db_path]["nodes"]
        self.index_ids: List[str] = list(coll._ids)
        self.index_rows: np.ndarray = coll.matrix()

    def search(self, query: str, **kw) -> Dict[str, Any]:
        return self.hg.search(query, return_cluster=True, db_path=self.db_path, **kw)

    def make_env(self):
        return self.Env(graph_dir=self.graph_dir, logging_enabled=False)

    def make_gym(self, max_steps: int = 6):
        cwd = os.getcwd()
        os.chdir(self.tmp)  # gym env leaves logging on and writes artifacts/phase2/<session> (gym_env.py:27)
        try:
            return self.Gym(graph_dir=self.graph_dir, max_steps=max_steps)
        finally:
            os.chdir(cwd)


def strip_search(out: Dict[str, Any]) -> Dict[str, Any]:
    """Drop fields that are not results (Chroma metadata with timestamps)."""
    return {
        "seeds": [{"node_id": s["node_id"], "semantic": s["semantic"]} for s in out["seeds"]],
        "expanded": [{"node_id": e["node_id"], "semantic": e["semantic"], "node_type": e["node_type"]}
                     for e in out["expanded"]],
        "reranked": out["reranked"],
    }


def reference_l1_relations(l1_nodes: List[Dict[str, Any]], **kw) -> List[Dict[str, Any]]:
    """``SemanticAggregator.generate_l1_relations`` of the unmodified reference (no model is loaded: the method does
    not touch ``self``).  Runs in a scratch directory because it writes ``artifacts/l1_*.json``; mutates ``l1_nodes``
    (``neighbors``) exactly as the reference does."""
    cls = stubs.import_reference_aggregator(REF_ROOT)
    agg = cls.__new__(cls)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="ahrag_ref_l1_"))
    try:
        return agg.generate_l1_relations(l1_nodes, **kw)
    finally:
        os.chdir(cwd)
