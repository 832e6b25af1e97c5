"""On-disk embedding store replacing the Chroma persistent directory of the reference
(hierarchical_graph.py:605-611, :660: PersistentClient / collection "nodes" / upsert).

Layout: ``<db_path>/nodes.npz`` (fp32 rows, row order = first-upsert order, as a Chroma collection keeps
ids) and ``<db_path>/nodes.json`` (ids, documents, metadatas).  Host-side only; the device-resident copy is
``engine.VectorIndex``.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional

import numpy as np


class VectorStore:
    def __init__(self, db_path: str) -> None:
        self.db_path = db_path
        self.ids: List[str] = []
        self.docs: List[Any] = []
        self.metas: List[Any] = []
        self.rows: np.ndarray = np.zeros((0, 0), dtype=np.float32)
        self._pos: Dict[str, int] = {}

    @staticmethod
    def _files(db_path: str):
        return os.path.join(db_path, "nodes.npz"), os.path.join(db_path, "nodes.json")

    @staticmethod
    def stamp(db_path: str) -> Optional[int]:
        npz, _ = VectorStore._files(db_path)
        return os.stat(npz).st_mtime_ns if os.path.exists(npz) else None

    @classmethod
    def open(cls, db_path: str, reset: bool = False) -> "VectorStore":
        st = cls(db_path)
        npz, js = cls._files(db_path)
        if reset:
            for p in (npz, js):
                if os.path.exists(p):
                    os.remove(p)
            return st
        if os.path.exists(npz) and os.path.exists(js):
            st.rows = np.load(npz)["rows"].astype(np.float32)
            with open(js, "r", encoding="utf-8") as f:
                meta = json.load(f)
            st.ids, st.docs, st.metas = meta["ids"], meta["documents"], meta["metadatas"]
            st._pos = {n: i for i, n in enumerate(st.ids)}
        return st

    def upsert(self, ids: List[str], documents: List[Any], rows: np.ndarray, metadatas: List[Any]) -> None:
        rows = np.asarray(rows, dtype=np.float32)
        if self.rows.size == 0:
            self.rows = np.zeros((0, rows.shape[1]), dtype=np.float32)
        new_rows = []
        for j, nid in enumerate(ids):
            if nid in self._pos:
                r = self._pos[nid]
                self.rows[r] = rows[j]
                self.docs[r], self.metas[r] = documents[j], metadatas[j]
            else:
                self._pos[nid] = len(self.ids)
                self.ids.append(nid)
                self.docs.append(documents[j])
                self.metas.append(metadatas[j])
                new_rows.append(rows[j])
        if new_rows:
            self.rows = np.concatenate([self.rows, np.stack(new_rows)], axis=0)

    def save(self) -> None:
        os.makedirs(self.db_path, exist_ok=True)
        npz, js = self._files(self.db_path)
        np.savez(npz, rows=self.rows)
        with open(js, "w", encoding="utf-8") as f:
            json.dump({"ids": self.ids, "documents": self.docs, "metadatas": self.metas}, f, ensure_ascii=False)
