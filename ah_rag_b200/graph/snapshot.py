"""Binary device snapshot of a loaded graph: the CSR + attribute arrays ``ahr_graph_create`` takes, cached next to
``structure.json`` so that the second process (or the other 7 ranks of a box) skips the export altogether.

SURVEY.md section 8 (f2): the reference re-parses the node-link JSON and re-walks NetworkX in every environment
(environment.py:73-77, hierarchical_graph.py:191-218).  The snapshot is keyed by the SHA-1 of the ``structure.json``
bytes it was built from (not by ``graph_hash`` (:220-236), which ignores edges) and by a format version; a stale or
foreign file is ignored, a read-only directory is not an error.

Layout: one uncompressed ``.npz``: ``row_ptr0..5`` int64, ``col_idx0..5`` int32, ``node_type`` / ``level`` u8,
``judge`` / ``judge_term`` / ``conf`` f64, ``node_ids`` (JSON text, uint8), ``key`` (ASCII, uint8).
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Any, Dict, Optional

import numpy as np

FORMAT_VERSION = 1
FILE_NAME = "device_snapshot.npz"


def key_of(structure_bytes: bytes) -> str:
    return f"v{FORMAT_VERSION}:" + hashlib.sha1(structure_bytes).hexdigest()


def path_in(directory: str) -> str:
    return os.path.join(directory, FILE_NAME)


def save(directory: str, key: str, arrays: Dict[str, Any]) -> bool:
    """Best effort (atomic rename); returns False when the directory cannot be written."""
    try:
        payload = {f"row_ptr{s}": np.ascontiguousarray(arrays["row_ptr"][s], dtype=np.int64) for s in range(6)}
        payload.update({f"col_idx{s}": np.ascontiguousarray(arrays["col_idx"][s], dtype=np.int32) for s in range(6)})
        for name, dt in (("node_type", np.uint8), ("level", np.uint8), ("judge", np.float64), ("judge_term", np.float64),
                         ("conf", np.float64)):
            payload[name] = np.ascontiguousarray(arrays[name], dtype=dt)
        payload["node_ids"] = np.frombuffer(json.dumps(arrays["node_ids"], ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
        payload["key"] = np.frombuffer(key.encode("ascii"), dtype=np.uint8)
        tmp = path_in(directory) + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            np.savez(f, **payload)
        os.replace(tmp, path_in(directory))
        return True
    except OSError:
        return False


def load(directory: str, key: str) -> Optional[Dict[str, Any]]:
    """The arrays of ``csr.csr_from_node_link`` if a snapshot with this key exists, else None."""
    p = path_in(directory)
    if not os.path.exists(p):
        return None
    try:
        with np.load(p, allow_pickle=False) as z:
            if bytes(z["key"]).decode("ascii") != key:
                return None
            node_ids = json.loads(bytes(z["node_ids"]).decode("utf-8"))
            out = {"n_nodes": len(node_ids), "node_ids": node_ids, "dense": {n: i for i, n in enumerate(node_ids)},
                   "row_ptr": [z[f"row_ptr{s}"] for s in range(6)], "col_idx": [z[f"col_idx{s}"] for s in range(6)]}
            for name in ("node_type", "level", "judge", "judge_term", "conf"):
                out[name] = z[name]
        if any(len(rp) != out["n_nodes"] + 1 for rp in out["row_ptr"]):
            return None
        return out
    except (OSError, KeyError, ValueError):
        return None
