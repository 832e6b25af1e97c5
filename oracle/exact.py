"""Exact similarity + top-k oracle (TEST INFRASTRUCTURE, not product code).

Restates the query the reference issues at ``hierarchical_graph.py:714`` against its cosine-space
Chroma collection (``:611``) as an exact scan -- see the header of ``oracle/exact_topk.c`` for the
contract (canonical order-fixed fp32 score, ranking ``(score desc, row id asc)``) and for why parity
at this seam is *unpinned* (chromadb / chroma-hnswlib absent and un-versioned in the reference).

Two implementations that must agree bit-for-bit (``tests/test_oracle.py``):
  * ``*_np``  -- numpy, sequential over the feature dimension (small inputs);
  * C library ``oracle/liboracle_exact.so`` built by ``oracle/Makefile`` (OpenMP, any size).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_exact.so")
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc only; no GPU, no reference sources)."""
    src = os.path.join(_HERE, "exact_topk.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_exact.so"])
    return _LIB_PATH


def _load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        lib.ahro_canonical_scores.argtypes = [vp, vp, i64, i32, i32, ctypes.c_int, vp]
        lib.ahro_exact_topk.argtypes = [vp, vp, i64, i32, i32, ctypes.c_int, i32, i64, vp, vp]
        lib.ahro_exact_topk_rowpar.argtypes = [vp, vp, i64, i32, i32, ctypes.c_int, i32, i64, i32, vp, vp]
        lib.ahro_rescore.argtypes = [vp, vp, i32, i32, ctypes.c_int, i32, vp, vp]
        lib.ahro_round4.argtypes = [ctypes.c_double]
        lib.ahro_round4.restype = ctypes.c_double
        lib.ahro_round4_array.argtypes = [vp, i64, vp]
        _lib = lib
    return _lib


# ---------------------------------------------------------------------------------------------
# dtype helpers: rows are fp32 arrays, or bf16 carried as raw uint16 (numpy has no bfloat16)
# ---------------------------------------------------------------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as raw uint16 (same rule as torch's cast)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def _as_f64(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.uint16:
        return bf16_bits_to_f32(a).astype(np.float64)
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def _dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.uint16:
        return 1
    raise TypeError(f"rows must be float32 or bf16-as-uint16, got {a.dtype}")


# ---------------------------------------------------------------------------------------------
# numpy restatement
# ---------------------------------------------------------------------------------------------
def canonical_scores_np(X: np.ndarray, Q: np.ndarray) -> np.ndarray:
    """``float32(sum_j double(q_j) * double(x_j))`` with j ascending, for all pairs -> [nq, n]."""
    Xd, Qd = _as_f64(X), _as_f64(Q)
    acc = np.zeros((Qd.shape[0], Xd.shape[0]), dtype=np.float64)
    for j in range(Xd.shape[1]):  # sequential over the feature dimension == the canonical order
        acc += Qd[:, j : j + 1] * Xd[:, j][None, :]
    return acc.astype(np.float32)


def topk_from_scores(S: np.ndarray, k: int, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Rank rows by ``(score desc, id asc)``; pad with (-inf, -1) when fewer than k rows."""
    nq, n = S.shape
    scores = np.full((nq, k), -np.inf, dtype=np.float32)
    ids = np.full((nq, k), -1, dtype=np.int64)
    rows = np.arange(n, dtype=np.int64)
    for qi in range(nq):
        order = np.lexsort((rows, -S[qi].astype(np.float64)))[:k]  # last key is primary
        scores[qi, : order.size] = S[qi, order]
        ids[qi, : order.size] = order + id_base
    return scores, ids


def exact_topk_np(X: np.ndarray, Q: np.ndarray, k: int, id_base: int = 0):
    return topk_from_scores(canonical_scores_np(X, Q), k, id_base)


# ---------------------------------------------------------------------------------------------
# C restatement
# ---------------------------------------------------------------------------------------------
def canonical_scores(X: np.ndarray, Q: np.ndarray) -> np.ndarray:
    lib = _load()
    X, Q = np.ascontiguousarray(X), np.ascontiguousarray(Q)
    assert X.dtype == Q.dtype and X.shape[1] == Q.shape[1]
    out = np.empty((Q.shape[0], X.shape[0]), dtype=np.float32)
    rc = lib.ahro_canonical_scores(X.ctypes.data, Q.ctypes.data, X.shape[0], X.shape[1], Q.shape[0],
                                   _dtype_code(X), out.ctypes.data)
    assert rc == 0
    return out


def exact_topk(X: np.ndarray, Q: np.ndarray, k: int, id_base: int = 0):
    lib = _load()
    X, Q = np.ascontiguousarray(X), np.ascontiguousarray(Q)
    assert X.dtype == Q.dtype and X.shape[1] == Q.shape[1]
    scores = np.empty((Q.shape[0], k), dtype=np.float32)
    ids = np.empty((Q.shape[0], k), dtype=np.int64)
    rc = lib.ahro_exact_topk(X.ctypes.data, Q.ctypes.data, X.shape[0], X.shape[1], Q.shape[0],
                             _dtype_code(X), k, id_base, scores.ctypes.data, ids.ctypes.data)
    assert rc == 0
    return scores, ids


def exact_topk_rowpar(X: np.ndarray, Q: np.ndarray, k: int, id_base: int = 0, n_threads: int = 0):
    """``exact_topk`` parallelised over row blocks (few queries over a large corpus); same result bit for bit.
    ``n_threads`` is explicit because launchers such as torchrun export OMP_NUM_THREADS=1."""
    lib = _load()
    X, Q = np.ascontiguousarray(X), np.ascontiguousarray(Q)
    assert X.dtype == Q.dtype and X.shape[1] == Q.shape[1]
    scores = np.empty((Q.shape[0], k), dtype=np.float32)
    ids = np.empty((Q.shape[0], k), dtype=np.int64)
    nt = int(n_threads) if n_threads > 0 else len(os.sched_getaffinity(0))
    rc = lib.ahro_exact_topk_rowpar(X.ctypes.data, Q.ctypes.data, X.shape[0], X.shape[1], Q.shape[0],
                                    _dtype_code(X), k, id_base, nt, scores.ctypes.data, ids.ctypes.data)
    assert rc == 0
    return scores, ids


def rescore(X: np.ndarray, Q: np.ndarray, cand_ids: np.ndarray) -> np.ndarray:
    lib = _load()
    X, Q = np.ascontiguousarray(X), np.ascontiguousarray(Q)
    cand = np.ascontiguousarray(cand_ids, dtype=np.int64)
    out = np.empty(cand.shape, dtype=np.float32)
    rc = lib.ahro_rescore(X.ctypes.data, Q.ctypes.data, X.shape[1], Q.shape[0], _dtype_code(X),
                          cand.shape[1], cand.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def merge_topk(scores_parts, ids_parts, k: int):
    """Merge per-shard candidate lists with the same comparator (what the device k-way merge does)."""
    S = np.concatenate(scores_parts, axis=1)
    I = np.concatenate(ids_parts, axis=1)
    nq = S.shape[0]
    out_s = np.full((nq, k), -np.inf, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        valid = I[qi] >= 0
        s, i = S[qi][valid], I[qi][valid]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        out_s[qi, : order.size] = s[order]
        out_i[qi, : order.size] = i[order]
    return out_s, out_i


def round4(x: float) -> float:
    """C restatement of CPython's ``round(x, 4)`` (the device routine's twin)."""
    return float(_load().ahro_round4(float(x)))


def round4_array(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _load().ahro_round4_array(x.ctypes.data, x.size, out.ctypes.data)
    return out
