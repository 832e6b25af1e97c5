"""ctypes binding of libahrag_b200.so (the C ABI declared in include/ahrag_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built, importing this module
raises.  The library itself refuses to run on anything but an sm_100 device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

from . import _build

AHR_DTYPE_F32, AHR_DTYPE_BF16 = 0, 1
AHR_MODE_BF16, AHR_MODE_EXACT = 0, 1
AHR_N_CSR = 6
CSR_PART_OUT, CSR_PART_IN, CSR_BEL_OUT, CSR_BEL_IN, CSR_REL_OUT, CSR_REL_IN = range(6)
NT_ENTITY, NT_SUMMARY, NT_HYPEREDGE, NT_OTHER = 0, 1, 2, 3
EK_SEED, EK_PARENT, EK_MEMBER = 1, 2, 3
OP_PARENTS, OP_CHILDREN, OP_RELATED = 0, 1, 2

ABI_VERSION = 200  # AHR_VERSION of include/ahrag_b200.h this binding was written against

# every symbol include/ahrag_b200.h declares (tests/test_abi.py checks the header against this list)
AHR_IPC_HANDLE_BYTES = 64

EXPORTS = [
    "ahr_last_error", "ahr_version", "ahr_launch_count", "ahr_index_create", "ahr_index_destroy", "ahr_topk",
    "ahr_topk_exact_begin", "ahr_topk_exact_finish",
    "ahr_topk_host", "ahr_merge_topk", "ahr_merge_topk_packed", "ahr_graph_create", "ahr_graph_destroy", "ahr_anchor_rerank", "ahr_anchor_rerank_rows", "ahr_expand",
    "ahr_debug_tc_scores", "ahr_rescore", "ahr_debug_round4", "ahr_exact_stats", "ahr_set_profiling",
    "ahr_get_profile", "ahr_env_reset", "ahr_env_step", "ahr_comm_create", "ahr_comm_export", "ahr_comm_connect",
    "ahr_comm_destroy", "ahr_comm_begin", "ahr_comm_exchange_merge", "ahr_topk_sharded", "ahr_lca_closure",
    "ahr_graph_shard_create", "ahr_graph_shard_destroy", "ahr_expand_shard_record_ints", "ahr_expand_shard_stage1",
    "ahr_expand_shard_stage2", "ahr_expand_shard_merge", "ahr_graph_shard_set_attrs", "ahr_anchor_shard_record_ints",
    "ahr_anchor_shard_stage1", "ahr_anchor_shard_stage2", "ahr_anchor_shard_merge",
]


class RerankParams(ctypes.Structure):
    """Mirror of ``struct ahr_rerank_params``."""
    _fields_ = [
        ("alpha", c_double), ("beta", c_double), ("gamma", c_double), ("delta", c_double),
        ("boost", c_double * 4),
        ("judge_min", c_double), ("conf_min", c_double),
        ("has_judge_min", c_int32), ("has_conf_min", c_int32),
        ("type_mask", c_int32), ("member_top_m", c_int32), ("parent_cap", c_int32), ("reserved", c_int32),
    ]


class EnvBatch(ctypes.Structure):
    """Mirror of ``struct ahr_env_batch`` (device pointers of the batched environment state)."""
    _fields_ = [("n_env", c_int32), ("cap_sel", c_int32), ("cap_fr", c_int32), ("reserved", c_int32)] + [
        (name, c_void_p) for name in ("sel_ids", "sel_cnt", "fr_ids", "fr_cnt", "shown_ids", "shown_score", "shown_sem",
                                      "shown_cnt", "n_seeds", "obs_step", "obs_sel", "obs_fr", "env_steps", "gym_steps",
                                      "last_action", "done", "overflow")]


def _declare(lib: ctypes.CDLL) -> None:
    vp = c_void_p
    lib.ahr_last_error.restype = c_char_p
    lib.ahr_last_error.argtypes = []
    lib.ahr_version.restype = c_int
    lib.ahr_launch_count.restype = c_int64
    lib.ahr_index_create.argtypes = [POINTER(vp), c_int, c_int64, c_int32, c_int, vp, c_int64]
    lib.ahr_index_destroy.argtypes = [vp]
    lib.ahr_topk.argtypes = [vp, vp, c_int32, c_int32, c_int, vp, vp, vp]
    lib.ahr_topk_exact_begin.argtypes = [vp, vp, c_int32, c_int32, vp, vp, vp]
    lib.ahr_topk_exact_finish.argtypes = [vp]
    lib.ahr_topk_host.argtypes = [vp, vp, c_int32, c_int32, c_int, vp, vp, vp]
    lib.ahr_merge_topk.argtypes = [c_int, vp, vp, c_int32, c_int32, c_int32, vp, vp, vp]
    lib.ahr_merge_topk_packed.argtypes = [c_int, vp, c_int64, c_int32, c_int32, c_int32, vp, vp, vp]
    lib.ahr_graph_create.argtypes = [POINTER(vp), c_int, c_int64, POINTER(vp), POINTER(vp), vp, vp, vp, vp, vp]
    lib.ahr_graph_destroy.argtypes = [vp]
    lib.ahr_anchor_rerank.argtypes = [vp, c_int32, c_int32, vp, vp, POINTER(RerankParams), c_int32,
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ahr_anchor_rerank_rows.argtypes = [vp, c_int32, c_int32, vp, vp, c_int64, vp, POINTER(RerankParams), c_int32,
                                           vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ahr_expand.argtypes = [vp, c_int32, vp, vp, vp, c_int32, vp, vp, vp]
    lib.ahr_debug_tc_scores.argtypes = [vp, vp, c_int32, vp, vp]
    lib.ahr_rescore.argtypes = [vp, vp, c_int32, c_int32, vp, vp, vp]
    lib.ahr_debug_round4.argtypes = [c_int, vp, c_int64, vp, vp]
    lib.ahr_exact_stats.argtypes = [vp, POINTER(c_int64), POINTER(c_int64)]
    lib.ahr_set_profiling.argtypes = [vp, c_int]
    lib.ahr_get_profile.argtypes = [vp, POINTER(c_double), POINTER(c_int64), POINTER(c_int32)]
    lib.ahr_env_reset.argtypes = [vp, POINTER(EnvBatch), vp, vp, vp, vp, c_int32, c_int32, vp, vp, vp]
    lib.ahr_comm_create.argtypes = [POINTER(vp), c_int, c_int, c_int, c_int64]
    lib.ahr_comm_export.argtypes = [vp, vp]
    lib.ahr_comm_connect.argtypes = [vp, vp]
    lib.ahr_comm_destroy.argtypes = [vp]
    lib.ahr_comm_begin.argtypes = [vp, c_int32, c_int32, POINTER(vp), POINTER(vp)]
    lib.ahr_comm_exchange_merge.argtypes = [vp, c_int32, c_int32, vp, vp, vp]
    lib.ahr_topk_sharded.argtypes = [vp, vp, vp, c_int32, c_int32, c_int, vp, vp, vp]
    lib.ahr_graph_shard_create.argtypes = [POINTER(vp), c_int, c_int64, c_int64, c_int64, POINTER(vp), POINTER(vp), vp]
    lib.ahr_graph_shard_destroy.argtypes = [vp]
    lib.ahr_expand_shard_record_ints.argtypes = [c_int32, c_int32, c_int32]
    lib.ahr_expand_shard_stage1.argtypes = [vp, c_int32, vp, vp, vp, c_int32, c_int32, vp, vp]
    lib.ahr_expand_shard_stage2.argtypes = [vp, c_int32, c_int32, vp, vp, vp, vp, c_int32, c_int32, vp, vp, vp]
    lib.ahr_expand_shard_merge.argtypes = [c_int, c_int64, c_int32, c_int32, vp, vp, vp, vp, c_int32, c_int32, vp, vp, vp, vp, vp]
    lib.ahr_graph_shard_set_attrs.argtypes = [vp, vp, vp, vp, vp]
    lib.ahr_anchor_shard_record_ints.argtypes = [c_int32, c_int32, POINTER(RerankParams)]
    lib.ahr_anchor_shard_stage1.argtypes = [vp, c_int32, c_int32, vp, POINTER(RerankParams), vp, vp]
    lib.ahr_anchor_shard_stage2.argtypes = [vp, c_int32, c_int32, c_int32, vp, vp, POINTER(RerankParams), vp, vp, vp]
    lib.ahr_anchor_shard_merge.argtypes = [vp, c_int32, c_int32, c_int32, vp, vp, vp, POINTER(RerankParams), c_int32, vp, vp,
                                           vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ahr_lca_closure.argtypes = [vp, POINTER(c_int32), c_int32, vp, POINTER(c_int64), vp]
    lib.ahr_env_step.argtypes = [vp, POINTER(EnvBatch), vp, c_int32, c_double, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("ahr_last_error", "ahr_launch_count", "ahr_expand_shard_record_ints", "ahr_anchor_shard_record_ints"):
            fn.restype = c_int
    lib.ahr_expand_shard_record_ints.restype = c_int64
    lib.ahr_anchor_shard_record_ints.restype = c_int64


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if sources are newer and nvcc is present) the CUDA library.  Fails loudly."""
    global _lib
    if _lib is None:
        if _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # no nvcc on this machine: only fatal if there is no prebuilt library
                if not os.path.exists(_build.LIB_PATH):
                    raise ImportError(
                        "ah_rag_b200: libahrag_b200.so is missing and could not be built "
                        f"({e}); run `python -c 'import __graft_entry__ as g; g.build()'` on a machine with nvcc. "
                        "There is no CPU fallback.") from e
        lib = ctypes.CDLL(_build.LIB_PATH)
        lib.ahr_version.restype = c_int
        if lib.ahr_version() != ABI_VERSION:  # a stale library against newer ctypes declarations is an ABI mismatch
            raise ImportError(f"ah_rag_b200: {_build.LIB_PATH} is ABI version {lib.ahr_version()}, the Python side expects "
                              f"{ABI_VERSION}; rebuild with `python -c 'import __graft_entry__ as g; g.build()'`")
        _declare(lib)
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ahr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libahrag_b200 {what} failed (code {rc}): {msg}")
