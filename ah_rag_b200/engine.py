"""Host-side handles over the C ABI: device memory, streams and lifetimes come from PyTorch; every
computation is a call into libahrag_b200.so with raw device pointers.  No CPU fallback anywhere.

  VectorIndex  -- device-resident embedding matrix + batched top-k          (replaces the Chroma collection,
                                                                               hierarchical_graph.py:605-611, :714)
  GraphCSR     -- device-resident typed CSR + node attributes               (replaces the NetworkX adjacency
                                                                               reads, hierarchical_graph.py:441-462)
  EnvBatchState -- device-resident state of B gym environments + reset/step  (replaces B sequential AHRAGGymEnv
                                                                               objects, gym_env.py:41-125)
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import EnvBatch, RerankParams, check

_NT_BITS = {"entity": 1 << 0, "summary": 1 << 1, "hyperedge": 1 << 2}


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the engine has no CPU path")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _stream_ptr(device: torch.device) -> int:
    """Raw handle of torch's current stream on ``device`` (the Stream-object route costs 2.3 us per call, this 0.3)."""
    return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())


class _on:
    """``with torch.cuda.device(dev)`` that costs nothing when ``dev`` already is the current device (every library entry
    point selects its handle's device itself; this only keeps torch's notion of the current device in step)."""
    __slots__ = ("dev", "ctx")

    def __init__(self, dev: torch.device) -> None:
        self.dev, self.ctx = dev, None

    def __enter__(self):
        if self.dev.index is not None and torch.cuda.current_device() != self.dev.index:
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def launch_count() -> int:
    """CUDA kernels launched by the library so far in this process (bench.py's ``gpu_launches``)."""
    return int(_lib.load().ahr_launch_count())


class VectorIndex:
    """Embedding index over rows already resident on one GPU.  ``rows``: [n, dim] float32 or bfloat16."""

    def __init__(self, rows: torch.Tensor, id_base: int = 0) -> None:
        self.lib = _lib.load()
        _require_cuda(rows, "rows")
        if rows.dim() != 2 or rows.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("rows must be a [n, dim] float32 or bfloat16 tensor")
        self.rows = rows  # keep alive: the library borrows the pointer
        self.device = rows.device
        self.n, self.dim = int(rows.shape[0]), int(rows.shape[1])
        self.dtype = rows.dtype
        self.id_base = int(id_base)
        self._h = ctypes.c_void_p()
        code = _lib.AHR_DTYPE_F32 if rows.dtype == torch.float32 else _lib.AHR_DTYPE_BF16
        with _on(self.device):
            check(self.lib.ahr_index_create(ctypes.byref(self._h), self.device.index or 0, self.n, self.dim, code,
                                            rows.data_ptr() if self.n else None, self.id_base), "ahr_index_create")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ahr_index_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _mode(self, mode: str) -> int:
        if mode == "bf16":
            return _lib.AHR_MODE_BF16
        if mode in ("exact", "fp32"):
            return _lib.AHR_MODE_EXACT
        raise ValueError(f"unknown mode {mode!r} (bf16 | exact)")

    def topk(self, q: torch.Tensor, k: int, mode: str = "bf16",
             out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """scores [nq, k] float32 and ids [nq, k] int64 on the device, sorted by (score desc, id asc).
        ``out``: preallocated contiguous (scores, ids) to write into (e.g. views of one exchange buffer)."""
        _require_cuda(q, "q")
        if q.dim() != 2 or q.shape[1] != self.dim or q.dtype != self.dtype:
            raise ValueError(f"q must be [nq, {self.dim}] {self.dtype}")
        nq = int(q.shape[0])
        if out is not None:
            scores, ids = out
            _require_cuda(scores, "out scores")
            _require_cuda(ids, "out ids")
            if scores.dtype != torch.float32 or ids.dtype != torch.int64 or scores.numel() != nq * k or ids.numel() != nq * k:
                raise ValueError("out must be (float32, int64) tensors of nq * k elements")
        else:
            scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_topk(self._h, q.data_ptr(), nq, int(k), self._mode(mode), scores.data_ptr(),
                                    ids.data_ptr(), _stream_ptr(self.device)), "ahr_topk")
        return scores, ids

    def topk_exact_begin(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """First half of an exact query: enqueues candidates + rescoring + certification and returns WITHOUT synchronising
        (``ahr_topk_exact_begin``).  The returned tensors are final, in stream order, once ``topk_exact_finish`` returned."""
        _require_cuda(q, "q")
        if q.dim() != 2 or q.shape[1] != self.dim or q.dtype != self.dtype:
            raise ValueError(f"q must be [nq, {self.dim}] {self.dtype}")
        nq = int(q.shape[0])
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        self._pending_q = q  # keep alive until finish
        with _on(self.device):
            check(self.lib.ahr_topk_exact_begin(self._h, q.data_ptr(), nq, int(k), scores.data_ptr(), ids.data_ptr(),
                                                _stream_ptr(self.device)), "ahr_topk_exact_begin")
        return scores, ids

    def topk_exact_finish(self) -> None:
        with _on(self.device):
            check(self.lib.ahr_topk_exact_finish(self._h), "ahr_topk_exact_finish")
        self._pending_q = None

    def topk_raw(self, q: torch.Tensor, k: int, mode: str, scores_ptr: int, ids_ptr: int) -> None:
        """Top-k written to raw device pointers ([nq, k] fp32 / int64), e.g. a slot of the peer exchange buffer."""
        _require_cuda(q, "q")
        if q.dim() != 2 or q.shape[1] != self.dim or q.dtype != self.dtype:
            raise ValueError(f"q must be [nq, {self.dim}] {self.dtype}")
        with _on(self.device):
            check(self.lib.ahr_topk(self._h, q.data_ptr(), int(q.shape[0]), int(k), self._mode(mode), scores_ptr, ids_ptr,
                                    _stream_ptr(self.device)), "ahr_topk")

    def topk_host(self, q_host: torch.Tensor, k: int, mode: str = "bf16",
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """End-to-end call with HOST buffers (H2D copy of the queries, top-k, D2H copy of the results)."""
        if q_host.is_cuda or not q_host.is_contiguous() or q_host.dtype != self.dtype:
            raise ValueError("q_host must be a contiguous host tensor of the index dtype")
        nq = int(q_host.shape[0])
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                   torch.empty((nq, k), dtype=torch.int64).pin_memory())
        with _on(self.device):
            check(self.lib.ahr_topk_host(self._h, q_host.data_ptr(), nq, int(k), self._mode(mode), out[0].data_ptr(),
                                         out[1].data_ptr(), _stream_ptr(self.device)), "ahr_topk_host")
        return out

    def set_profiling(self, enable: bool) -> None:
        check(self.lib.ahr_set_profiling(self._h, 1 if enable else 0), "ahr_set_profiling")

    def get_profile(self) -> Dict[str, Any]:
        """Summed CUDA-event duration of the K1 launches since the last call (device time on the launch stream)."""
        ms, n, s = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int32()
        check(self.lib.ahr_get_profile(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(s)), "ahr_get_profile")
        return {"k1_ms_total": float(ms.value), "k1_launches": int(n.value), "n_slabs": int(s.value)}

    def exact_stats(self) -> Tuple[int, int]:
        a, b = ctypes.c_int64(), ctypes.c_int64()
        check(self.lib.ahr_exact_stats(self._h, ctypes.byref(a), ctypes.byref(b)), "ahr_exact_stats")
        return int(a.value), int(b.value)

    def rescore(self, q: torch.Tensor, cand_ids: torch.Tensor) -> torch.Tensor:
        _require_cuda(q, "q")
        _require_cuda(cand_ids, "cand_ids")
        out = torch.empty(cand_ids.shape, dtype=torch.float32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_rescore(self._h, q.data_ptr(), int(q.shape[0]), int(cand_ids.shape[1]),
                                       cand_ids.data_ptr(), out.data_ptr(), _stream_ptr(self.device)), "ahr_rescore")
        return out

    def debug_tc_scores(self, q: torch.Tensor) -> torch.Tensor:
        """Raw tensor-core accumulators for every (query, row) pair -- tests only, small inputs."""
        _require_cuda(q, "q")
        out = torch.full((int(q.shape[0]), self.n), float("nan"), dtype=torch.float32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_debug_tc_scores(self._h, q.data_ptr(), int(q.shape[0]), out.data_ptr(),
                                               _stream_ptr(self.device)), "ahr_debug_tc_scores")
        return out


def merge_topk(scores: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device k-way merge of [parts, nq, k] candidate lists (the step after the NCCL all-gather)."""
    _require_cuda(scores, "scores")
    _require_cuda(ids, "ids")
    parts, nq, k = (int(x) for x in scores.shape)
    if scores.dtype != torch.float32 or ids.dtype != torch.int64 or tuple(ids.shape) != (parts, nq, k):
        raise ValueError("scores must be float32 [parts, nq, k] and ids int64 of the same shape")
    lib = _lib.load()
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    with _on(scores.device):
        check(lib.ahr_merge_topk(scores.device.index or 0, scores.data_ptr(), ids.data_ptr(), parts, nq, k,
                                 out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(scores.device)), "ahr_merge_topk")
    return out_s, out_i


def merge_topk_packed(packed: torch.Tensor, parts: int, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge straight out of the all-gathered exchange buffer: ``parts`` records of [nq*k fp32 | nq*k int64]."""
    _require_cuda(packed, "packed")
    if packed.dtype != torch.uint8 or packed.numel() % parts:
        raise ValueError("packed must be a uint8 buffer of parts equal records")
    lib = _lib.load()
    out_s = torch.empty((nq, k), dtype=torch.float32, device=packed.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=packed.device)
    with _on(packed.device):
        check(lib.ahr_merge_topk_packed(packed.device.index or 0, packed.data_ptr(), packed.numel() // parts, parts, nq, k,
                                        out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(packed.device)),
              "ahr_merge_topk_packed")
    return out_s, out_i


def round4_device(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "x")
    out = torch.empty_like(x)
    with _on(x.device):
        check(_lib.load().ahr_debug_round4(x.device.index or 0, x.data_ptr(), x.numel(), out.data_ptr(),
                                           _stream_ptr(x.device)), "ahr_debug_round4")
    return out


@dataclass
class SearchParams:
    """Resolved knobs of HierarchicalGraph.search (hierarchical_graph.py:681-699)."""
    alpha: float = 0.6
    beta: float = 0.2
    gamma: float = 0.1
    delta: float = 0.1
    judge_overall_min: Optional[float] = None
    confidence_min: Optional[float] = None
    type_filter: Optional[Sequence[str]] = None
    member_top_m: int = 5
    layer_boost: Optional[Dict[str, float]] = None
    parent_cap: int = 2

    def to_c(self) -> RerankParams:
        p = RerankParams()
        p.alpha, p.beta, p.gamma, p.delta = float(self.alpha), float(self.beta), float(self.gamma), float(self.delta)
        lb = self.layer_boost if self.layer_boost is not None else {"entity": 0.0, "summary": 1.0, "hyperedge": 0.0}
        # float(layer_boost_map.get(node_type, 0.0)): a node without node_type looks up the key None
        p.boost[0] = float(lb.get("entity", 0.0))
        p.boost[1] = float(lb.get("summary", 0.0))
        p.boost[2] = float(lb.get("hyperedge", 0.0))
        p.boost[3] = float(lb.get(None, 0.0))
        p.has_judge_min = 0 if self.judge_overall_min is None else 1
        p.judge_min = 0.0 if self.judge_overall_min is None else float(self.judge_overall_min)
        p.has_conf_min = 0 if self.confidence_min is None else 1
        p.conf_min = 0.0 if self.confidence_min is None else float(self.confidence_min)
        mask = 0
        if self.type_filter:  # `if type_filter and node_type not in type_filter` -- empty list == no filter
            for t in self.type_filter:
                mask |= _NT_BITS.get(t, 0)
            mask |= 1 << 30  # keep the mask non-zero even when no known type is listed (everything filtered)
        p.type_mask = mask
        p.member_top_m = max(0, int(self.member_top_m))
        p.parent_cap = int(self.parent_cap)
        return p


class GraphCSR:
    """Device-resident typed CSR + node attributes built from ``csr.csr_from_networkx`` / ``csr_from_edges``."""

    def __init__(self, arrays: Dict[str, Any], device: torch.device | str = "cuda") -> None:
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphCSR needs a CUDA device: the engine has no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_nodes = int(arrays["n_nodes"])
        dev = self.device
        self.row_ptr = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64)).to(dev) for a in arrays["row_ptr"]]
        self.col_idx = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev) for a in arrays["col_idx"]]
        self.node_type = torch.as_tensor(np.ascontiguousarray(arrays["node_type"], dtype=np.uint8)).to(dev)
        self.level = torch.as_tensor(np.ascontiguousarray(arrays["level"], dtype=np.uint8)).to(dev)
        self.judge = torch.as_tensor(np.ascontiguousarray(arrays["judge"], dtype=np.float64)).to(dev)
        self.judge_term = torch.as_tensor(np.ascontiguousarray(arrays["judge_term"], dtype=np.float64)).to(dev)
        self.conf = torch.as_tensor(np.ascontiguousarray(arrays["conf"], dtype=np.float64)).to(dev)
        rp = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.row_ptr])
        ci = (ctypes.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in self.col_idx])
        self._h = ctypes.c_void_p()
        with _on(dev):
            check(self.lib.ahr_graph_create(ctypes.byref(self._h), dev.index, self.n_nodes, rp, ci,
                                            self.node_type.data_ptr(), self.level.data_ptr(), self.judge.data_ptr(),
                                            self.judge_term.data_ptr(), self.conf.data_ptr()), "ahr_graph_create")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ahr_graph_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def nbytes(self) -> int:
        ts = self.row_ptr + self.col_idx + [self.node_type, self.level, self.judge, self.judge_term, self.conf]
        return sum(t.numel() * t.element_size() for t in ts)

    def anchor_rerank(self, seed_nodes: torch.Tensor, seed_scores: torch.Tensor, params: SearchParams, top_k: int,
                      want_cluster: bool = True, row_to_node: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Batched hierarchical_graph.py:719-789.  seed_scores fp32 [nq, n_seeds]; seed_nodes either int32 dense node ids
        (-1 = none) or, with ``row_to_node`` (int32 [n_rows]), the int64 index rows exactly as ``VectorIndex.topk`` returns
        them -- the row -> node lookup then happens inside the kernel (``ahr_anchor_rerank_rows``)."""
        _require_cuda(seed_nodes, "seed_nodes")
        _require_cuda(seed_scores, "seed_scores")
        want = torch.int64 if row_to_node is not None else torch.int32
        if seed_nodes.dtype != want or seed_scores.dtype != torch.float32 or seed_nodes.shape != seed_scores.shape:
            raise ValueError(f"seed_nodes must be {want} and seed_scores float32 of the same [nq, n_seeds] shape")
        if row_to_node is not None:
            _require_cuda(row_to_node, "row_to_node")
            if row_to_node.dtype != torch.int32:
                raise ValueError("row_to_node must be int32")
        nq, ns = int(seed_nodes.shape[0]), int(seed_nodes.shape[1])
        dev = self.device
        cp = params.to_c()
        cap = ns * (1 + max(cp.parent_cap, cp.member_top_m))
        out = {
            "nodes": torch.empty((nq, top_k), dtype=torch.int32, device=dev),
            "semantic": torch.empty((nq, top_k), dtype=torch.float64, device=dev),
            "score": torch.empty((nq, top_k), dtype=torch.float64, device=dev),
            "count": torch.empty((nq,), dtype=torch.int32, device=dev),
        }
        null = None
        if want_cluster:
            out["exp_nodes"] = torch.empty((nq, cap), dtype=torch.int32, device=dev)
            out["exp_semantic"] = torch.empty((nq, cap), dtype=torch.float64, device=dev)
            out["exp_kind"] = torch.empty((nq, cap), dtype=torch.uint8, device=dev)
            out["exp_count"] = torch.empty((nq,), dtype=torch.int32, device=dev)
            out["seed_semantic"] = torch.empty((nq, ns), dtype=torch.float64, device=dev)
        if nq == 0:
            return out
        tail = (ctypes.byref(cp), int(top_k),
                out["nodes"].data_ptr(), out["semantic"].data_ptr(), out["score"].data_ptr(), out["count"].data_ptr(),
                out["exp_nodes"].data_ptr() if want_cluster else null,
                out["exp_semantic"].data_ptr() if want_cluster else null,
                out["exp_kind"].data_ptr() if want_cluster else null,
                out["exp_count"].data_ptr() if want_cluster else null,
                out["seed_semantic"].data_ptr() if want_cluster else null,
                _stream_ptr(dev))
        with _on(dev):
            if row_to_node is not None:
                check(self.lib.ahr_anchor_rerank_rows(self._h, nq, ns, seed_nodes.data_ptr(), row_to_node.data_ptr(),
                                                      int(row_to_node.numel()), seed_scores.data_ptr(), *tail),
                      "ahr_anchor_rerank_rows")
            else:
                check(self.lib.ahr_anchor_rerank(self._h, nq, ns, seed_nodes.data_ptr(), seed_scores.data_ptr(), *tail),
                      "ahr_anchor_rerank")
        return out

    def lca_closure(self, node_ids: Sequence[int]) -> torch.Tensor:
        """environment.py:243-299, closure step: dense ids (int64, ascending) of the nodes that lie below every input
        over belongs_to and have no belongs_to parent with the same property (``ahr_lca_closure``)."""
        ids = (ctypes.c_int32 * max(1, len(node_ids)))(*[int(v) for v in node_ids])
        flags = torch.empty((self.n_nodes,), dtype=torch.uint8, device=self.device)
        with _on(self.device):
            # no count requested: the library call is asynchronous, torch.nonzero below is the one synchronisation
            check(self.lib.ahr_lca_closure(self._h, ids, len(node_ids), flags.data_ptr(), None, _stream_ptr(self.device)),
                  "ahr_lca_closure")
        return torch.nonzero(flags, as_tuple=False).view(-1)

    def expand(self, ops: torch.Tensor, in_ids: torch.Tensor, in_offsets: torch.Tensor, limit: int = 10
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Batched environment.py:393-480.  ops int32 [n_env]; in_ids int32 (CSR by in_offsets int32 [n_env+1])."""
        for t, name in ((ops, "ops"), (in_ids, "in_ids"), (in_offsets, "in_offsets")):
            _require_cuda(t, name)
            if t.dtype != torch.int32:
                raise ValueError(f"{name} must be int32")
        n_env = int(ops.shape[0])
        out_ids = torch.empty((n_env, limit), dtype=torch.int32, device=self.device)
        out_cnt = torch.empty((n_env,), dtype=torch.int32, device=self.device)
        if n_env == 0:
            return out_ids, out_cnt
        with _on(self.device):
            check(self.lib.ahr_expand(self._h, n_env, ops.data_ptr(), in_ids.data_ptr() if in_ids.numel() else None,
                                      in_offsets.data_ptr(), int(limit), out_ids.data_ptr(), out_cnt.data_ptr(),
                                      _stream_ptr(self.device)), "ahr_expand")
        return out_ids, out_cnt


_NP_OF = {torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8, torch.float64: np.float64}


class GraphShard:
    """One rank's node range [lo, hi) of the typed CSR (``ahr_graph_shard_*``, csrc/graph_shard.cu): the adjacency rows of
    the owned nodes (all six slots, neighbours as GLOBAL dense ids) and their node types.  ``arrays`` is the full-graph
    dict of ``csr.csr_from_networkx`` (numpy, or torch tensors on the host or the device); only the owned slice is kept."""

    def __init__(self, arrays: Dict[str, Any], lo: int, hi: int, device: torch.device | str = "cuda") -> None:
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphShard needs a CUDA device: the engine has no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self.n_nodes, self.lo, self.hi = int(arrays["n_nodes"]), int(lo), int(hi)
        self.row_ptr, self.col_idx = [], []

        def own(x, lo_, hi_, dt):
            """the slice [lo_, hi_) of a numpy array or of a torch tensor (host or device) as its own device tensor"""
            if isinstance(x, torch.Tensor):
                return x[lo_:hi_].to(device=dev, dtype=dt).clone().contiguous()
            return torch.as_tensor(np.ascontiguousarray(x[lo_:hi_], dtype=_NP_OF[dt])).to(dev)

        for s_ in range(_lib.AHR_N_CSR):
            rp = arrays["row_ptr"][s_]
            a, b = int(rp[self.lo]), int(rp[self.hi])
            self.row_ptr.append(own(rp, self.lo, self.hi + 1, torch.int64) - a)
            self.col_idx.append(own(arrays["col_idx"][s_], a, b, torch.int32))
        self.node_type = own(arrays["node_type"], self.lo, self.hi, torch.uint8)
        self.level, self.judge = own(arrays["level"], self.lo, self.hi, torch.uint8), own(arrays["judge"], self.lo, self.hi, torch.float64)
        self.judge_term = own(arrays["judge_term"], self.lo, self.hi, torch.float64)
        self.conf = own(arrays["conf"], self.lo, self.hi, torch.float64)
        rp = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.row_ptr])
        ci = (ctypes.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in self.col_idx])
        self._h = ctypes.c_void_p()
        with _on(dev):
            check(self.lib.ahr_graph_shard_create(ctypes.byref(self._h), dev.index, self.n_nodes, self.lo, self.hi, rp, ci,
                                                  self.node_type.data_ptr() if self.hi > self.lo else None),
                  "ahr_graph_shard_create")
            if self.hi > self.lo:
                check(self.lib.ahr_graph_shard_set_attrs(self._h, self.level.data_ptr(), self.judge.data_ptr(),
                                                         self.judge_term.data_ptr(), self.conf.data_ptr()), "ahr_graph_shard_set_attrs")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ahr_graph_shard_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def nbytes(self) -> int:
        ts = self.row_ptr + self.col_idx + [self.node_type, self.level, self.judge, self.judge_term, self.conf]
        return sum(t.numel() * t.element_size() for t in ts)

    # ---- anchor expansion + rerank over the sharded graph (ahr_anchor_shard_*) ----
    def anchor_record_ints(self, stage: int, n_seeds: int, cp) -> int:
        n = int(self.lib.ahr_anchor_shard_record_ints(stage, int(n_seeds), ctypes.byref(cp)))
        if n < 0:
            raise ValueError("sharded anchor expansion: bad n_seeds / params")
        return n

    def anchor_stage1(self, seed_nodes: torch.Tensor, params: "SearchParams") -> torch.Tensor:
        nq, ns = int(seed_nodes.shape[0]), int(seed_nodes.shape[1])
        cp = params.to_c()
        out = torch.empty((nq, self.anchor_record_ints(1, ns, cp)), dtype=torch.int32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_anchor_shard_stage1(self._h, nq, ns, seed_nodes.data_ptr(), ctypes.byref(cp), out.data_ptr(),
                                                   _stream_ptr(self.device)), "ahr_anchor_shard_stage1")
        return out

    def anchor_stage2(self, world: int, bounds, seed_nodes: torch.Tensor, params: "SearchParams", gathered1) -> torch.Tensor:
        nq, ns = int(seed_nodes.shape[0]), int(seed_nodes.shape[1])
        cp = params.to_c()
        out = torch.empty((nq, max(1, self.anchor_record_ints(2, ns, cp))), dtype=torch.int32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_anchor_shard_stage2(self._h, nq, ns, int(world), bounds.data_ptr(), seed_nodes.data_ptr(),
                                                   ctypes.byref(cp), gathered1.data_ptr(), out.data_ptr(),
                                                   _stream_ptr(self.device)), "ahr_anchor_shard_stage2")
        return out

    def anchor_merge(self, world: int, bounds, seed_nodes: torch.Tensor, seed_scores: torch.Tensor, params: "SearchParams",
                     top_k: int, gathered1, gathered2, want_cluster: bool = True) -> Dict[str, torch.Tensor]:
        nq, ns = int(seed_nodes.shape[0]), int(seed_nodes.shape[1])
        dev = self.device
        cp = params.to_c()
        cap = ns * (1 + max(cp.parent_cap, cp.member_top_m))
        out = {"nodes": torch.empty((nq, top_k), dtype=torch.int32, device=dev),
               "semantic": torch.empty((nq, top_k), dtype=torch.float64, device=dev),
               "score": torch.empty((nq, top_k), dtype=torch.float64, device=dev),
               "count": torch.empty((nq,), dtype=torch.int32, device=dev)}
        if want_cluster:
            out["exp_nodes"] = torch.empty((nq, cap), dtype=torch.int32, device=dev)
            out["exp_semantic"] = torch.empty((nq, cap), dtype=torch.float64, device=dev)
            out["exp_kind"] = torch.empty((nq, cap), dtype=torch.uint8, device=dev)
            out["exp_count"] = torch.empty((nq,), dtype=torch.int32, device=dev)
            out["seed_semantic"] = torch.empty((nq, ns), dtype=torch.float64, device=dev)
        ptr = lambda k_: out[k_].data_ptr() if want_cluster else None
        with _on(dev):
            check(self.lib.ahr_anchor_shard_merge(self._h, nq, ns, int(world), bounds.data_ptr(), seed_nodes.data_ptr(),
                                                  seed_scores.data_ptr(), ctypes.byref(cp), int(top_k), gathered1.data_ptr(),
                                                  gathered2.data_ptr(), out["nodes"].data_ptr(), out["semantic"].data_ptr(),
                                                  out["score"].data_ptr(), out["count"].data_ptr(), ptr("exp_nodes"),
                                                  ptr("exp_semantic"), ptr("exp_kind"), ptr("exp_count"), ptr("seed_semantic"),
                                                  _stream_ptr(dev)), "ahr_anchor_shard_merge")
        return out

    def record_ints(self, stage: int, limit: int, max_ids: int) -> int:
        n = int(self.lib.ahr_expand_shard_record_ints(stage, int(limit), int(max_ids)))
        if n < 0:
            raise ValueError(f"sharded expansion: limit must be in [1, 32] and max_ids >= 1 (got {limit}, {max_ids})")
        return n

    def stage1(self, ops, in_ids, in_offsets, max_ids: int, limit: int) -> torch.Tensor:
        n_env = int(ops.shape[0])
        out = torch.empty((n_env, self.record_ints(1, limit, max_ids)), dtype=torch.int32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_expand_shard_stage1(self._h, n_env, ops.data_ptr(), in_ids.data_ptr(), in_offsets.data_ptr(),
                                                   int(max_ids), int(limit), out.data_ptr(), _stream_ptr(self.device)),
                  "ahr_expand_shard_stage1")
        return out

    def stage2(self, world: int, bounds, ops, in_ids, in_offsets, max_ids: int, limit: int, gathered1) -> torch.Tensor:
        n_env = int(ops.shape[0])
        out = torch.empty((n_env, self.record_ints(2, limit, max_ids)), dtype=torch.int32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_expand_shard_stage2(self._h, n_env, int(world), bounds.data_ptr(), ops.data_ptr(), in_ids.data_ptr(),
                                                   in_offsets.data_ptr(), int(max_ids), int(limit), gathered1.data_ptr(),
                                                   out.data_ptr(), _stream_ptr(self.device)), "ahr_expand_shard_stage2")
        return out

    def merge(self, world: int, bounds, ops, in_ids, in_offsets, max_ids: int, limit: int, gathered1, gathered2):
        n_env = int(ops.shape[0])
        out_ids = torch.empty((n_env, limit), dtype=torch.int32, device=self.device)
        out_cnt = torch.empty((n_env,), dtype=torch.int32, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_expand_shard_merge(self.device.index, self.n_nodes, n_env, int(world), bounds.data_ptr(),
                                                  ops.data_ptr(), in_ids.data_ptr(), in_offsets.data_ptr(), int(max_ids), int(limit),
                                                  gathered1.data_ptr(), gathered2.data_ptr(), out_ids.data_ptr(), out_cnt.data_ptr(),
                                                  _stream_ptr(self.device)), "ahr_expand_shard_merge")
        return out_ids, out_cnt


class EnvBatchState:
    """Device-resident state of ``n_envs`` gym environments over one ``GraphCSR`` (``ahr_env_reset`` /
    ``ahr_env_step``): selection and frontier sets, the current observation, step counters.  Outputs are
    preallocated CUDA tensors overwritten by every call: ``features`` [B, 84] fp32, ``mask`` [B, 6] u8,
    ``rewards`` [B] fp64, ``done`` [B] u8.  Nothing here synchronises with the host."""

    FEATS = 84
    N_ACTIONS = 6

    def __init__(self, csr: GraphCSR, n_envs: int, max_steps: int = 6, repeat_penalty: float = 0.02,
                 cap_selection: Optional[int] = None, cap_frontier: Optional[int] = None) -> None:
        self.csr = csr
        self.n_envs = int(n_envs)
        self.max_steps = max(1, int(max_steps))
        self.repeat_penalty = float(repeat_penalty)
        dev = csr.device
        B = self.n_envs
        # an episode adds at most 3 selected and 10 frontier nodes per step (gym_env.py:94-100)
        self.cap_sel = int(cap_selection or max(3 * (self.max_steps + 2), 32))
        self.cap_fr = int(cap_frontier or (10 + 10 * (self.max_steps + 2)))
        i32 = dict(dtype=torch.int32, device=dev)
        self.state = {
            "sel_ids": torch.zeros((B, self.cap_sel), **i32), "sel_cnt": torch.zeros((B,), **i32),
            "fr_ids": torch.zeros((B, self.cap_fr), **i32), "fr_cnt": torch.zeros((B,), **i32),
            "shown_ids": torch.full((B, 10), -1, **i32),
            "shown_score": torch.zeros((B, 10), dtype=torch.float64, device=dev),
            "shown_sem": torch.zeros((B, 10), dtype=torch.float64, device=dev),
            "shown_cnt": torch.zeros((B,), **i32), "n_seeds": torch.zeros((B,), **i32),
            "obs_step": torch.zeros((B,), **i32), "obs_sel": torch.zeros((B,), **i32),
            "obs_fr": torch.zeros((B,), **i32), "env_steps": torch.zeros((B,), **i32),
            "gym_steps": torch.zeros((B,), **i32), "last_action": torch.full((B,), -1, **i32),
            "done": torch.zeros((B,), dtype=torch.uint8, device=dev), "overflow": torch.zeros((1,), **i32),
        }
        self.features = torch.zeros((B, self.FEATS), dtype=torch.float32, device=dev)
        self.mask = torch.zeros((B, self.N_ACTIONS), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((B,), dtype=torch.float64, device=dev)
        self.done = torch.zeros((B,), dtype=torch.uint8, device=dev)
        eb = EnvBatch()
        eb.n_env, eb.cap_sel, eb.cap_fr, eb.reserved = B, self.cap_sel, self.cap_fr, 0
        for name, t in self.state.items():
            setattr(eb, name, t.data_ptr())
        self._eb = eb
        self._ready = False

    def reset(self, rerank: Dict[str, torch.Tensor], n_seeds: int, features: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None):
        """``rerank`` = output of ``GraphCSR.anchor_rerank`` for one query per environment (top_k <= 10).
        ``features`` / ``mask``: optional contiguous destinations ([B, 84] fp32 / [B, 6] u8) instead of the handle's own."""
        nodes, sem, score, count = rerank["nodes"], rerank["semantic"], rerank["score"], rerank["count"]
        if int(nodes.shape[0]) != self.n_envs:
            raise ValueError(f"expected {self.n_envs} reranked lists, got {int(nodes.shape[0])}")
        for t, name in ((nodes, "nodes"), (sem, "semantic"), (score, "score"), (count, "count")):
            _require_cuda(t, name)
        features = self.features if features is None else features
        mask = self.mask if mask is None else mask
        dev = self.csr.device
        if self.n_envs:
            with _on(dev):
                check(self.csr.lib.ahr_env_reset(self.csr._h, ctypes.byref(self._eb), nodes.data_ptr(), sem.data_ptr(),
                                                 score.data_ptr(), count.data_ptr(), int(nodes.shape[1]), int(n_seeds),
                                                 features.data_ptr(), mask.data_ptr(), _stream_ptr(dev)),
                      "ahr_env_reset")
        self._ready = True
        return features, mask

    def capture_rollout(self, rerank: Dict[str, torch.Tensor], n_seeds: int, actions: torch.Tensor):
        """Reset + ``T`` steps as ONE CUDA graph (``actions``: int32 [T, B] on the device, read at replay time, so a policy
        can overwrite it between replays).  The seven launches of a 6-step episode are 8-15 us of device time each but
        ~90 us of Python / driver launch cost when issued one by one; a replay is a single submission.
        Returns ``(replay, out)``: ``replay()`` enqueues the episode on the current stream, ``out`` holds the per-step
        results it overwrites -- features [T + 1, B, 84] (index 0 = after reset), rewards [T, B], done [T, B], mask
        [T + 1, B, 6]."""
        _require_cuda(actions, "actions")
        if actions.dtype != torch.int32 or actions.dim() != 2 or int(actions.shape[1]) != self.n_envs:
            raise ValueError(f"actions must be int32 [T, {self.n_envs}]")
        T, B, dev = int(actions.shape[0]), self.n_envs, self.csr.device
        out = {"features": torch.zeros((T + 1, B, self.FEATS), dtype=torch.float32, device=dev),
               "mask": torch.zeros((T + 1, B, self.N_ACTIONS), dtype=torch.uint8, device=dev),
               "rewards": torch.zeros((T, B), dtype=torch.float64, device=dev),
               "done": torch.zeros((T, B), dtype=torch.uint8, device=dev)}

        def episode():  # every launch writes straight into its slice of `out`: 1 + T kernels, no copies
            self.reset(rerank, n_seeds, out["features"][0], out["mask"][0])
            for t in range(T):
                self.step(actions[t], out["features"][t + 1], out["rewards"][t], out["done"][t], out["mask"][t + 1])

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            episode()  # warm-up outside the capture (first-use attribute settings, lazy module loads)
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                episode()
        torch.cuda.current_stream(dev).wait_stream(side)
        return graph.replay, out

    def step(self, actions: torch.Tensor, features: Optional[torch.Tensor] = None, rewards: Optional[torch.Tensor] = None,
             done: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        """``actions`` int32 [B] on the device, values 0..5 (gym_env.py:14-20).  Optional contiguous destinations
        replace the handle's own output tensors (e.g. slices of a [T, B, ...] trajectory buffer)."""
        if not self._ready:
            raise RuntimeError("Call reset() first")
        _require_cuda(actions, "actions")
        if actions.dtype != torch.int32 or tuple(actions.shape) != (self.n_envs,):
            raise ValueError(f"actions must be int32 of shape ({self.n_envs},)")
        features = self.features if features is None else features
        rewards = self.rewards if rewards is None else rewards
        done = self.done if done is None else done
        mask = self.mask if mask is None else mask
        dev = self.csr.device
        if self.n_envs:
            with _on(dev):
                check(self.csr.lib.ahr_env_step(self.csr._h, ctypes.byref(self._eb), actions.data_ptr(), self.max_steps,
                                                self.repeat_penalty, features.data_ptr(), rewards.data_ptr(),
                                                done.data_ptr(), mask.data_ptr(), _stream_ptr(dev)),
                      "ahr_env_step")
        return features, rewards, done, mask

    def check_overflow(self) -> None:
        """Raises if a selection / frontier list outgrew its capacity (synchronises)."""
        if int(self.state["overflow"].item()):
            raise RuntimeError("EnvBatchState: selection/frontier capacity exceeded; raise cap_selection/cap_frontier")


class PeerExchange:
    """Handle over ``ahr_comm``: the cross-GPU candidate exchange through NVLink peer memory (csrc/exchange.cu).
    Rendezvous (once): every rank creates its buffer and ``export_handle()``s 64 bytes, the ranks swap them over any
    transport, every rank ``connect(handles)``s.  ``dist.ShardedIndex.enable_peer_exchange`` does this over the process
    group with collective-safe error handling."""

    def __init__(self, device: torch.device, rank: int, world: int, max_record_bytes: int, swap_handles=None) -> None:
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world)
        self.max_record_bytes = int(max_record_bytes)
        self._h = ctypes.c_void_p()
        with _on(self.device):
            check(self.lib.ahr_comm_create(ctypes.byref(self._h), self.device.index or 0, self.rank, self.world,
                                           self.max_record_bytes), "ahr_comm_create")
        if swap_handles is not None and self.world > 1:
            self.connect(swap_handles(self.export_handle()))

    def export_handle(self) -> bytes:
        mine = ctypes.create_string_buffer(_lib.AHR_IPC_HANDLE_BYTES)
        with _on(self.device):
            check(self.lib.ahr_comm_export(self._h, mine), "ahr_comm_export")
        return bytes(mine.raw)

    def connect(self, handles: Sequence[bytes]) -> None:
        if len(handles) != self.world or any(h is None or len(h) != _lib.AHR_IPC_HANDLE_BYTES for h in handles):
            raise RuntimeError("connect needs one 64-byte handle per rank, in rank order")
        blob = ctypes.create_string_buffer(b"".join(handles), self.world * _lib.AHR_IPC_HANDLE_BYTES)
        with _on(self.device):
            check(self.lib.ahr_comm_connect(self._h, blob), "ahr_comm_connect")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ahr_comm_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def begin(self, nq: int, k: int) -> Tuple[int, int]:
        """Start an exchange: raw device pointers (scores, ids) of this rank's slot."""
        ps, pi = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.lib.ahr_comm_begin(self._h, int(nq), int(k), ctypes.byref(ps), ctypes.byref(pi)), "ahr_comm_begin")
        return int(ps.value), int(pi.value)

    def topk_sharded(self, index: "VectorIndex", q: torch.Tensor, k: int, mode: str = "bf16") -> Tuple[torch.Tensor, torch.Tensor]:
        """Local top-k over ``index`` (this rank's node range), exchange with the peers, merge -- one library call
        (``ahr_topk_sharded``: the slab reduction of the local top-k pushes to the peers itself)."""
        _require_cuda(q, "q")
        if q.dim() != 2 or q.shape[1] != index.dim or q.dtype != index.dtype:
            raise ValueError(f"q must be [nq, {index.dim}] {index.dtype}")
        nq = int(q.shape[0])
        out_s = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_topk_sharded(index._h, self._h, q.data_ptr(), nq, int(k), index._mode(mode), out_s.data_ptr(),
                                            out_i.data_ptr(), _stream_ptr(self.device)), "ahr_topk_sharded")
        return out_s, out_i

    def exchange_merge(self, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Push this rank's slot to the peers, wait for theirs, merge: [nq, k] results, identical on every rank."""
        out_s = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        with _on(self.device):
            check(self.lib.ahr_comm_exchange_merge(self._h, int(nq), int(k), out_s.data_ptr(), out_i.data_ptr(),
                                                   _stream_ptr(self.device)), "ahr_comm_exchange_merge")
        return out_s, out_i
