"""CPU retrieval baseline (TEST / BENCH INFRASTRUCTURE): best-effort restatement of the reference's
single-process CPU retrieval path for the similarity half.

The reference answers one query at a time through Chroma's HNSW index (hierarchical_graph.py:706-717),
which cannot be installed here; the closest honest CPU stand-in that returns the SAME answer as the engine's
contract (exact cosine, ties -> lower id up to fp32 GEMM rounding) is a blocked fp32 GEMM + top-k on all host
cores: ``torch`` ``Q @ X_chunk.T`` + ``torch.topk`` with a running merge.  bench.py times it on a bounded
sample of the workload and scales linearly in the number of corpus rows (stated in its ``sample`` field).
"""
from __future__ import annotations

import os
import time
from typing import Dict, Tuple

import torch


def torch_topk_pass(X: torch.Tensor, Q: torch.Tensor, k: int, chunk: int = 262144) -> Tuple[torch.Tensor, torch.Tensor]:
    """One retrieval pass on the CPU: X [n, d] fp32, Q [nq, d] fp32 -> (scores, ids) [nq, k]."""
    best_s = torch.full((Q.shape[0], k), float("-inf"))
    best_i = torch.full((Q.shape[0], k), -1, dtype=torch.int64)
    for r0 in range(0, X.shape[0], chunk):
        xs = X[r0:r0 + chunk]
        s = Q @ xs.T
        kk = min(k, xs.shape[0])
        cs, ci = torch.topk(s, kk, dim=1)
        cat_s = torch.cat([best_s, cs], dim=1)
        cat_i = torch.cat([best_i, ci + r0], dim=1)
        best_s, order = torch.topk(cat_s, k, dim=1)
        best_i = torch.gather(cat_i, 1, order)
    return best_s, best_i


def time_sample(n_rows_full: int, dim: int, k: int, sample_rows: int, sample_queries: int, repeats: int = 1,
                seed: int = 99) -> Dict[str, object]:
    """Time the CPU pass on ``sample_rows x dim`` fp32 rows and ``sample_queries`` queries; report queries/s
    extrapolated linearly in rows to the full corpus."""
    g = torch.Generator().manual_seed(seed)
    X = torch.randn((sample_rows, dim), generator=g)
    X /= X.norm(dim=1, keepdim=True)
    Q = torch.randn((sample_queries, dim), generator=g)
    Q /= Q.norm(dim=1, keepdim=True)
    torch_topk_pass(X[: min(sample_rows, 65536)], Q[:8], k)  # warm the thread pool / BLAS
    best = float("inf")
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        torch_topk_pass(X, Q, k)
        best = min(best, time.perf_counter() - t0)
    qps_sample = sample_queries / best
    return {
        "value": qps_sample * (sample_rows / float(n_rows_full)),
        "unit": "queries/s",
        "cores": int(torch.get_num_threads()),
        "host_cpus": int(os.cpu_count() or 0),
        "kind": "port",
        "seconds": best,
        "sample": (f"torch fp32 GEMM+topk on {sample_queries} queries x {sample_rows} rows x {dim} dims "
                   f"({best:.2f} s), scaled by rows to the {n_rows_full}-row corpus"),
    }
