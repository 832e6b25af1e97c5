"""Relations between L1 topics -- drop-in for ``SemanticAggregator.generate_l1_relations``
(src/ah_rag/aggregate/semantic_aggregator.py:594-680).

The reference walks all T*(T-1)/2 topic pairs in Python: member-set overlap, Jaccard, one ``np.dot`` of the normalised
centroids per pair.  Here the three T x T matrices are built at once:

  overlap   sparse incidence product  M . M^T  (M: topics x members, scipy)                        exact integers
  jaccard   overlap / (|A| + |B| - overlap) in float64 -- the same IEEE division as len(inter) / len(union)  exact
  cosine    the T x T contraction of the L2-normalised fp32 centroids runs on the TENSOR CORES (K1, hi/lo bf16 split of the
            fp32 rows, ~2^-16 accurate); only the pairs that can matter -- tensor score within the error bound of
            ``min_cosine``, or kept by the overlap / Jaccard rule -- are re-scored canonically (``VectorIndex.rescore``,
            csrc/topk_exact.cu: float32(sum_j double(a_j) * double(b_j)) = the correctly rounded fp32 dot product), so every
            cosine that is compared with ``min_cosine`` or reported is the canonical one.
            ``np.dot`` on fp32 vectors (BLAS sdot) accumulates in an implementation-defined order, so the two differ by a
            few fp32 ulps (<= ~2e-6 here): the 4-decimal ``cosine`` / ``weight`` fields agree except on exact rounding
            boundaries, and a pair whose cosine sits within that distance of ``min_cosine`` may fall on the other side.
            ``cosine_fn`` lets a caller (and the parity tests) supply the matrix instead.

Rules, rounding, ordering, the optional top-k pruning (stable sort by the ROUNDED weight), the ``neighbors`` lists
written back into ``l1_nodes`` and the two JSON artifacts follow the reference line by line.  No CPU fallback for the
cosine matrix: without ``cosine_fn`` a CUDA device is required.
"""
from __future__ import annotations

import json
import os
from typing import Any, Callable, Dict, List, Optional

import numpy as np

ARTIFACT_DIR = "artifacts"


def _normalised(v: Optional[List[float]]) -> Optional[np.ndarray]:
    if v is None:
        return None
    arr = np.asarray(v, dtype=np.float32)
    n = np.linalg.norm(arr)
    return arr / n if n > 0 else arr


def _device_cosines(rows: np.ndarray, device: str = "cuda", needed: Optional[np.ndarray] = None,
                    min_cosine: Optional[float] = None) -> np.ndarray:
    """[T, T] dot products of the rows (each already L2-normalised), computed on the GPU.

    ``needed is None``: every entry is the canonical fp32 score (one thread per pair, the first implementation).
    Otherwise (``needed``: bool [T, T], ``min_cosine``): the contraction runs on the tensor cores and only the pairs with
    ``needed`` set or a tensor score >= min_cosine - eps are re-scored canonically; all other entries keep the tensor-core
    value, which is provably below ``min_cosine`` (eps = the certification bound of the exact top-k mode for fp32 rows)."""
    import torch
    from .. import engine
    dev = torch.device(device)
    x = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).to(dev)
    index = engine.VectorIndex(x)
    t, d = int(x.shape[0]), int(x.shape[1])
    ids = torch.arange(t, device=dev, dtype=torch.int64)
    if needed is None or min_cosine is None or d % 8 != 0:
        out = np.empty((t, t), dtype=np.float32)
        step = max(1, min(t, (1 << 22) // max(t, 1)))  # <= 4M (query, candidate) pairs per launch
        for q0 in range(0, t, step):
            q1 = min(t, q0 + step)
            cand = ids.unsqueeze(0).expand(q1 - q0, t).contiguous()
            out[q0:q1] = index.rescore(x[q0:q1].contiguous(), cand).cpu().numpy()
        return out
    tc = index.debug_tc_scores(x)  # [T, T] tensor-core accumulators (fp32 rows: hi/lo bf16 split, three operand segments)
    norms = x.double().norm(dim=1)
    eps = float(((3 * d / 2 + 64) * 2.0 ** -23 + 2.0 ** -16) * float(norms.max()) ** 2)  # api.cu: exact_eps_rel, fp32 rows
    cand = (tc >= (min_cosine - eps)) | torch.from_numpy(np.ascontiguousarray(needed)).to(dev)
    cand &= torch.triu(torch.ones((t, t), dtype=torch.bool, device=dev), diagonal=1)  # the rules read a < b only
    counts = cand.sum(dim=1)
    m = int(counts.max().item()) if t else 0
    out = tc.clone()
    if m > 0:
        # ragged candidate lists, padded with -1 (re-scored as -inf and never scattered)
        order = torch.argsort(cand.to(torch.int8), dim=1, descending=True, stable=True)[:, :m]
        valid = torch.arange(m, device=dev)[None, :] < counts[:, None]
        cand_ids = torch.where(valid, order, torch.full_like(order, -1)).contiguous()
        sc = index.rescore(x, cand_ids)
        r, c = torch.nonzero(valid, as_tuple=True)
        out[r, cand_ids[r, c]] = sc[r, c]
        out[cand_ids[r, c], r] = sc[r, c]
    return out.cpu().numpy()


def generate_l1_relations(l1_nodes: List[Dict[str, Any]], min_overlap: int = 3, min_jaccard: float = 0.2,
                          min_cosine: float = 0.5, top_k: Optional[int] = None,
                          artifact_dir: Optional[str] = ARTIFACT_DIR,
                          cosine_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                          device: str = "cuda") -> List[Dict[str, Any]]:
    """Undirected ``related_summary`` edges between L1 topics; ``l1_nodes`` gets its ``neighbors`` lists in place.
    ``artifact_dir=None`` skips the two JSON files the reference writes (``l1_edges.json``, ``l1_nodes.json``)."""
    if not l1_nodes:
        return []
    # later entries with the same topic_id replace earlier ones, as the reference's dict comprehensions do
    centroids = {int(n["topic_id"]): _normalised(n.get("centroid")) for n in l1_nodes}
    members = {int(n["topic_id"]): set(n.get("members", [])) for n in l1_nodes}
    confidence = {int(n["topic_id"]): float(n.get("confidence", 5.0)) for n in l1_nodes}
    topic_ids = sorted(centroids)
    t = len(topic_ids)

    # ---- overlap / union sizes for every pair: sparse incidence product -----------------------------------------
    from scipy import sparse
    vocab: Dict[Any, int] = {}
    rows_i: List[int] = []
    cols_i: List[int] = []
    for r, tid in enumerate(topic_ids):
        for m in members[tid]:
            rows_i.append(r)
            cols_i.append(vocab.setdefault(m, len(vocab)))
    inc = sparse.csr_matrix((np.ones(len(rows_i), dtype=np.int64), (rows_i, cols_i)), shape=(t, max(1, len(vocab))))
    overlap = np.asarray((inc @ inc.T).todense(), dtype=np.int64)
    size = np.asarray([len(members[tid]) for tid in topic_ids], dtype=np.int64)
    union = size[:, None] + size[None, :] - overlap
    with np.errstate(divide="ignore", invalid="ignore"):
        jaccard = np.where(union > 0, overlap / np.maximum(union, 1), 0.0)  # int / int -> float64, as in Python

    # ---- centroid cosines for every pair --------------------------------------------------------------------------
    has_c = np.asarray([centroids[tid] is not None for tid in topic_ids], dtype=bool)
    cosine = np.zeros((t, t), dtype=np.float64)
    if has_c.any():
        sel = np.flatnonzero(has_c)
        mat = np.stack([centroids[topic_ids[i]] for i in sel]).astype(np.float32)
        if cosine_fn is not None:
            c = np.asarray(cosine_fn(mat), dtype=np.float32)
        else:  # pairs kept by the set rules need their canonical cosine as well (it is reported in the edge)
            set_rule = ((overlap >= min_overlap) | (jaccard >= min_jaccard))[np.ix_(sel, sel)]
            c = np.asarray(_device_cosines(mat, device, needed=set_rule, min_cosine=float(min_cosine)), dtype=np.float32)
        cosine[np.ix_(sel, sel)] = c.astype(np.float64)  # float(np.float32) is exact

    # ---- rules on the upper triangle, in (a, b) order with a < b ------------------------------------------------
    both_empty = (size[:, None] == 0) & (size[None, :] == 0)
    keep = ((overlap >= min_overlap) | (jaccard >= min_jaccard) | (cosine >= min_cosine)) & ~both_empty
    ia, ib = np.nonzero(np.triu(keep, k=1))  # row-major: a ascending, then b ascending
    edges: List[Dict[str, Any]] = []
    for a, b in zip(ia.tolist(), ib.tolist()):
        jac, cos = float(jaccard[a, b]), float(cosine[a, b])
        ta, tb = topic_ids[a], topic_ids[b]
        edges.append({"source": int(ta), "target": int(tb), "relation_type": "related_summary",
                      "weight": round(0.5 * jac + 0.5 * cos, 4), "overlap": int(overlap[a, b]),
                      "jaccard": round(jac, 4), "cosine": round(cos, 4),
                      "confidence": round((confidence.get(ta, 5.0) + confidence.get(tb, 5.0)) / 2.0, 2)})
    if top_k is not None and len(edges) > top_k:
        edges = sorted(edges, key=lambda e: e["weight"], reverse=True)[:top_k]

    if artifact_dir is not None:
        os.makedirs(artifact_dir, exist_ok=True)
        with open(os.path.join(artifact_dir, "l1_edges.json"), "w", encoding="utf-8") as f:
            json.dump(edges, f, ensure_ascii=False, indent=2)
    adjacency: Dict[int, List[Any]] = {int(n["topic_id"]): [] for n in l1_nodes}
    for e in edges:
        adjacency[e["source"]].append((e["target"], e["weight"]))
        adjacency[e["target"]].append((e["source"], e["weight"]))
    for n in l1_nodes:
        near = sorted(adjacency.get(int(n["topic_id"]), []), key=lambda x: x[1], reverse=True)
        n["neighbors"] = [{"topic_id": tid, "weight": w} for tid, w in near]
    if artifact_dir is not None:
        with open(os.path.join(artifact_dir, "l1_nodes.json"), "w", encoding="utf-8") as f:
            json.dump(l1_nodes, f, ensure_ascii=False, indent=2)
    return edges
