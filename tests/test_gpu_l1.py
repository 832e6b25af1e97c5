"""GPU side of generate_l1_relations: the all-pairs centroid cosines come from the device
(canonical fp32 dot products, bit-equal to the oracle), the edges agree with the unmodified reference's up to the
few-ulp freedom of its fp32 ``np.dot`` (tolerance stated below)."""
import copy
import json
import os

import numpy as np
import pytest

from oracle import exact
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _cases():
    with open(os.path.join(gu.GOLDEN_DIR, "l1_relations.json"), "r", encoding="utf-8") as f:
        return json.load(f)


def test_device_cosines_are_the_canonical_scores():
    from ah_rag_b200.aggregate.l1_relations import _device_cosines, _normalised
    rng = np.random.default_rng(5)
    for t, d in ((7, 16), (130, 384), (300, 48)):
        mat = np.stack([_normalised(rng.normal(size=d).tolist()) for _ in range(t)]).astype(np.float32)
        got = _device_cosines(mat)
        want = exact.canonical_scores(mat, mat)  # float32(sum_j double(a_j) * double(b_j)), j ascending
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.abs(got.astype(np.float64) - mat.astype(np.float64) @ mat.astype(np.float64).T).max() <= 1e-6


def test_l1_relations_on_device_against_reference_golden():
    """Tolerance: the reference's cosine is an fp32 BLAS dot (accumulation order unspecified), ours the correctly
    rounded one: |difference| <= 2e-6 before rounding, so 4-decimal fields may differ by one unit (1e-4) and a pair may
    change sides only if its cosine is within 2e-4 of ``min_cosine``.  Overlap and Jaccard are exact."""
    from ah_rag_b200.aggregate.l1_relations import generate_l1_relations
    for c in _cases():
        kw = dict(c["kwargs"])
        top_k = kw.pop("top_k", None)
        edges = generate_l1_relations(copy.deepcopy(c["l1_nodes"]), artifact_dir=None, top_k=top_k, **kw)
        got = {(e["source"], e["target"]): e for e in edges}
        want = {(e["source"], e["target"]): e for e in c["edges"]}
        min_cos = kw.get("min_cosine", 0.5)
        for key in set(got) & set(want):
            g, w = got[key], want[key]
            assert g["overlap"] == w["overlap"] and g["jaccard"] == w["jaccard"] and g["confidence"] == w["confidence"]
            assert abs(g["cosine"] - w["cosine"]) <= 1.0001e-4 and abs(g["weight"] - w["weight"]) <= 1.0001e-4, key
        odd = set(got) ^ set(want)
        if top_k is None:
            for key in odd:
                e = got.get(key) or want.get(key)
                assert abs(e["cosine"] - min_cos) <= 2e-4, (key, e)
        else:
            assert len(edges) == len(c["edges"]) and len(odd) <= 2


def test_tensor_core_cosines_with_sparse_canonical_rescoring():
    """f3 on the tensor cores: the T x T contraction runs through K1 (fp32 rows as hi/lo bf16 halves); every pair that is
    compared with min_cosine at close range or reported carries the canonical score, all others are provably below it."""
    from ah_rag_b200.aggregate.l1_relations import _device_cosines, _normalised
    rng = np.random.default_rng(9)
    for t, d, min_cos in ((300, 48, 0.3), (700, 384, 0.15), (64, 16, 0.5)):
        base = rng.normal(size=(8, d))
        mat = np.stack([_normalised((base[i % 8] + 0.6 * rng.normal(size=d)).tolist()) for i in range(t)]).astype(np.float32)
        needed = rng.random((t, t)) < 0.02
        got = _device_cosines(mat, needed=needed, min_cosine=min_cos)
        want = exact.canonical_scores(mat, mat)
        up = np.triu(np.ones((t, t), dtype=bool), k=1)
        must = up & (needed | (want >= min_cos))
        assert must.sum() > 0
        assert np.array_equal(got[must].view(np.uint32), want[must].view(np.uint32))      # canonical where it matters
        assert np.array_equal((got >= min_cos) & up, (want >= min_cos) & up)              # same side of the threshold everywhere
        assert np.abs(got - want).max() <= 1e-4                                           # the rest: tensor-core accuracy
