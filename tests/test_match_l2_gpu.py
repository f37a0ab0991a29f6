"""GPU parity: the tcgen05 float-descriptor matcher (dunk_knn2_l2) vs cv2 4.13.0
BFMatcher(NORM_L2).knnMatch k=2 goldens and the oracle: neighbour indices identical (ties -> lower
train index), distances within 1e-6 relative (exact f32 re-rank; OpenCV's SIMD sum associates
differently).  The tensor-core stage only nominates candidates, so exactness must not depend on TF32."""
import os

import numpy as np
import pytest

import synthdata
from oracle import match_oracle as mo

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_golden.npz"))


@pytest.mark.parametrize("name", list(synthdata.L2_CASES))
def test_knn2_l2_vs_cv2_golden(dunk, ctx, name):
    q, t = synthdata.l2_descriptors(name)
    idx, dist, stats = dunk.feature_extraction.knn2_l2(q, t, ctx)
    assert np.array_equal(idx, G[f"{name}_idx"])
    assert np.abs(dist - G[f"{name}_dist"]).max() <= 1e-6 * max(1.0, float(G[f"{name}_dist"].max()))


def test_many_slabs_and_ragged_sizes_vs_oracle(dunk, ctx):
    rng = np.random.default_rng(5)
    for nq, nt, dim in ((1, 2, 64), (130, 70001, 64), (257, 40013, 128)):
        t = rng.normal(size=(nt, dim)).astype(np.float32)
        q = (t[rng.integers(0, nt, nq)] * 1.01 + rng.normal(0, 0.3, (nq, dim))).astype(np.float32)
        idx, dist, stats = dunk.feature_extraction.knn2_l2(q, t, ctx)
        oi, od = mo.knn2_l2(q, t)
        assert np.array_equal(idx, oi), (nq, nt, dim, stats)
        assert np.abs(dist - od).max() <= 1e-6 * float(od.max())


def test_near_ties_force_the_exact_fallback_and_stay_exact(dunk, ctx):
    """train rows that differ from each other below TF32 resolution: stage A cannot rank them, stage B's
    proof fails for most queries, the exact fallback answers — results still equal the oracle"""
    rng = np.random.default_rng(6)
    base = rng.normal(size=(1, 64)).astype(np.float32)
    t = (base + rng.normal(0, 2e-4, (3000, 64))).astype(np.float32)
    q = (base + rng.normal(0, 2e-4, (64, 64))).astype(np.float32)
    idx, dist, stats = dunk.feature_extraction.knn2_l2(q, t, ctx)
    oi, od = mo.knn2_l2(q, t)
    assert stats[0] > 0
    assert np.array_equal(idx, oi)


def test_errors(dunk, ctx):
    fe = dunk.feature_extraction
    q = np.zeros((4, 64), np.float32)
    with pytest.raises(dunk.DunkError) as e:
        fe.knn2_l2(q, np.zeros((1, 64), np.float32), ctx)
    assert e.value.code == -211
    with pytest.raises(dunk.DunkError):
        fe.knn2_l2(np.zeros((4, 48), np.float32), np.zeros((9, 48), np.float32), ctx)


def test_ratio_filter_matches_oracle(dunk, ctx):
    q, t = synthdata.l2_descriptors("a")
    m = dunk.feature_extraction.get_knn_matches_l2(q, t, 2, 0.5, ctx)
    oi, od = mo.knn2_l2(q, t)
    keep = od[:, 0] < od[:, 1] * np.float32(0.5)
    assert np.array_equal(m["query_idx"], np.nonzero(keep)[0]) and np.array_equal(m["train_idx"], oi[keep, 0])
    assert 0 < len(m) < len(q)
