"""CPU: the matcher oracle reproduces the committed cv2 4.13.0 golden vectors."""
import numpy as np
import pytest

from oracle import match_oracle as mo

CASES = ["rand", "ties", "dups", "tiny"]


@pytest.mark.parametrize("name", CASES)
def test_knn2_matches_cv2_golden(match_golden, name):
    g = match_golden
    idx, dist = mo.knn2(g[f"{name}_q"], g[f"{name}_t"])
    assert np.array_equal(idx, g[f"{name}_idx"])
    assert np.array_equal(dist, g[f"{name}_dist"])


@pytest.mark.parametrize("name", CASES)
def test_crosscheck_matches_cv2_golden(match_golden, name):
    g = match_golden
    qs, ti, dd = mo.crosscheck_match(g[f"{name}_q"], g[f"{name}_t"])
    ref = g[f"{name}_cross"]
    assert np.array_equal(qs, ref[:, 0]) and np.array_equal(ti, ref[:, 1])
    assert np.array_equal(dd.astype(np.int32), ref[:, 2])


def test_ratio_filter_is_f32_strict():
    idx = np.array([[3, 4], [5, 6], [7, 8]])
    dist = np.array([[3, 10], [30, 100], [0, 0]], dtype=np.int32)
    qi, ti, d = mo.ratio_filter(idx, dist, 0.3)
    # f32: 10*0.3f = 3.00000012 rounds (ties-to-even) to 3.0 -> 3 < 3 dropped;
    # 100*0.3f rounds up to 30.0000019 -> 30 kept; 0 < 0 dropped
    assert qi.tolist() == [1] and ti.tolist() == [5]


def test_sharded_merge_equals_unsharded(match_golden):
    g = match_golden
    for name in ["rand", "ties", "dups"]:
        q, t = g[f"{name}_q"], g[f"{name}_t"]
        cuts = [0, 17, t.shape[0] // 3, t.shape[0] // 3 + 1, t.shape[0]]
        parts = [mo.knn2(q, t[a:b], index_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
        idx, dist = mo.merge_top2(parts)
        assert np.array_equal(idx, g[f"{name}_idx"]) and np.array_equal(dist, g[f"{name}_dist"])


def test_knn_match_needs_two_train_rows(match_golden):
    g = match_golden
    with pytest.raises(IndexError):
        mo.knn_match(g["tiny_q"], g["tiny_t"][:1], 0.7)


def test_random_db_rows_layout():
    r = mo.random_db_rows(1000, 7, row_offset=5)
    assert r.shape == (1000, 61) and (r[:, 60] <= 63).all()
    again = mo.random_db_rows(10, 7, row_offset=5 + 990)
    assert np.array_equal(again, r[990:])
    assert abs(np.unpackbits(r[:, :60]).mean() - 0.5) < 0.01


def test_l2_oracle_matches_cv2_golden():
    """BFMatcher(NORM_L2).knnMatch k=2 on f32 descriptors: indices identical (incl. duplicate-row ties),
    distances within 1e-6 relative"""
    import os
    import synthdata
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_golden.npz"))
    for name in synthdata.L2_CASES:
        q, t = synthdata.l2_descriptors(name)
        assert np.allclose([q.astype(np.float64).sum(), t.astype(np.float64).sum()], G[f"{name}_checksum"], rtol=0, atol=1e-9)
        idx, dist = mo.knn2_l2(q, t)
        assert np.array_equal(idx, G[f"{name}_idx"])
        assert np.abs(dist - G[f"{name}_dist"]).max() <= 1e-6 * max(1.0, float(G[f"{name}_dist"].max()))
