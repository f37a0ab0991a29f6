"""CPU: the warp oracle reproduces cv2 4.13.0 `warpPerspective(INTER_LINEAR, BORDER_CONSTANT, 1)` (the
call at homographier/src/homographier/mod.rs:286-294) bit-exactly on the committed goldens; the DB
keyed-read oracle is sanity-checked against brute force."""
import os

import numpy as np
import pytest

from oracle import db_oracle as do
from oracle import warp_oracle as wo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "warp_golden.npz"))


@pytest.mark.parametrize("i", range(6))
def test_warp_gray_bit_exact(i):
    for j, (w, h) in enumerate(G["sizes"]):
        assert np.array_equal(wo.warp_perspective(G["gray"], G["H"][i], int(w), int(h), 1), G[f"g{i}_{j}"])


@pytest.mark.parametrize("i", [1, 2])
def test_warp_bgra_bit_exact(i):
    assert np.array_equal(wo.warp_perspective(G["bgra"], G["H"][i], 264, 200, 1), G[f"c{i}"])


def test_reference_test_warp_image_empty():
    """mod.rs:682-707: the identity homography leaves every pixel unchanged"""
    img = np.arange(4 * 4 * 4, dtype=np.uint8).reshape(4, 4, 4)
    assert np.array_equal(wo.warp_perspective(img, np.eye(3)), img)


def test_db_oracle_orders_by_response_then_row():
    rng = np.random.default_rng(0)
    n = 500
    x, y = rng.uniform(0, 100, n).astype(np.float32), rng.uniform(0, 100, n).astype(np.float32)
    r = rng.choice(np.float32([0.1, 0.2, 0.3]), n)
    im = rng.integers(1, 5, n).astype(np.int32)
    idx = do.select_rows(x, y, r, im, [0, 1, 0, 1], f_lod=1, box=(10.5, 10.5, 80.2, 90.0), limit=50)
    assert len(idx) == 50
    keys = [(-float(r[i]), int(i)) for i in idx]
    assert keys == sorted(keys)
    assert all(im[i] in (2, 4) and 10 <= x[i] <= 81 and 10 <= y[i] <= 90 for i in idx)


def test_lod_oracle_resamplers():
    """box mean of a constant-gradient band is exact; Lanczos keeps constants and is symmetric"""
    from oracle import lod_oracle as lo
    y, x = np.mgrid[0:64, 0:64].astype(np.float32)
    band = (2 * x + 3 * y).astype(np.float32)
    a = lo.resample_window(band, 0, 0, 16, 16, 4, "area")
    assert np.allclose(a, 2 * (4 * np.arange(16)[None, :] + 1.5) + 3 * (4 * np.arange(16)[:, None] + 1.5))
    c = lo.resample_window(np.full((64, 64), 7.0, np.float32), 0, 0, 16, 16, 4, "lanczos")
    assert np.allclose(c, 7.0, atol=1e-6)
    first, w = lo.taps(2, "lanczos")
    assert len(w) == 12 and np.allclose(w, w[::-1]) and first == -6
    assert np.array_equal(lo.resample_window(band, 3, 5, 8, 8, 1, "area"), band[5:13, 3:11])
