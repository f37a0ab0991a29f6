"""CPU: the LoD tiler oracle's box-mean resampler against cv2.resize(INTER_AREA) (integer decimation factors: the same
box mean, cv2 accumulates in f32, the oracle and the CUDA kernel in f64 with a fixed tap order) and the tiling
bookkeeping of preprocessor/src/main.rs:212-301.  The Lanczos path stays unpinned (GDAL absent) and says so."""
import numpy as np
import pytest

from oracle import lod_oracle as lo

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_box_mean_matches_cv2_inter_area_within_one_ulp(scale):
    r = np.random.default_rng(scale)
    band = r.uniform(0, 1, (8 * 37, 8 * 53)).astype(np.float32)
    tw, th = band.shape[1] // scale, band.shape[0] // scale
    out = lo.resample_window(band, 0, 0, tw, th, scale, "area")
    ref = cv2.resize(band[:th * scale, :tw * scale], (tw, th), interpolation=cv2.INTER_AREA)
    # values in [0, 1): one f32 ulp is at most 2^-24 ~ 6e-8; cv2's f32 accumulation may be off by one more
    assert np.abs(out.astype(np.float64) - ref).max() <= 1.2e-7
    # an offset window reads the same pixels as the same window cut out first
    x0, y0 = 3 * scale, 5 * scale
    sub = lo.resample_window(band, x0, y0, tw - 8, th - 8, scale, "area")
    ref2 = cv2.resize(band[y0:y0 + (th - 8) * scale, x0:x0 + (tw - 8) * scale], (tw - 8, th - 8), interpolation=cv2.INTER_AREA)
    assert np.abs(sub.astype(np.float64) - ref2).max() <= 1.2e-7


def test_scale_one_is_a_copy():
    band = np.arange(48, dtype=np.float32).reshape(6, 8)
    assert np.array_equal(lo.resample_window(band, 2, 1, 4, 3, 1, "area"), band[1:4, 2:6])
    assert np.array_equal(lo.resample_window(band, 2, 1, 4, 3, 1, "lanczos"), band[1:4, 2:6])
