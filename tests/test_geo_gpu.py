"""GPU: radiometric pre-step, raster_to_mat and pixel -> ECEF object points through the C ABI vs the
oracle (bit-exact bytes; ECEF within 1e-6 m = 1.5e-13 relative) and the reference's own test constants."""
import ctypes as C

import numpy as np
import pytest

from oracle import geo_oracle as go

pytestmark = pytest.mark.gpu


def test_reference_pure_tests(dunk, ctx):
    ie = dunk.image_extractor
    assert ie.f32_to_u8(0.2, 0.1, 0.3, ctx) == 186                                   # mod.rs:546-555
    with pytest.raises(ie.PixelConversion):
        ie.f32_to_u8(float("nan"), 0.1, 0.3, ctx)                                    # mod.rs:557-566
    m = ie.band_merger([[0.0, 0.5, 1.0]] * 3, ie.BandsMinMax(-1.0, 2.0, -1.0, 2.0, -1.0, 2.0), ctx=ctx)
    assert len(m) == 3 and m[0, 0] == 155                                            # mod.rs:625-646
    assert abs(ie.gamma_correction(0.5) - 0.7297401) < 1e-7                           # mod.rs:517-526
    with pytest.raises(ie.PixelConversion):
        ie.gamma_correction(1.5)


def test_band_merger_bit_exact_full_tile(dunk, ctx):
    rng = np.random.default_rng(0)
    n = 1024 * 1024
    bands = [rng.uniform(0.0, 0.35, n).astype(np.float32) for _ in range(3)]
    bands[0][::1001] = np.nan
    bands[1][::1001] = np.nan
    bands[2][::2002] = np.nan                       # all-NaN pixels -> alpha 0
    bands[1][5::777] = 0.5                          # above max -> 0
    bands[2][7::555] = -0.1                         # below min -> 0
    mm = (0.0017, 0.31, 0.002, 0.3, 0.0, 0.33)
    got = dunk.image_extractor.band_merger(bands, dunk.image_extractor.BandsMinMax(*mm), ctx=ctx)
    exp = go.band_merger(*bands, mm)
    assert np.array_equal(got, exp)
    assert (got[:, 3] == 0).sum() == (exp[:, 3] == 0).sum() > 500
    got_bgra = dunk.image_extractor.band_merger(bands, dunk.image_extractor.BandsMinMax(*mm), bgra=True, ctx=ctx)
    assert np.array_equal(got_bgra, exp[:, [2, 1, 0, 3]])


def test_raster_to_mat_matches_host_mirror(dunk, ctx):
    from cubesat_apds_b200._lib import check, load, ptr
    rng = np.random.default_rng(1)
    w, h = 37, 21
    rgba = rng.integers(0, 256, (w * h, 4), dtype=np.uint8)
    out = np.empty((h, w, 4), np.uint8)
    check(load().dunk_raster_to_mat(ctx.handle, ptr(rgba), w, h, ptr(out)))
    assert np.array_equal(out, go.raster_to_mat(rgba, w, h))
    assert np.array_equal(out, dunk.homographier.raster_to_mat(rgba, w, h).mat)


def test_world_coordinates(dunk, ctx):
    fd = dunk.feature_database
    rng = np.random.default_rng(2)
    gt_d = [9.5, 9e-5, 0.0, 56.3, 0.0, -9e-5]                  # north-up dataset, ~10 m pixels
    gt_e = [9.4, 2.7e-4, 1e-6, 56.4, -2e-6, -2.7e-4]           # coarser, slightly rotated DEM
    heights = rng.uniform(0, 170, (1200, 1500))
    g = fd.Geotransform(gt_d, gt_e, heights, ctx)
    px, py = rng.uniform(0, 10980, 50000), rng.uniform(0, 10980, 50000)
    xyz, miss = g.world_coordinates(px, py)
    exp, ok = go.world_coordinates(px, py, gt_d, gt_e, heights, 1500, 1200)
    assert miss == int((~ok).sum())
    assert np.abs(xyz[ok] - exp[ok]).max() < 1e-6 and np.isnan(xyz[~ok]).all()
    # no elevation transform: height 0 (elevationdb.rs:76-79); the reference's coordinate_converter constants
    g0 = fd.Geotransform([9.68505, 1.0, 0.0, 56.105169, 0.0, 1.0], ctx=ctx)
    x, y, z = g0.get_world_coordinates(0.0, 0.0)
    assert abs(x - 3514316.2468943615) < 1e-6 and abs(y - 599769.3477405359) < 1e-6
    far = fd.Geotransform(gt_d, gt_e, heights[:4, :4], ctx)
    with pytest.raises(fd.NotFound):
        far.get_world_coordinates(10000.0, 10000.0)
    g.close(); g0.close(); far.close()


def test_gamma_threshold_table_is_exhaustively_exact(dunk, ctx):
    """the 255-threshold table behind f32_to_u8 equals the direct pow formula on every f32 in [0, 1]"""
    import ctypes as C
    from cubesat_apds_b200._lib import check, load
    bad = C.c_uint64(123)
    check(load().dunk_selftest_gamma_lut(ctx.handle, C.byref(bad)))
    assert bad.value == 0
