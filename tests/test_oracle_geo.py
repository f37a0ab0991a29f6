"""CPU: the radiometric / geodesy oracle is pinned by the reference's own pure tests
(geotiff_extractor/src/image_extractor/mod.rs:517-555, 625-646; feature_database/src/elevationdb.rs:166-178)."""
import numpy as np

from oracle import geo_oracle as go


def test_gamma_correct_input():
    """mod.rs:517-526: gamma_correction(0.5) == 0.7297401"""
    assert go.gamma_correction(0.5) == np.float32(0.7297401)


def test_gamma_out_of_range():
    """mod.rs:528-544"""
    assert np.isnan(go.gamma_correction(1.5)) and np.isnan(go.gamma_correction(-0.5))


def test_convert_f32_to_u8():
    """mod.rs:546-555 (== 186) and the NaN case (-> unwrap_or(0) inside band_merger)"""
    assert go.f32_to_u8(0.2, 0.1, 0.3) == 186
    assert go.f32_to_u8(np.nan, 0.1, 0.3) == 0


def test_merging_bands():
    """mod.rs:625-646: merged_bands[0].r == 155"""
    m = go.band_merger([0.0, 0.5, 1.0], [0.0, 0.5, 1.0], [0.0, 0.5, 1.0], (-1.0, 2.0, -1.0, 2.0, -1.0, 2.0))
    assert m.shape == (3, 4) and m[0, 0] == 155 and (m[:, 3] == 255).all()
    n = go.band_merger([np.nan], [np.nan], [np.nan], (0, 1, 0, 1, 0, 1))
    assert n.tolist() == [[0, 0, 0, 0]]


def test_coordinate_converter():
    """elevationdb.rs:166-178: (lat 56.105169, lon 9.68505, h 0) -> ECEF x, y"""
    x, y, z = go.geodetic_to_ecef(56.105169, 9.68505, 0.0)
    assert abs(x - 3514316.2468943615) < 1e-6 and abs(y - 599769.3477405359) < 1e-6
    assert abs(np.sqrt(x * x + y * y + z * z) - 6.364e6) < 2e3


def test_geotransform_invert_round_trip():
    for gt in ([9.0, 1e-4, 0.0, 57.0, 0.0, -1e-4], [9.0, 1e-4, 2e-5, 57.0, -3e-5, -1e-4]):
        inv = go.geotransform_invert(gt)
        gx, gy = go.geotransform_apply(gt, 123.0, 456.0)
        px, py = go.geotransform_apply(inv, gx, gy)
        assert abs(px - 123.0) < 1e-6 and abs(py - 456.0) < 1e-6
    assert go.geotransform_invert([0, 1, 2, 0, 2, 4]) is None
