"""Oracle for the steps either side of the hot path — TEST INFRASTRUCTURE ONLY.

* radiometric pre-step `band_merger` / `f32_to_u8` / `gamma_correction`
  (geotiff_extractor/src/image_extractor/mod.rs:346-378, 402-422): three f32 bands -> RGBA8 through
  min-max normalisation, gamma 1/2.2 and round-half-away; pinned by the reference's own pure tests
  (mod.rs:517-555, 625-646: gamma(0.5) == 0.7297401, f32_to_u8(0.2, 0.1, 0.3) == 186,
  merging_bands r == 155).
* `raster_to_mat` (homographier/src/homographier/mod.rs:183-220): RGBA -> BGRA swizzle.
* pixel -> ECEF object points `get_world_coordinates` (feature_database/src/elevationdb.rs:64-104):
  GDAL geotransform apply / invert, nearest-pixel elevation lookup, EPSG:4326 -> EPSG:4978 which PROJ
  evaluates in closed form (cart.cpp: N = a / sqrt(1 - es sin^2 phi)); pinned by the reference's
  `coordinate_converter` test (elevationdb.rs:166-178).
"""
import numpy as np

GAMMA_VALUE = np.float32(1.0) / np.float32(2.2)        # const GAMMA_VALUE: f32 = 1.0 / 2.2
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_ES = 2 * WGS84_F - WGS84_F * WGS84_F
DEG_TO_RAD = 0.017453292519943296


def gamma_correction(v):
    """f32::powf(v, 1/2.2) (correctly rounded: evaluated in f64 and rounded once); NaN where v is outside 0..=1"""
    v = np.asarray(v, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        ok = (v >= 0) & (v <= 1)
        out = np.power(np.where(ok, v, 0).astype(np.float64), np.float64(GAMMA_VALUE)).astype(np.float32)
    return np.where(ok, out, np.float32(np.nan))


def f32_to_u8(v, vmin, vmax):
    """.unwrap_or(0) applied: NaN input or a normalised value outside 0..=1 gives 0"""
    v = np.asarray(v, dtype=np.float32)
    vmin, vmax = np.float32(vmin), np.float32(vmax)
    with np.errstate(invalid="ignore", divide="ignore"):
        fl = (v - vmin) / (vmax - vmin)
    g = gamma_correction(fl)
    x = g * np.float32(255.0)
    with np.errstate(invalid="ignore"):
        r = np.floor(x + np.float32(0.5))               # f32::round for x >= 0 (x + 0.5 is exact below 2^23)
    return np.where(np.isnan(g) | np.isnan(v), 0, r).astype(np.uint8)


def band_merger(red, green, blue, min_max):
    """min_max = (red_min, red_max, green_min, green_max, blue_min, blue_max) (f64, cast to f32 as the reference does)"""
    red, green, blue = (np.asarray(b, dtype=np.float32) for b in (red, green, blue))
    out = np.empty(red.shape + (4,), np.uint8)
    out[..., 0] = f32_to_u8(red, min_max[0], min_max[1])
    out[..., 1] = f32_to_u8(green, min_max[2], min_max[3])
    out[..., 2] = f32_to_u8(blue, min_max[4], min_max[5])
    out[..., 3] = np.where(np.isnan(red) & np.isnan(green) & np.isnan(blue), 0, 255)
    return out


def raster_to_mat(rgba, w, h):
    return np.ascontiguousarray(np.asarray(rgba, np.uint8).reshape(h, w, 4)[..., [2, 1, 0, 3]])


def geotransform_apply(gt, x, y):
    return gt[0] + x * gt[1] + y * gt[2], gt[3] + x * gt[4] + y * gt[5]


def geotransform_invert(gt):
    """GDALInvGeoTransform"""
    gt = [float(v) for v in gt]
    if gt[2] == 0.0 and gt[4] == 0.0 and gt[1] != 0.0 and gt[5] != 0.0:
        return [-gt[0] / gt[1], 1.0 / gt[1], 0.0, -gt[3] / gt[5], 0.0, 1.0 / gt[5]]
    det = gt[1] * gt[5] - gt[2] * gt[4]
    mag = max(max(abs(gt[1]), abs(gt[2])), max(abs(gt[4]), abs(gt[5])))
    if abs(det) <= 1e-10 * mag * mag:
        return None
    inv = 1.0 / det
    return [(gt[2] * gt[3] - gt[0] * gt[5]) * inv, gt[5] * inv, -gt[2] * inv,
            (-gt[1] * gt[3] + gt[0] * gt[4]) * inv, -gt[4] * inv, gt[1] * inv]


def geodetic_to_ecef(lat_deg, lon_deg, h):
    phi = np.asarray(lat_deg, np.float64) * DEG_TO_RAD
    lam = np.asarray(lon_deg, np.float64) * DEG_TO_RAD
    h = np.asarray(h, np.float64)
    s = np.sin(phi)
    N = WGS84_A / np.sqrt(1.0 - WGS84_ES * s * s)
    c = np.cos(phi)
    return (N + h) * c * np.cos(lam), (N + h) * c * np.sin(lam), (N * (1.0 - WGS84_ES) + h) * s


def round_half_away(v):
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))


def world_coordinates(px, py, gt_dataset, gt_elevation=None, heights=None, x_size=0, y_size=0):
    """elevationdb.rs:64-90.  Returns (xyz [n,3], ok [n]); ok is False where the elevation pixel does not
    exist (the reference's diesel NotFound)."""
    px, py = np.asarray(px, np.float64), np.asarray(py, np.float64)
    gx, gy = geotransform_apply(gt_dataset, px, py)
    ok = np.ones(px.shape, bool)
    h = np.zeros(px.shape)
    if gt_elevation is not None:
        inv = geotransform_invert(gt_elevation)
        ex, ey = geotransform_apply(inv, gx, gy)
        ix, iy = round_half_away(ex).astype(np.int64), round_half_away(ey).astype(np.int64)
        idx = iy * x_size + ix                       # row id - 1
        ok = (idx >= 0) & (idx < x_size * y_size)
        h = np.where(ok, np.asarray(heights, np.float64).ravel()[np.clip(idx, 0, x_size * y_size - 1)], np.nan)
    X, Y, Z = geodetic_to_ecef(gy, gx, h)            # convert_coordinates(coordinates.1, coordinates.0, height)
    return np.stack([X, Y, Z], -1), ok
