"""Oracle for the level-of-detail tiler — TEST INFRASTRUCTURE ONLY.

Restates preprocessor/src/main.rs:197-327: tile = scene >> (lods - 1); per lod the
(tile << lod)-sized windows are read at tile resolution, converted by band_merger and handed to the
extractor; keypoints go back to scene pixels as x * 2^lod + window offset; one ref_image row per tile.
The reference reads the windows through GDAL `read_as(.., Some(ResampleAlg::Lanczos))`
(geotiff_extractor/src/image_extractor/mod.rs:332-343), whose pixels also depend on GDAL's overview
selection inside the COG — GDAL is not available here and the reference holds no fixture for it:
PARITY UNPINNED for the GDAL / Lanczos pixel values; the box mean is pinned against cv2.resize(INTER_AREA) to one f32
ulp (tests/test_oracle_lod.py).  Two deterministic resamplers are restated, both with
f64 accumulation in row-major tap order and renormalisation at the scene border: the box mean and the
Lanczos-3 convolution stretched by the decimation factor (the kernel GDAL's RasterIO uses when
down-sampling)."""
import math

import numpy as np

from . import geo_oracle as go


def lanczos3(x):
    if x == 0.0:
        return 1.0
    if abs(x) >= 3.0:
        return 0.0
    px = math.pi * x
    return 3.0 * math.sin(px) * math.sin(px / 3.0) / (px * px)


def taps(scale, resample):
    if scale == 1:
        return 0, [1.0]
    if resample == "area":
        return -(scale // 2), [1.0] * scale
    first = -3 * scale
    return first, [lanczos3((k + first + 0.5) / scale) for k in range(6 * scale)]


def resample_window(band, x0, y0, tile_w, tile_h, scale, resample):
    """band [H, W] f32; window origin (x0, y0) in scene pixels; returns [tile_h, tile_w] f32"""
    H, W = band.shape
    if scale == 1:
        return band[y0:y0 + tile_h, x0:x0 + tile_w].copy()
    if resample == "area" and x0 + tile_w * scale <= W and y0 + tile_h * scale <= H:
        # box mean of an interior window: the same sequential f64 sums as the general path below (unit weights,
        # taps kx = 0..scale-1 then ky = 0..scale-1), formed with strided slices instead of per-tap gathers
        win = band[y0:y0 + tile_h * scale, x0:x0 + tile_w * scale].astype(np.float64)
        rows = np.zeros((tile_h * scale, tile_w))
        for kx in range(scale):
            rows = rows + win[:, kx::scale]
        rows = rows / float(scale)
        out = np.zeros((tile_h, tile_w))
        for ky in range(scale):
            out = out + rows[ky::scale]
        return (out / float(scale)).astype(np.float32)
    first, w = taps(scale, resample)
    cx = x0 + np.arange(tile_w) * scale + scale // 2
    cy = y0 + np.arange(tile_h) * scale + scale // 2
    b64 = band.astype(np.float64)
    # horizontal pass per source row that is needed, sequential tap order
    ys = np.arange(cy[0] + first, cy[-1] + first + len(w))
    ys_ok = (ys >= 0) & (ys < H)
    rows = np.zeros((len(ys), tile_w))
    wsum_x = np.zeros(tile_w)
    for kx, wk in enumerate(w):
        xx = cx + first + kx
        ok = (xx >= 0) & (xx < W)
        wsum_x = wsum_x + np.where(ok, wk, 0.0)
        vals = b64[np.clip(ys, 0, H - 1)][:, np.clip(xx, 0, W - 1)]
        rows = rows + np.where(ok[None, :], wk * vals, 0.0)
    rows = rows / wsum_x[None, :]
    out = np.zeros((tile_h, tile_w))
    wsum_y = np.zeros(tile_h)
    for ky, wk in enumerate(w):
        yy = cy + first + ky
        ok = (yy >= 0) & (yy < H)
        wsum_y = wsum_y + np.where(ok, wk, 0.0)
        idx = np.clip(yy - ys[0], 0, len(ys) - 1)
        out = out + np.where(ok[:, None], wk * rows[idx], 0.0)
    return (out / wsum_y[:, None]).astype(np.float32)


def lod_tiles(red, green, blue, min_max, lods, resample="area"):
    """yields (lod, col, row, x_start, y_start, scale, BGRA tile [tile_h, tile_w, 4] u8) in the device's order"""
    H, W = red.shape
    tw, th = W >> (lods - 1), H >> (lods - 1)
    for lod in range(lods):
        s = 1 << lod
        for row in range(H // (th * s)):
            for col in range(W // (tw * s)):
                x0, y0 = col * tw * s, row * th * s
                b = [resample_window(x, x0, y0, tw, th, s, resample) for x in (red, green, blue)]
                rgba = go.band_merger(b[0], b[1], b[2], min_max)
                yield lod, col, row, x0, y0, s, np.ascontiguousarray(rgba[..., [2, 1, 0, 3]])
