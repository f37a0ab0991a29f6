"""Host-side mirror of the radiometric pre-step of `geotiff_lib::image_extractor`
(geotiff_extractor/src/image_extractor/mod.rs:346-378, 402-422): `BandsMinMax`, `band_merger`,
`f32_to_u8`, `gamma_correction`.  GDAL raster I/O itself is out of scope; the arithmetic — three f32
bands to RGBA8 — runs on the GPU through `dunk_band_merger`."""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DunkError, check, default_context, ptr


class BandsMinMax(NamedTuple):
    """mod.rs:110-118"""
    red_min: float
    red_max: float
    green_min: float
    green_max: float
    blue_min: float
    blue_max: float


class PixelConversion(Exception):
    """mod.rs:120-125 — kinds: GammaOutOfRange, FloatToIntegerError, NotANumber"""

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind


def band_merger(bands: Sequence[np.ndarray], min_max: BandsMinMax, bgra: bool = False,
                ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """mod.rs:346-378 — bands in the order R, G, B (equal-length f32 vectors) -> [n, 4] RGBA8.
    `bgra=True` fuses homographier::raster_to_mat's channel swizzle (mod.rs:183-220)."""
    ctx = ctx or default_context()
    r, g, b = (np.ascontiguousarray(x, dtype=np.float32).ravel() for x in bands[:3])
    if not (r.shape == g.shape == b.shape):
        raise DunkError(_lib.ERR_VEC_LENGTH, "band_merger: bands differ in length")
    out = np.empty((r.shape[0], 4), dtype=np.uint8)
    mm = np.ascontiguousarray(min_max, dtype=np.float64)
    check(_lib.load().dunk_band_merger(ctx.handle, ptr(r), ptr(g), ptr(b), r.shape[0], ptr(mm), int(bool(bgra)), ptr(out)))
    return out


def f32_to_u8(input_value: float, min: float, max: float, ctx: Optional[_lib.Context] = None) -> int:
    """mod.rs:410-422 — Err(NotANumber) / Err(GammaOutOfRange) become PixelConversion"""
    v = np.float32(input_value)
    if np.isnan(v):
        raise PixelConversion("NotANumber")
    fl = (v - np.float32(min)) / (np.float32(max) - np.float32(min))
    if not (0.0 <= fl <= 1.0):
        raise PixelConversion("GammaOutOfRange")
    px = band_merger([[v], [v], [v]], BandsMinMax(min, max, min, max, min, max), ctx=ctx)
    return int(px[0, 0])


def gamma_correction(input_value: float, ctx: Optional[_lib.Context] = None) -> float:
    """mod.rs:402-408 — only the 8-bit result is observable through the device kernel; the f32 value is
    returned from the same correctly-rounded evaluation powf(v, 1/2.2)"""
    v = np.float32(input_value)
    if not (0.0 <= v <= 1.0):
        raise PixelConversion("GammaOutOfRange")
    return float(np.float32(np.power(np.float64(v), np.float64(np.float32(1.0) / np.float32(2.2)))))
