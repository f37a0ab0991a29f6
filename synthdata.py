"""Synthetic imagery for benchmarks and tests: multi-scale smoothed noise tiles and
known-homography warps.  Input generation for bench.py / tests only (not the product; may use
cv2.resize for speed when it is importable)."""
from __future__ import annotations

import numpy as np

# SURVEY 8d config 1
H_CONFIG1 = np.array([[0.98, -0.12, 60.0], [0.10, 1.03, -40.0], [1e-5, -2e-5, 1.0]])


def _upsample(a: np.ndarray, h: int, w: int) -> np.ndarray:
    """bicubic up-sampling of a coarse noise field to (h, w)"""
    try:   # SURVEY 8d recipe uses cv2.resize(INTER_CUBIC); much faster for multi-megapixel scenes
        import cv2 as _cv
        return _cv.resize(a, (w, h), interpolation=_cv.INTER_CUBIC)
    except Exception:
        from scipy import ndimage
        return ndimage.zoom(a, (h / a.shape[0], w / a.shape[1]), order=3, mode="nearest", grid_mode=True)[:h, :w]


def synth_image(h: int, w: int | None = None, seed: int = 0) -> np.ndarray:
    """SURVEY 8d `synth`: sum over s in {1,2,4,8,16,32} of resize(N(0,1)[(n/s+1)^2], cubic) * sqrt(s),
    min-max normalised to u8 (~2-3 k AKAZE keypoints per 1024^2)."""
    w = w or h
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float64)
    for s in (1, 2, 4, 8, 16, 32):
        n = rng.standard_normal((h // s + 1, w // s + 1))
        acc += _upsample(n, h, w) * np.sqrt(s)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return (acc * 255).astype(np.uint8)


def synth_scene(size: int, seed: int = 11) -> np.ndarray:
    """Large scene for configs 4/5: the same multi-scale noise field as `synth_image`, but mapped to u8
    with a fixed contrast (mean +- 4.6 sigma -> 0..255, the range/sigma ratio a 1024^2 `synth_image` has)
    instead of a global min-max, so that a 1024^2 window carries the keypoint density of a tile."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((size, size), np.float32)
    for s in (1, 2, 4, 8, 16, 32):
        n = rng.standard_normal((size // s + 1, size // s + 1)).astype(np.float32)
        acc += _upsample(n, size, size).astype(np.float32) * np.float32(np.sqrt(s))
    acc = (acc - acc.mean()) / acc.std()
    return np.clip(np.rint(127.5 + acc * (255.0 / 9.2)), 0, 255).astype(np.uint8)


def warp_perspective(img: np.ndarray, H: np.ndarray, out_h: int, out_w: int, border: int = 1) -> np.ndarray:
    """dst(x, y) = src(H^-1 (x, y)) with bilinear interpolation and a constant border (the reference
    warps with BORDER_CONSTANT value 1, homographier mod.rs:286-294).  H maps src -> dst."""
    Hi = np.linalg.inv(H)
    ys, xs = np.mgrid[0:out_h, 0:out_w].astype(np.float64)
    d = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
    sx = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / d
    sy = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / d
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    fx = (sx - x0).astype(np.float32); fy = (sy - y0).astype(np.float32)
    h, w = img.shape[:2]
    src = img.astype(np.float32)

    def at(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(ok, v, np.float32(border))
    out = (at(y0, x0) * (1 - fx) + at(y0, x0 + 1) * fx) * (1 - fy) + (at(y0 + 1, x0) * (1 - fx) + at(y0 + 1, x0 + 1) * fx) * fy
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def window_homography(x0: float, y0: float, seed: int, jitter: float = 1.0) -> np.ndarray:
    """scene -> frame homography of a query frame looking at the scene window with top-left (x0, y0):
    small rotation / scale / perspective around the window (config 5: "known-H warp of a window")."""
    r = np.random.default_rng(seed)
    ang = r.uniform(-0.12, 0.12) * jitter
    s = 1.0 + r.uniform(-0.05, 0.05) * jitter
    A = np.array([[s * np.cos(ang), -s * np.sin(ang), 0.0], [s * np.sin(ang), s * np.cos(ang), 0.0],
                  [r.uniform(-2e-5, 2e-5) * jitter, r.uniform(-2e-5, 2e-5) * jitter, 1.0]])
    T = np.array([[1.0, 0, -x0], [0, 1.0, -y0], [0, 0, 1.0]])
    C = np.array([[1.0, 0, -512.0], [0, 1.0, -512.0], [0, 0, 1.0]])
    return np.linalg.inv(C) @ A @ C @ T


L2_CASES = {"a": (300, 5000, 64), "b": (77, 1031, 128), "c": (129, 257, 64)}


def l2_descriptors(name: str):
    """Seeded unit-norm f32 descriptor sets of the float-matcher golden cases (tests/golden/l2_golden.npz
    stores only cv2's answers): queries are noisy copies of train rows; case "c" plants exact duplicates."""
    nq, nt, dim = L2_CASES[name]
    rng = np.random.default_rng(31 + ord(name))
    t = rng.normal(size=(nt, dim)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    q = (t[rng.integers(0, nt, nq)] + rng.normal(0, 0.08, (nq, dim))).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if name == "c":
        t[100:104] = t[7]                      # distance ties -> lower train index
        q[:8] = t[7]
    return np.ascontiguousarray(q), np.ascontiguousarray(t)
