"""Synthetic imagery for benchmarks and tests: multi-scale smoothed noise tiles and
known-homography warps.  Input generation for bench.py / tests only (not the product; may use
cv2.resize for speed when it is importable)."""
from __future__ import annotations

import functools

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


# ------------------------------------------------------------------------------------------------------------
# BASELINE config 5 geometry: the config-4 scene is a Sentinel-2-like tile (10980 x 10980 px of 10 m) with a GDAL
# geotransform and a DEM; every query frame is what a nadir-ish pinhole camera 500 km up sees of a 1024^2 window.
# Input generation only (numpy); the product computes world coordinates with its own kernel.
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_ES = 2 * WGS84_F - WGS84_F * WGS84_F
SCENE_LAT0, SCENE_LON0, SCENE_GSD = 57.05, 9.92, 10.0        # Aalborg; metres per pixel
CAMERA_F = 50000.0                                          # pixels: 500 km altitude at 10 m GSD
CAMERA_K = np.array([[CAMERA_F, 0.0, 512.0], [0.0, CAMERA_F, 512.0], [0.0, 0.0, 1.0]])
DEM_CELL = 10                                               # scene pixels per DEM cell (100 m posts)


def scene_geotransform():
    """GDAL geotransform of the scene: x -> lon, y -> lat (north up), ~10 m square pixels at SCENE_LAT0"""
    dx = SCENE_GSD / (111320.0 * np.cos(np.radians(SCENE_LAT0)))
    dy = SCENE_GSD / 110574.0
    return np.array([SCENE_LON0, dx, 0.0, SCENE_LAT0, 0.0, -dy])


@functools.lru_cache(maxsize=4)
def scene_dem(size: int):
    """smooth synthetic relief (+-60 m) on 100 m posts covering the scene; returns (geotransform, heights [ny, nx]).
    Seen from 500 km with <= 2 degrees off-nadir a 60 m relief moves a point by <= 60 m * tan(2 deg) = 0.2 px, so a
    homography-warped frame is consistent with the camera viewing this terrain to below the matcher's noise, while
    the object points are far enough from coplanar for EPnP (an exact plane makes its null space degenerate)."""
    gt = scene_geotransform()
    n = size // DEM_CELL + 2
    gt_e = np.array([gt[0], gt[1] * DEM_CELL, 0.0, gt[3], 0.0, gt[5] * DEM_CELL])
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    h = 45.0 * np.sin(xx / 23.0) * np.cos(yy / 17.0) + 15.0 * np.sin(xx / 5.3 + 1.0) * np.sin(yy / 6.1)
    return gt_e, np.ascontiguousarray(h)


def geodetic_to_ecef(lat_deg, lon_deg, h):
    phi, lam = np.radians(np.asarray(lat_deg, np.float64)), np.radians(np.asarray(lon_deg, np.float64))
    s, c = np.sin(phi), np.cos(phi)
    N = WGS84_A / np.sqrt(1.0 - WGS84_ES * s * s)
    return np.stack([(N + h) * c * np.cos(lam), (N + h) * c * np.sin(lam), (N * (1.0 - WGS84_ES) + h) * s], -1)


def scene_points_ecef(px, py, size: int):
    """scene pixel -> ECEF through the same chain the reference uses (geotransform -> nearest DEM post -> EPSG:4978)"""
    gt = scene_geotransform()
    gt_e, h = scene_dem(size)
    px, py = np.asarray(px, np.float64), np.asarray(py, np.float64)
    lon, lat = gt[0] + px * gt[1], gt[3] + py * gt[5]
    ex, ey = (lon - gt_e[0]) / gt_e[1], (lat - gt_e[3]) / gt_e[5]
    ix = np.clip(np.floor(ex + 0.5).astype(np.int64), 0, h.shape[1] - 1)
    iy = np.clip(np.floor(ey + 0.5).astype(np.int64), 0, h.shape[0] - 1)
    return geodetic_to_ecef(lat, lon, h[iy, ix])


def scene_origin(size: int):
    """ECEF of the scene centre at height 0: the local origin subtracted before the f32 rounding of PnP inputs"""
    gt = scene_geotransform()
    return geodetic_to_ecef(gt[3] + 0.5 * size * gt[5], gt[0] + 0.5 * size * gt[1], 0.0)


def camera_view(x0: float, y0: float, seed: int, size: int):
    """A camera looking at the scene window with top-left (x0, y0).  Returns (H scene px -> frame px, R, t, fit
    residual px): camera-frame point = R (X_ecef - origin) + t; H is the least-squares homography through the
    projections of a 9 x 9 grid of window points (DEM heights included), i.e. the planar approximation of the view."""
    r = np.random.default_rng(seed)
    gt = scene_geotransform()
    origin = scene_origin(size)
    cx, cy = x0 + 512.0, y0 + 512.0
    lat, lon = gt[3] + cy * gt[5], gt[0] + cx * gt[1]
    target = geodetic_to_ecef(lat, lon, 0.0)
    phi, lam = np.radians(lat), np.radians(lon)
    east = np.array([-np.sin(lam), np.cos(lam), 0.0])
    north = np.array([-np.sin(phi) * np.cos(lam), -np.sin(phi) * np.sin(lam), np.cos(phi)])
    up = np.array([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)])
    alt = CAMERA_F * SCENE_GSD * (1.0 + r.uniform(-0.05, 0.05))          # +-5 % scale
    tilt, az, yaw = np.radians(r.uniform(0.0, 2.0)), r.uniform(0, 2 * np.pi), r.uniform(-0.12, 0.12)
    C = target + alt * (up + np.tan(tilt) * (np.cos(az) * east + np.sin(az) * north))
    z = (target - C) / np.linalg.norm(target - C)                         # optical axis, towards the ground
    x_ref = np.cos(yaw) * east - np.sin(yaw) * north                      # image x ~ east, image y ~ south
    y = np.cross(z, x_ref); y /= np.linalg.norm(y)
    x = np.cross(y, z)
    R = np.stack([x, y, z])
    t = -R @ (C - origin)
    g = np.linspace(0.0, 1023.0, 9)
    gx, gy = np.meshgrid(x0 + g, y0 + g)
    P = (scene_points_ecef(gx.ravel(), gy.ravel(), size) - origin) @ R.T + t
    uv = np.stack([CAMERA_F * P[:, 0] / P[:, 2] + 512.0, CAMERA_F * P[:, 1] / P[:, 2] + 512.0], 1)
    # least-squares homography (DLT on normalised coordinates) scene px -> frame px
    src = np.stack([gx.ravel() - cx, gy.ravel() - cy], 1) / 512.0
    dst = (uv - 512.0) / 512.0
    A = []
    for (sx, sy), (u, v) in zip(src, dst):
        A.append([-sx, -sy, -1, 0, 0, 0, u * sx, u * sy, u])
        A.append([0, 0, 0, -sx, -sy, -1, v * sx, v * sy, v])
    h = np.linalg.svd(np.asarray(A))[2][-1].reshape(3, 3)
    Tn_src = np.array([[1 / 512.0, 0, -cx / 512.0], [0, 1 / 512.0, -cy / 512.0], [0, 0, 1.0]])
    Tn_dst = np.array([[512.0, 0, 512.0], [0, 512.0, 512.0], [0, 0, 1.0]])
    H = Tn_dst @ h @ Tn_src
    H /= H[2, 2]
    q = np.c_[gx.ravel(), gy.ravel(), np.ones(81)] @ H.T
    resid = float(np.abs(q[:, :2] / q[:, 2:] - uv).max())
    return H, R, t, resid


def config5_views(n: int, size: int, seed0: int):
    """n camera views of random windows of the scene: (H [n,3,3] scene->frame, R [n,3,3], t [n,3], max fit residual)"""
    rng = np.random.default_rng(seed0)
    Hs, Rs, ts, worst = [], [], [], 0.0
    for i in range(n):
        x0, y0 = rng.uniform(96, size - 1024 - 96, 2)
        H, R, t, res = camera_view(x0, y0, seed0 + 1 + i, size)
        Hs.append(H); Rs.append(R); ts.append(t); worst = max(worst, res)
    return np.stack(Hs), np.stack(Rs), np.stack(ts), worst


def pose_errors(rvecs, tvecs, found, Rs, ts):
    """(rotation error in degrees, camera-centre error in metres) of recovered poses; inf where not found"""
    rot, pos = [], []
    for rv, tv, ok, R, t in zip(rvecs, tvecs, found, Rs, ts):
        if not ok:
            rot.append(np.inf); pos.append(np.inf); continue
        th = np.linalg.norm(rv)
        if th < 1e-12:
            Re = np.eye(3)
        else:
            k = rv / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            Re = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
        rot.append(float(np.degrees(np.arccos(np.clip((np.trace(Re @ R.T) - 1) / 2, -1, 1)))))
        pos.append(float(np.linalg.norm(-Re.T @ tv + R.T @ t)))
    return np.array(rot), np.array(pos)
