"""Oracle for stage 1 (AKAZE detect + MLDB describe) — TEST INFRASTRUCTURE ONLY.

numpy restatement of `cv::AKAZE::detectAndCompute` as configured by the reference at
feature_extraction/src/lib.rs:64-79: DESCRIPTOR_MLDB, descriptor_size 0, 3 channels, threshold
1e-3, 4 octaves x 4 sublevels, DIFF_PM_G2, max_points.  The arithmetic lives in OpenCV's
features2d (`kaze/AKAZEFeatures.cpp`, `kaze/nldiffusion_functions.cpp`, `kaze/fed.cpp`; not
vendored in the reference — opencv crate 0.88.8 over the system libopencv 4.x); the procedure
is SURVEY.md Appendix A.  Pinned against cv2 4.13.0 `detectAndCompute` output on seeded
synthetic images (tests/golden/akaze_golden.npz, tests/golden/make_golden.py) within the
tolerances stated in tests/test_oracle_akaze.py — parity pinned (tolerance-based: f32 filter
reassociation makes bit equality of keypoints impossible without OpenCV's exact SIMD order).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

# AKAZEOptions defaults + the reference's arguments (lib.rs:64-73)
OMAX, NSUBLEVELS = 4, 4
SOFFSET, DERIVATIVE_FACTOR = 1.6, 1.5
DTHRESHOLD = 1e-3
KCONTRAST_PERCENTILE, KCONTRAST_NBINS = 0.7, 300
PATTERN_SIZE = 10


# ----------------------------------------------------------------------------- primitives
def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F)"""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (k / k.sum()).astype(F)


def _pad(img, r, mode):
    return np.pad(img, r, mode={"replicate": "edge", "reflect101": "reflect"}[mode])


def sep_filter(img: np.ndarray, kx: np.ndarray, ky: np.ndarray, border: str) -> np.ndarray:
    """cv::sepFilter2D(img, CV_32F, kx, ky): kx along x then ky along y, f32 accumulation."""
    img = img.astype(F, copy=False)
    rx, ry = len(kx) // 2, len(ky) // 2
    p = _pad(img, ((0, 0), (rx, rx)), border) if rx else img
    tmp = np.zeros_like(img)
    w = img.shape[1]
    for i, c in enumerate(kx):
        if c != 0:
            tmp += F(c) * p[:, i:i + w]
    p = _pad(tmp, ((ry, ry), (0, 0)), border) if ry else tmp
    out = np.zeros_like(img)
    h = img.shape[0]
    for i, c in enumerate(ky):
        if c != 0:
            out += F(c) * p[i:i + h, :]
    return out


def gaussian_blur(img, ksize, sigma):
    """cv::GaussianBlur(img, (k,k), sigma, sigma, BORDER_REPLICATE)"""
    k = gaussian_kernel(ksize, sigma)
    return sep_filter(img, k, k, "replicate")


def scharr(img, dx, dy):
    """cv::Scharr(img, CV_32F, dx, dy, scale 1, BORDER_DEFAULT) — un-normalised [3,10,3] x [-1,0,1]"""
    d = np.array([-1, 0, 1], F)
    s = np.array([3, 10, 3], F)
    return sep_filter(img, d if dx else s, d if dy else s, "reflect101")


def halfsample_area(img):
    """cv::resize(img, (w/2, h/2), INTER_AREA) for exact 2x decimation = 2x2 box mean; general
    sizes use the fractional-coverage area weights."""
    h, w = img.shape
    nh, nw = h // 2, w // 2
    if h == 2 * nh and w == 2 * nw:
        a = img.astype(F)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) * F(0.25)).astype(F)
    return resize_area(img, nw, nh)


def _area_weights(src, dst):
    """computeResizeAreaTab: fractional coverage weights of dst cell over src cells."""
    scale = src / dst
    W = np.zeros((dst, src), np.float64)
    for d in range(dst):
        fs, fe = d * scale, (d + 1) * scale
        cell = min(scale, src - fs)
        s0, s1 = int(math.ceil(fs)), int(math.floor(fe))
        s1 = min(s1, src)
        s0 = min(s0, s1)
        if s0 - fs > 1e-3:
            W[d, s0 - 1] = (s0 - fs) / cell
        for s in range(s0, s1):
            W[d, s] = 1.0 / cell
        if fe - s1 > 1e-3 and s1 < src:
            W[d, s1] = min(min(fe - s1, 1.0), cell) / cell
    return W.astype(F)


def resize_area(img, nw, nh):
    h, w = img.shape
    Wx, Wy = _area_weights(w, nw), _area_weights(h, nh)
    return (Wy @ img.astype(F) @ Wx.T).astype(F)


# ----------------------------------------------------------------------------- FED
def _is_prime(n):
    if n < 2:
        return False
    for d in range(2, int(math.isqrt(n)) + 1):
        if n % d == 0:
            return False
    return True


def fed_tau_by_process_time(T: float, M: int = 1, tau_max: float = 0.25, reordering: bool = True):
    """kaze/fed.cpp fed_tau_by_process_time -> list of f32 step sizes."""
    t = F(T) / F(M)
    n = int(math.ceil(float(np.sqrt(F(3.0) * t / F(tau_max) + F(0.25)) - F(0.5) - F(1.0e-8))))
    if n <= 0:
        return []
    scale = F(3.0) * t / (F(tau_max) * F(n * (n + 1)))
    c = F(1.0) / (F(4.0) * F(n) + F(2.0))
    d = scale * F(tau_max) / F(2.0)
    tauh = []
    for k in range(n):
        h = F(np.cos(F(np.pi) * (F(2.0) * F(k) + F(1.0)) * c, dtype=F))
        tauh.append(F(d / (h * h)))
    if not reordering:
        return tauh
    kappa = n // 2
    prime = n + 1
    while not _is_prime(prime):
        prime += 1
    tau = []
    k = 0
    for _ in range(n):
        while True:
            index = ((k + 1) * kappa) % prime - 1
            if index < n:
                break
            k += 1
        tau.append(tauh[index])
        k += 1
    return tau


def cv_round(x):
    return int(np.rint(x))


def level_table(width: int, height: int):
    """AKAZEFeatures::Allocate_Memory_Evolution"""
    levels = []
    for o in range(OMAX):
        power = 1 << o
        rf = F(1.0) / F(power)
        lw, lh = int(F(width) * rf), int(F(height) * rf)
        if (lw < 80 or lh < 40) and o != 0:
            break
        for j in range(NSUBLEVELS):
            esigma = F(SOFFSET) * F(math.pow(2.0, F(j) / F(NSUBLEVELS) + o))
            sigma_size = cv_round(float(esigma) * DERIVATIVE_FACTOR / power)
            levels.append(dict(w=lw, h=lh, esigma=float(esigma), sigma_size=sigma_size,
                               etime=float(F(0.5) * (esigma * esigma)), octave=o, sublevel=j, ratio=float(power),
                               border=cv_round(10.0 * math.sqrt(2.0) * sigma_size) + 1))
    for i in range(1, len(levels)):
        levels[i]["tau"] = fed_tau_by_process_time(levels[i]["etime"] - levels[i - 1]["etime"])
    levels[0]["tau"] = []
    return levels


# ----------------------------------------------------------------------------- scale space
def to_gray_f32(img: np.ndarray) -> np.ndarray:
    """cvtColor(BGR2GRAY / BGRA2GRAY) (8-bit fixed point, 15-bit coefficients:
    (B*3735 + G*19235 + R*9798 + 16384) >> 15 — verified bit-exact against cv2 4.13.0)
    then convertTo(CV_32F, 1/255)."""
    if img.ndim == 3 and img.shape[2] in (3, 4):
        b, g, r = (img[..., i].astype(np.int64) for i in range(3))
        gray = ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
    else:
        gray = img.reshape(img.shape[0], img.shape[1])
    return gray.astype(F) * F(1.0 / 255.0)


def compute_kcontrast(Lx, Ly, perc=KCONTRAST_PERCENTILE, nbins=KCONTRAST_NBINS):
    lx, ly = Lx[1:-1, 1:-1], Ly[1:-1, 1:-1]
    modg = np.sqrt(lx * lx + ly * ly).astype(F)
    hmax = F(modg.max())
    if hmax == 0:
        return F(0.03)
    q = (modg * (F(nbins - 1) / hmax)).astype(F)
    hist = np.bincount(q.astype(np.int32).ravel(), minlength=nbins)
    nthreshold = int(F(modg.size - hist[0]) * F(perc))
    nelements = 0
    for k in range(1, nbins):
        if nelements >= nthreshold:
            return F(hmax * F(k) / F(nbins))
        nelements += int(hist[k])
    return F(0.03)


def pm_g2(Lx, Ly, k):
    inv_k = F(1.0) / (F(k) * F(k))
    return (F(1.0) / (F(1.0) + inv_k * (Lx * Lx + Ly * Ly))).astype(F)


def nld_step(Lt, c, step_size):
    """nldiffusion_functions.cpp nld_step_scalar: Lstep = step * sum over 4 neighbours of
    (c + c_nb)(L_nb - L), missing neighbours dropped, the four image corners... (first/last row
    corner pixels are written 0)."""
    L = Lt
    xp = np.zeros_like(L); xn = np.zeros_like(L); yp = np.zeros_like(L); yn = np.zeros_like(L)
    xp[:, :-1] = (c[:, :-1] + c[:, 1:]) * (L[:, 1:] - L[:, :-1])
    xn[:, 1:] = (c[:, 1:] + c[:, :-1]) * (L[:, :-1] - L[:, 1:])
    yp[:-1, :] = (c[:-1, :] + c[1:, :]) * (L[1:, :] - L[:-1, :])
    yn[1:, :] = (c[1:, :] + c[:-1, :]) * (L[:-1, :] - L[1:, :])
    step = ((xp + xn) + yp) + yn
    step = (step * F(step_size)).astype(F)
    step[0, 0] = step[0, -1] = step[-1, 0] = step[-1, -1] = 0
    return step


def derivative_kernels(dx, dy, scale):
    """compute_derivative_kernels: Scharr-like kernels of size 3 + 2(scale-1)."""
    ksize = 3 + 2 * (scale - 1)
    if scale == 1:
        d = np.array([-1, 0, 1], F)
        s = (np.array([3, 10, 3], F) / F(32.0)).astype(F)
        return (d if dx else s), (d if dy else s)
    w = F(10.0) / F(3.0)
    norm = F(1.0) / (F(2.0) * F(scale) * (w + F(2.0)))
    out = []
    for order in (dx, dy):
        k = np.zeros(ksize, F)
        if order == 0:
            k[0], k[ksize // 2], k[-1] = norm, w * norm, norm
        else:
            k[0], k[ksize // 2], k[-1] = -1, 0, 1
        out.append(k)
    return out[0], out[1]


def build_scale_space(img: np.ndarray):
    """Create_Nonlinear_Scale_Space + Compute_Determinant_Hessian_Response.
    Returns levels (dicts with Lt, Lx, Ly, Ldet and the level constants) and kcontrast."""
    gray = to_gray_f32(img)
    h, w = gray.shape
    lv = level_table(w, h)
    lv[0]["Lsmooth"] = gaussian_blur(gray, 9, SOFFSET)
    lv[0]["Lt"] = lv[0]["Lsmooth"].copy()
    kcontrast = F(0.03)
    if len(lv) > 1:
        sm = gaussian_blur(gray, 5, 1.0)
        kcontrast = compute_kcontrast(scharr(sm, 1, 0), scharr(sm, 0, 1))
    k = kcontrast
    for i in range(1, len(lv)):
        e = lv[i]
        if e["octave"] > lv[i - 1]["octave"]:
            Lt = resize_area(lv[i - 1]["Lt"], e["w"], e["h"]) if (lv[i - 1]["w"] != 2 * e["w"] or lv[i - 1]["h"] != 2 * e["h"]) \
                else halfsample_area(lv[i - 1]["Lt"])
            k = F(k * F(0.75))
        else:
            Lt = lv[i - 1]["Lt"].copy()
        e["Lsmooth"] = gaussian_blur(Lt, 5, 1.0)
        Lflow = pm_g2(scharr(e["Lsmooth"], 1, 0), scharr(e["Lsmooth"], 0, 1), k)
        for tau in e["tau"]:
            Lt = (Lt + nld_step(Lt, Lflow, F(tau) * F(0.5))).astype(F)
        e["Lt"] = Lt
    for e in lv:
        s = e["sigma_size"]
        dxkx, dxky = derivative_kernels(1, 0, s)
        dykx, dyky = derivative_kernels(0, 1, s)
        e["Lx"] = sep_filter(e["Lsmooth"], dxkx, dxky, "reflect101")
        e["Ly"] = sep_filter(e["Lsmooth"], dykx, dyky, "reflect101")
        Lxx = sep_filter(e["Lx"], dxkx, dxky, "reflect101")
        Lxy = sep_filter(e["Lx"], dykx, dyky, "reflect101")
        Lyy = sep_filter(e["Ly"], dykx, dyky, "reflect101")
        e["Ldet"] = ((Lxx * Lyy - Lxy * Lxy) * F(s ** 4)).astype(F)
    return lv, float(kcontrast)


# ----------------------------------------------------------------------------- extrema
def local_maxima(e, thr=DTHRESHOLD):
    """strict 3x3 maxima above the detector threshold inside the level's border."""
    L = e["Ldet"]
    h, w = L.shape
    m = np.zeros((h, w), bool)
    b = e["border"]
    if b + 1 >= h or b + 1 >= w:
        return m
    c = L[b:h - b, b:w - b]
    ok = c > F(thr)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                ok &= c > L[b + dy:h - b + dy, b + dx:w - b + dx]
    m[b:h - b, b:w - b] = ok
    return m


def _find_neighbor(mask, x, y, r):
    """first keypoint in a row-major scan of the HALF-OPEN window [y-r, y+r) x [x-r, x+r) that lies
    within L2 distance r (<=).  The half-open bounds matter: a neighbour at exactly +r along an axis
    is not seen (established differentially against cv2 4.13.0 on dense images)."""
    h, w = mask.shape
    for cy in range(max(0, y - r), min(h, y + r)):
        for cx in range(max(0, x - r), min(w, x + r)):
            if mask[cy, cx] and (cx - x) ** 2 + (cy - y) ** 2 <= r * r:
                return cx, cy
    return None


def same_scale_maxima(e, thr=DTHRESHOLD):
    """FindKeypointsSameScale: 3x3 maxima in row-major order; a maximum with an already accepted
    keypoint of the same level within sigma_size replaces it when stronger, else is dropped.
    (Behaviour established differentially against cv2 4.13.0: exact keypoint sets on 13 images.)"""
    raw = local_maxima(e, thr)
    L, r = e["Ldet"], e["sigma_size"]
    m = np.zeros_like(raw)
    ys, xs = np.nonzero(raw)
    for y, x in zip(ys, xs):
        nb = _find_neighbor(m, x, y, r)
        if nb is None:
            m[y, x] = True
        elif L[y, x] > L[nb[1], nb[0]]:
            m[nb[1], nb[0]] = False
            m[y, x] = True
    return m


def find_scale_space_extrema(lv):
    """Find_Scale_Space_Extrema: per-level maxima, then two one-directional passes — a keypoint
    that is STRONGER than its (first found) neighbour in the adjacent level deletes that
    neighbour; a weaker keypoint is left for the other pass to delete."""
    masks = [same_scale_maxima(e) for e in lv]
    n = len(lv)
    for i in range(1, n):                      # against the lower level
        diff = int(lv[i]["ratio"]) // int(lv[i - 1]["ratio"])
        r = lv[i]["sigma_size"] * diff
        ys, xs = np.nonzero(masks[i])
        for y, x in zip(ys, xs):
            nb = _find_neighbor(masks[i - 1], x * diff, y * diff, r)
            if nb is not None and lv[i]["Ldet"][y, x] > lv[i - 1]["Ldet"][nb[1], nb[0]]:
                masks[i - 1][nb[1], nb[0]] = False
    for i in range(n - 2, -1, -1):             # against the upper level
        diff = int(lv[i + 1]["ratio"]) // int(lv[i]["ratio"])
        r = lv[i + 1]["sigma_size"]
        ys, xs = np.nonzero(masks[i])
        for y, x in zip(ys, xs):
            nb = _find_neighbor(masks[i + 1], x // diff, y // diff, r)
            if nb is not None and lv[i]["Ldet"][y, x] > lv[i + 1]["Ldet"][nb[1], nb[0]]:
                masks[i + 1][nb[1], nb[0]] = False
    return masks


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def subpixel_refine(lv, masks):
    """Do_Subpixel_Refinement -> KP_DTYPE array (angle still unset)."""
    out = []
    for i, e in enumerate(lv):
        L = e["Ldet"]
        ratio = F(e["ratio"])
        ys, xs = np.nonzero(masks[i])
        for y, x in zip(ys, xs):
            Dx = F(0.5) * (L[y, x + 1] - L[y, x - 1])
            Dy = F(0.5) * (L[y + 1, x] - L[y - 1, x])
            Dxx = L[y, x + 1] + L[y, x - 1] - F(2.0) * L[y, x]
            Dyy = L[y + 1, x] + L[y - 1, x] - F(2.0) * L[y, x]
            Dxy = F(0.25) * (L[y + 1, x + 1] + L[y - 1, x - 1] - L[y - 1, x + 1] - L[y + 1, x - 1])
            det = Dxx * Dyy - Dxy * Dxy
            if det == 0:
                dx = dy = F(0)      # cv::solve fails -> dst stays 0
            else:
                dx = F((-Dx * Dyy + Dy * Dxy) / det)
                dy = F((-Dy * Dxx + Dx * Dxy) / det)
            if abs(dx) > 1.0 or abs(dy) > 1.0:
                continue
            out.append((F(x) * ratio + dx * ratio + F(0.5) * (ratio - F(1.0)),
                        F(y) * ratio + dy * ratio + F(0.5) * (ratio - F(1.0)),
                        F(2.0) * F(e["esigma"]) * F(DERIVATIVE_FACTOR), 0.0, L[y, x], e["octave"], i))
    return np.array(out, dtype=KP_DTYPE)


# ----------------------------------------------------------------------------- orientation
def _gauss25():
    g = np.zeros((7, 7), F)
    for i in range(7):
        for j in range(7):
            g[i, j] = round(math.exp(-(i * i + j * j) / (2 * 2.5 * 2.5)) / (2 * math.pi * 2.5 * 2.5), 8)
    return g


_G25 = _gauss25()
_ORI = [(i, j, _G25[abs(i), abs(j)]) for i in range(-6, 7) for j in range(-6, 7) if i * i + j * j < 36]
_ORI_X = np.array([o[0] for o in _ORI]); _ORI_Y = np.array([o[1] for o in _ORI])
_ORI_W = np.array([o[2] for o in _ORI], F)

_P1 = F(0.9997878412794807) * F(180 / np.pi)
_P3 = F(-0.3258083974640975) * F(180 / np.pi)
_P5 = F(0.1555786518463281) * F(180 / np.pi)
_P7 = F(-0.04432655554792128) * F(180 / np.pi)


def fast_atan2_deg(y, x):
    """cv::hal::fastAtan32f polynomial (degrees), vectorised."""
    y = np.asarray(y, F); x = np.asarray(x, F)
    ax, ay = np.abs(x), np.abs(y)
    eps = F(np.finfo(np.float64).eps)
    swap = ax < ay
    num = np.where(swap, ax, ay); den = np.where(swap, ay, ax)
    c = (num / (den + eps)).astype(F)
    c2 = c * c
    a = ((((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c).astype(F)
    a = np.where(swap, F(90.0) - a, a)
    a = np.where(x < 0, F(180.0) - a, a)
    a = np.where(y < 0, F(360.0) - a, a)
    return a.astype(F)


def get_angle(x, y):
    """kaze/utils.h getAngle"""
    x, y = F(x), F(y)
    if x >= 0 and y >= 0:
        return F(np.arctan(y / x)) if x != 0 or y != 0 else F(0)
    if x < 0 and y >= 0:
        return F(np.pi) - F(np.arctan(-y / x))
    if x < 0 and y < 0:
        return F(np.pi) + F(np.arctan(y / x))
    return F(2.0 * np.pi) - F(np.arctan(-y / x))


def main_orientation(kp, lv):
    """Compute_Main_Orientation -> angle in degrees."""
    e = lv[int(kp["class_id"])]
    ratio = F(e["ratio"])
    scale = cv_round(F(0.5) * kp["size"] / ratio)
    x0, y0 = cv_round(kp["x"] / ratio), cv_round(kp["y"] / ratio)
    ys, xs = y0 + _ORI_Y * scale, x0 + _ORI_X * scale
    resX = (_ORI_W * e["Lx"][ys, xs]).astype(F)
    resY = (_ORI_W * e["Ly"][ys, xs]).astype(F)
    ang = (fast_atan2_deg(resY, resX) * F(np.pi / 180.0)).astype(F)
    slices, win = 42, 7
    step = F(2.0 * np.pi / slices)
    b = (ang / step).astype(np.int32)
    b[(b < 0) | (b >= slices)] = 0
    # quantized_counting_sort: within a slice indices end up in DESCENDING order
    order = np.lexsort((-np.arange(len(b)), b))
    cum = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=slices))])

    def wsum(lo, hi):
        sx = F(0); sy = F(0)
        for idx in order[lo:hi]:
            sx = F(sx + resX[idx]); sy = F(sy + resY[idx])
        return sx, sy

    maxX, maxY = wsum(cum[0], cum[win])
    maxN = maxX * maxX + maxY * maxY
    for sn in range(1, slices - win + 1):
        if cum[sn] == cum[sn - 1] and cum[sn + win] == cum[sn + win - 1]:
            continue
        sx, sy = wsum(cum[sn], cum[sn + win])
        n = sx * sx + sy * sy
        if n > maxN:
            maxN, maxX, maxY = n, sx, sy
    for sn in range(slices - win + 1, slices):
        remain = sn + win - slices
        if cum[sn] == cum[sn - 1] and cum[remain] == cum[remain - 1]:
            continue
        sx, sy = wsum(cum[sn], cum[slices])
        sx2, sy2 = wsum(cum[0], cum[remain])
        # one running sum in the source: continue accumulating
        sx = F(sx); sy = F(sy)
        for idx in order[cum[0]:cum[remain]]:
            sx = F(sx + resX[idx]); sy = F(sy + resY[idx])
        n = sx * sx + sy * sy
        if n > maxN:
            maxN, maxX, maxY = n, sx, sy
    # cv2 4.13.0 stores fastAtan2(maxY, maxX) (degrees); verified differentially: with the exact
    # atan the angles are off by up to 0.0092 deg and 10 % of descriptors differ by a few bits
    return F(fast_atan2_deg(np.array([maxY], F), np.array([maxX], F))[0])


# ----------------------------------------------------------------------------- MLDB
def mldb_descriptor(kp, lv):
    """MLDB_Full_Descriptor_Invoker::Get_MLDB_Full_Descriptor -> 61 bytes."""
    e = lv[int(kp["class_id"])]
    Lt, Lx, Ly = e["Lt"], e["Lx"], e["Ly"]
    h, w = Lt.shape
    ratio = F(1 << int(kp["octave"]))
    scale = cv_round(F(0.5) * kp["size"] / ratio)
    xf, yf = F(kp["x"] / ratio), F(kp["y"] / ratio)
    angle = F(kp["angle"] * F(np.pi / 180.0))
    co, si = F(np.cos(angle)), F(np.sin(angle))
    desc = np.zeros(61, np.uint8)
    dpos = 0
    for z, step in enumerate((PATTERN_SIZE, (PATTERN_SIZE * 2 + 2) // 3, (PATTERN_SIZE + 1) // 2)):
        vals = []
        starts = list(range(-PATTERN_SIZE, PATTERN_SIZE, step))
        for i in starts:
            for j in starts:
                k = np.arange(i, i + step, dtype=np.int32)[:, None]
                l = np.arange(j, j + step, dtype=np.int32)[None, :]
                kf, lf = k.astype(F), l.astype(F)
                sy = yf + ((lf * co) * F(scale) + (kf * si) * F(scale))
                sx = xf + ((-lf * si) * F(scale) + (kf * co) * F(scale))
                y1 = np.rint(sy).astype(np.int64).ravel()
                x1 = np.rint(sx).astype(np.int64).ravel()
                ok = (y1 >= 0) & (y1 < h) & (x1 >= 0) & (x1 < w)
                y1, x1 = y1[ok], x1[ok]
                di = dx = dy = F(0)
                ri, rx, ry = Lt[y1, x1], Lx[y1, x1], Ly[y1, x1]
                rry = rx * co + ry * si
                rrx = -rx * si + ry * co
                for t in range(len(ri)):            # sequential f32 accumulation like the source
                    di = F(di + ri[t]); dx = F(dx + rrx[t]); dy = F(dy + rry[t])
                if len(ri):
                    inv = F(1.0) / F(len(ri))
                    di, dx, dy = F(di * inv), F(dx * inv), F(dy * inv)
                vals.append((di, dx, dy))
        v = np.array(vals, F)
        n = len(vals)
        for c in range(3):
            for a in range(n):
                for b2 in range(a + 1, n):
                    if v[a, c] > v[b2, c]:
                        desc[dpos >> 3] |= 1 << (dpos & 7)
                    dpos += 1
    return desc


def detect_and_compute(img, max_points=(1 << 18) - 1, want_desc=True):
    """AKAZE::detectAndCompute -> (KP_DTYPE array, N x 61 u8)."""
    lv, _ = build_scale_space(img)
    masks = find_scale_space_extrema(lv)
    kps = subpixel_refine(lv, masks)
    if 0 < max_points < len(kps):
        keep = np.argsort(-kps["response"], kind="stable")[:max_points]
        kps = kps[keep]
    desc = np.zeros((len(kps), 61), np.uint8)
    for i in range(len(kps)):
        kps["angle"][i] = main_orientation(kps[i], lv)
        if want_desc:
            desc[i] = mldb_descriptor(kps[i], lv)
    return kps, desc
