"""Oracle for stage 3 (RANSAC homography) — TEST INFRASTRUCTURE ONLY.

Restates `cv::findHomography(src, dst, RANSAC, thr)` as the reference calls it
(homographier/src/homographier/mod.rs:231-259: 5-argument overload => maxIters 2000,
confidence 0.995).  The arithmetic lives in OpenCV's calib3d (`ptsetreg.cpp`, `fundam.cpp`,
`levmarq.cpp`; not vendored in the reference — opencv crate 0.88.8 over the system libopencv
4.x).  The procedure below is SURVEY.md Appendix C; it is pinned against cv2 4.13.0 outputs in
tests/golden/ransac_golden.npz (tests/golden/make_golden.py) — parity pinned (masks bit-exact,
H within 1e-4 relative).
"""
from __future__ import annotations

import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)
CV_RNG_COEFF = 4164903690


class CvRNG:
    """cv::RNG — multiply-with-carry; RANSACPointSetRegistrator seeds it with (uint64)-1."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * CV_RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else a + self.next() % (b - a)


def have_collinear_points(p, count):
    """modules/calib3d/src/precomp.hpp haveCollinearPoints: only the LAST point is tested."""
    i = count - 1
    for j in range(i):
        dx1 = float(p[j, 0]) - float(p[i, 0])
        dy1 = float(p[j, 1]) - float(p[i, 1])
        for k in range(j):
            dx2 = float(p[k, 0]) - float(p[i, 0])
            dy2 = float(p[k, 1]) - float(p[i, 1])
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2])
            + a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]))


def check_subset(s, d, count=4):
    if have_collinear_points(s, count) or have_collinear_points(d, count):
        return False
    if count == 4:
        neg = 0
        for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
            A = [[float(s[i, 0]), float(s[i, 1]), 1.0] for i in t]
            B = [[float(d[i, 0]), float(d[i, 1]), 1.0] for i in t]
            neg += (_det3(A) * _det3(B)) < 0
        if neg != 0 and neg != 4:
            return False
    return True


def get_subset(src, dst, rng, max_attempts=10000, model_points=4):
    n = src.shape[0]
    for _ in range(max_attempts):
        idx = []
        for i in range(model_points):
            v = rng.uniform(0, n)
            while v in idx:
                v = rng.uniform(0, n)
            idx.append(v)
        if check_subset(src[idx], dst[idx], model_points):
            return idx
    return None


def dlt_homography(src, dst):
    """HomographyEstimatorCallback::runKernel — normalised DLT, LtL 9x9, smallest eigenvector."""
    M = src.astype(np.float64)
    m = dst.astype(np.float64)
    n = M.shape[0]
    cM, cm = M.sum(0) / n, m.sum(0) / n
    sM, sm = np.abs(M - cM).sum(0), np.abs(m - cm).sum(0)
    if (np.abs(sM) < DBL_EPSILON).any() or (np.abs(sm) < DBL_EPSILON).any():
        return None
    sM, sm = n / sM, n / sm
    inv_hnorm = np.array([[1 / sm[0], 0, cm[0]], [0, 1 / sm[1], cm[1]], [0, 0, 1]])
    hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    x, y = (m[:, 0] - cm[0]) * sm[0], (m[:, 1] - cm[1]) * sm[1]
    X, Y = (M[:, 0] - cM[0]) * sM[0], (M[:, 1] - cM[1]) * sM[1]
    z, o = np.zeros(n), np.ones(n)
    Lx = np.stack([X, Y, o, z, z, z, -x * X, -x * Y, -x], 1)
    Ly = np.stack([z, z, z, X, Y, o, -y * X, -y * Y, -y], 1)
    LtL = Lx.T @ Lx + Ly.T @ Ly
    w, v = np.linalg.eigh(LtL)
    h0 = v[:, 0].reshape(3, 3)
    H = inv_hnorm @ h0 @ hnorm2
    if H[2, 2] == 0:
        return None
    return H / H[2, 2]


def reproj_err_f32(H, src, dst):
    """HomographyEstimatorCallback::computeError — everything in f32, no FMA contraction."""
    Hf = H.astype(np.float32).ravel()
    x, y = src[:, 0].astype(np.float32), src[:, 1].astype(np.float32)
    one = np.float32(1.0)
    ww = one / ((Hf[6] * x + Hf[7] * y) + one)
    dx = ((Hf[0] * x + Hf[1] * y) + Hf[2]) * ww - dst[:, 0].astype(np.float32)
    dy = ((Hf[3] * x + Hf[4] * y) + Hf[5]) * ww - dst[:, 1].astype(np.float32)
    return dx * dx + dy * dy


def find_inliers(H, src, dst, thr):
    err = reproj_err_f32(H, src, dst)
    mask = err <= np.float32(thr * thr)
    return mask, int(mask.sum())


def cv_round(x):
    return int(np.rint(x))      # round-half-even, like cvRound (SSE cvtsd2si)


def ransac_update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < DBL_MIN:
        return 0
    num, denom = np.log(num), np.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else cv_round(num / denom)


def hypothesis_stream(src, dst, count):
    """The first `count` minimal sample index sets of OpenCV's fixed-seed RANSAC stream (no early
    stop) — the "identical seeded hypothesis sets" of the parity contract."""
    rng = CvRNG()
    out = []
    for _ in range(count):
        idx = get_subset(src, dst, rng)
        if idx is None:
            break
        out.append(idx)
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def score_hypotheses(src, dst, samples, thr):
    """per-hypothesis inlier count and minimal-sample H (NaN where the solve fails)."""
    counts = np.zeros(len(samples), dtype=np.int32)
    Hs = np.full((len(samples), 9), np.nan)
    for i, idx in enumerate(samples):
        H = dlt_homography(src[idx], dst[idx])
        if H is None:
            continue
        Hs[i] = H.ravel()
        counts[i] = find_inliers(H, src, dst, thr)[1]
    return counts, Hs


def ransac_loop(src, dst, thr, max_iters=2000, confidence=0.995):
    """RANSACPointSetRegistrator::run.  Returns (H, mask, iterations actually run) or None."""
    n = src.shape[0]
    rng = CvRNG()
    niters = max_iters
    best_count, best_H, best_mask = 0, None, None
    it = 0
    while it < niters:
        idx = get_subset(src, dst, rng)
        if idx is None:
            if it == 0:
                return None
            break
        H = dlt_homography(src[idx], dst[idx])
        if H is not None:
            mask, good = find_inliers(H, src, dst, thr)
            if good > max(best_count, 3):
                best_count, best_H, best_mask = good, H, mask
                niters = ransac_update_num_iters(confidence, (n - good) / n, 4, niters)
        it += 1
    if best_H is None:
        return None
    return best_H, best_mask, it


def _refine_residual_jac(h8, src, dst, want_jac=True):
    """HomographyRefineCallback::compute (fundam.cpp)."""
    Mx, My = src[:, 0].astype(np.float64), src[:, 1].astype(np.float64)
    ww = h8[6] * Mx + h8[7] * My + 1.0
    ww = np.where(np.abs(ww) > DBL_EPSILON, 1.0 / ww, 0.0)
    xi = (h8[0] * Mx + h8[1] * My + h8[2]) * ww
    yi = (h8[3] * Mx + h8[4] * My + h8[5]) * ww
    r = np.empty(2 * Mx.shape[0])
    r[0::2] = xi - dst[:, 0].astype(np.float64)
    r[1::2] = yi - dst[:, 1].astype(np.float64)
    if not want_jac:
        return r, None
    J = np.zeros((2 * Mx.shape[0], 8))
    J[0::2, 0], J[0::2, 1], J[0::2, 2] = Mx * ww, My * ww, ww
    J[0::2, 6], J[0::2, 7] = -Mx * ww * xi, -My * ww * xi
    J[1::2, 3], J[1::2, 4], J[1::2, 5] = Mx * ww, My * ww, ww
    J[1::2, 6], J[1::2, 7] = -Mx * ww * yi, -My * ww * yi
    return r, J


def lm_refine(H, src, dst, max_iters=10):
    """LMSolverImpl::run (calib3d/src/levmarq.cpp, 4.x) on the 8 free entries of H."""
    eps = FLT_EPSILON
    x = (H.ravel() / H[2, 2])[:8].copy()
    lx = 8
    r, J = _refine_residual_jac(x, src, dst)
    S = float(r @ r)
    A = J.T @ J
    v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = _solve_eig(Ap, v)
        xd = x - d
        rd, _ = _refine_residual_jac(xd, src, dst, want_jac=False)
        Sd = float(rd @ rd)
        temp_d = 2.0 * v - A @ d
        dS = float(d @ temp_d)
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                Ai = _inv_eig(A)
                maxval = max(DBL_EPSILON, float(np.abs(np.diag(Ai)).max()))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = _refine_residual_jac(x, src, dst)
            A = J.T @ J
            v = J.T @ r
        it += 1
        proceed = it < max_iters and np.abs(d).max() >= eps and np.abs(r).max() >= eps
        if not proceed:
            break
    return np.append(x, 1.0).reshape(3, 3)


def _solve_eig(A, b):
    """cv::solve(A, b, DECOMP_EIG) for symmetric A: x = V diag(1/w) V^T b (tiny w -> 0)."""
    w, V = np.linalg.eigh(A)
    thr = 2 * DBL_EPSILON * np.abs(w).sum()      # SVBkSb threshold
    y = V.T @ b
    y = np.where(np.abs(w) > thr, y / np.where(w == 0, 1, w), 0.0)
    return V @ y


def _inv_eig(A):
    w, V = np.linalg.eigh(A)
    thr = 2 * DBL_EPSILON * np.abs(w).sum()
    wi = np.where(np.abs(w) > thr, 1.0 / np.where(w == 0, 1, w), 0.0)
    return (V * wi) @ V.T


def find_homography_ransac(src, dst, thr=3.0, max_iters=2000, confidence=0.995):
    """cv::findHomography(src, dst, RANSAC, thr): returns (H 3x3 f64, mask N u8) or (None, None)."""
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 2)
    n = src.shape[0]
    if n < 4:
        raise ValueError("-28: findHomography needs at least 4 point pairs")
    if n == 4:
        H = dlt_homography(src, dst)
        if H is None:
            return None, None
        return H, np.ones(4, np.uint8)
    res = ransac_loop(src, dst, thr, max_iters, confidence)
    if res is None:
        return None, None
    H, mask, _ = res
    s, d = src[mask], dst[mask]
    H2 = dlt_homography(s, d)
    if H2 is not None:
        H = H2
    H = lm_refine(H, s, d, 10)
    # cv2 4.13.0 re-derives the returned mask from the refined H
    mask2, _ = find_inliers(H, src, dst, thr)
    return H, mask2.astype(np.uint8)


def lmeds_loop(src, dst, max_iters=2000, confidence=0.995):
    """LMeDSPointSetRegistrator::run (calib3d/src/ptsetreg.cpp): the same sample stream as RANSAC, a fixed number of
    iterations (outlier ratio 0.45 -> 55 for confidence 0.995), the model with the smallest MEDIAN f32 reprojection
    error wins (std::nth_element at count / 2, strict <), inliers are the points within
    sigma = 2.5 * 1.4826 * (1 + 5 / (count - 4)) * sqrt(median).  Returns (H, mask, sigma, iterations) or None."""
    n = src.shape[0]
    rng = CvRNG()
    niters = max(ransac_update_num_iters(confidence, 0.45, 4, max_iters), 3)
    min_median, best_H = np.finfo(np.float64).max, None
    it = 0
    while it < niters:
        idx = get_subset(src, dst, rng, max_attempts=1000)     # getSubset's default; only RANSAC passes 10000
        if idx is None:
            if it == 0:
                return None
            break
        H = dlt_homography(src[idx], dst[idx])
        if H is not None:
            median = float(np.sort(reproj_err_f32(H, src, dst))[n // 2])
            if median < min_median:
                min_median, best_H = median, H
        it += 1
    if best_H is None:
        return None
    sigma = max(2.5 * 1.4826 * (1.0 + 5.0 / (n - 4)) * np.sqrt(min_median), 0.001)
    mask, good = find_inliers(best_H, src, dst, sigma)
    if good < 4:
        return None
    return best_H, mask, sigma, it


def find_homography_lmeds(src, dst, thr=3.0, max_iters=2000, confidence=0.995):
    """cv::findHomography(src, dst, LMEDS, thr): returns (H 3x3 f64, mask N u8) or (None, None).  The model is
    re-estimated on the sigma-inliers (DLT + LM, as for RANSAC); the RETURNED mask is the inlier set of the final H
    under `thr` (cv2 4.13.0, established against cv2 on 30 noisy problems where the three candidate rules differ)."""
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 2)
    n = src.shape[0]
    if n < 4:
        raise ValueError("-28: findHomography needs at least 4 point pairs")
    if n == 4:
        H = dlt_homography(src, dst)
        if H is None:
            return None, None
        return H, np.ones(4, np.uint8)
    res = lmeds_loop(src, dst, max_iters, confidence)
    if res is None:
        return None, None
    H, mask, _, _ = res
    s, d = src[mask], dst[mask]
    H2 = dlt_homography(s, d)
    if H2 is not None:
        H = H2
    H = lm_refine(H, s, d, 10)
    mask2, _ = find_inliers(H, src, dst, thr)
    return H, mask2.astype(np.uint8)


def find_homography_rho(src, dst, thr=3.0, max_iters=2000, confidence=0.995):
    """HomographyMethod::RHO (homographier/src/homographier/mod.rs:25-31, passed through at :243-250).

    PARITY UNPINNED for this method: OpenCV's rho.cpp (PROSAC sampling in the order of the pairs, SPRT verification,
    its own xorshift stream, Cholesky LM refinement, mask of the un-refined best model) is not available in the
    reference tree and is not restated.  The product serves a RHO request with the RANSAC estimator above — same
    contract (threshold, 2000 iterations, confidence 0.995, robust H + inlier mask) — and the tests check cv2's RHO
    goldens by tolerance only (tests/golden/rho_golden.npz: cv2's own RHO and RANSAC differ by up to 2.3e-2 there)."""
    return find_homography_ransac(src, dst, thr, max_iters, confidence)
