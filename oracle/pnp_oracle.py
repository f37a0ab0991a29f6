"""Oracle for stage 3b (PnP-RANSAC pose) — TEST INFRASTRUCTURE ONLY.

Restates `cv::solvePnPRansac(obj, img, K, dist=0, rvec, tvec, false, iters, thr, conf, inliers,
SOLVEPNP_EPNP)` as the reference calls it (homographier/src/homographier/mod.rs:320-369; zero
distortion is forced at :344, EPNP is the default method at :359).  The arithmetic lives in
OpenCV's calib3d (`solvepnp.cpp`, `epnp.cpp`, `ptsetreg.cpp`) and core (`lapack.cpp` JacobiSVD,
`calibration_base.cpp` Rodrigues / projectPoints); not vendored in the reference (opencv crate
0.88.8 over the system libopencv 4.x).  Pinned against cv2 4.13.0 outputs in
tests/golden/pnp_golden.npz (tests/golden/make_golden.py).

What the restatement follows (all established differentially against cv2 4.13.0):
  * f64 inputs are converted to f32 before the RANSAC loop (solvepnp.cpp: `convertTo(CV_32F)`),
    the final solve on the inliers runs on those f32 values widened back to f64;
  * minimal sample = 5 points (EPnP kernel), same MWC sample stream as findHomography, no
    checkSubset; a model is accepted when good > max(best, 4);
  * EPnP (Lepetit/Moreno-Noguer/Fua) with OpenCV's conventions: control points from the one-sided
    Jacobi SVD of PW0^T PW0 (the SIGNS of its singular vectors change the answer under noise, so
    the Jacobi sweep order is reproduced), betas from approximations 1..3 + 5 Gauss-Newton steps,
    Procrustes R,t, candidate with the smallest mean reprojection error wins;
  * error = squared f32 distance between the f32 image point and the f64 projection rounded to
    f32; inlier iff err <= (float)(thr^2).
For 5-point samples M^T M has a 2-dimensional null space whose basis OpenCV derives from rounding
noise, so per-hypothesis poses are only reproducible to ~1e-6; the final pose (>= 6 inliers) is
reproducible to ~1e-9 whenever the winning inlier set agrees.
"""
from __future__ import annotations

import math

import numpy as np

from .ransac_oracle import CvRNG, ransac_update_num_iters

DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)


# ----------------------------------------------------------------------------- cv::SVD (Jacobi)
def jacobi_svd(A):
    """cv::SVD::compute for an m x n f64 matrix with m >= n < 25 (core/src/lapack.cpp
    JacobiSVDImpl_: one-sided Hestenes rotations on the rows of A^T, cyclic (i, j) order,
    eps = 10 DBL_EPSILON, <= max(m, 30) sweeps, then a selection sort by decreasing singular
    value).  Returns (w[n], U[m x n], Vt[n x n]) with OpenCV's signs."""
    A = np.array(A, dtype=np.float64)
    m, n = A.shape
    assert m >= n
    At = A.T.copy()                      # n rows of length m
    Vt = np.eye(n)
    W = (At * At).sum(1)
    eps = DBL_EPSILON * 10
    for _ in range(max(m, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps * math.sqrt(a * b):
                    continue
                p *= 2.0
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = math.sqrt(delta / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                t0 = c * At[i] + s * At[j]
                t1 = -s * At[i] + c * At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = float(t0 @ t0), float(t1 @ t1)
                changed = True
                v0 = c * Vt[i] + s * Vt[j]
                v1 = -s * Vt[i] + c * Vt[j]
                Vt[i], Vt[j] = v0, v1
        if not changed:
            break
    W = np.sqrt((At * At).sum(1))
    for i in range(n - 1):
        j = i
        for k in range(i + 1, n):
            if W[j] < W[k]:
                j = k
        if i != j:
            W[[i, j]] = W[[j, i]]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    U = np.zeros((m, n))
    for i in range(n):
        sd = W[i]
        U[:, i] = At[i] * (1.0 / sd if sd > DBL_MIN else 0.0)   # (zero singular values: OpenCV fills a random vector)
    return W, U, Vt


def _svd_solve(A, b):
    """cv::solve(A, b, x, DECOMP_SVD): least squares through the SVD (SVBkSb threshold)."""
    w, U, Vt = jacobi_svd(A)
    thr = DBL_EPSILON * 2 * w.sum()
    y = U.T @ b
    y = np.where(w > thr, y / np.where(w > thr, w, 1.0), 0.0)
    return Vt.T @ y


def _svd_inv3(A):
    w, U, Vt = jacobi_svd(A)
    thr = DBL_EPSILON * 2 * w.sum()
    wi = np.where(w > thr, 1.0 / np.where(w > thr, w, 1.0), 0.0)
    return (Vt.T * wi) @ U.T


def _qr_solve(A, b):
    """epnp::qr_solve — Householder least squares of the 6 x 4 Gauss-Newton system."""
    return np.linalg.lstsq(A, b, rcond=None)[0]


# ----------------------------------------------------------------------------- cv::Rodrigues
def rodrigues_to_matrix(r):
    r = np.asarray(r, dtype=np.float64).ravel()
    theta = math.sqrt(float(r @ r))
    if theta < DBL_EPSILON:
        return np.eye(3)
    c, s = math.cos(theta), math.sin(theta)
    c1 = 1.0 - c
    k = r / theta
    rrt = np.outer(k, k)
    rx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


def rodrigues_to_vector(R):
    w, U, Vt = jacobi_svd(np.asarray(R, dtype=np.float64))
    R = U @ Vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = math.sqrt(float(r @ r) * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5
    c = 1.0 if c > 1.0 else (-1.0 if c < -1.0 else c)
    theta = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        r[0] = math.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        r[1] = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        r[2] = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(r[0]) < abs(r[1]) and abs(r[0]) < abs(r[2]) and ((R[1, 2] > 0) != (r[1] * r[2] > 0)):
            r[2] = -r[2]
        return r * (theta / math.sqrt(float(r @ r)))
    return r * (theta / (2 * s))


# ----------------------------------------------------------------------------- EPnP
_PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def epnp(obj, img, K, f32_normalised=False):
    """epnp::compute_pose (calib3d/src/epnp.cpp).  obj n x 3, img n x 2 (pixels), K 3 x 3; all f64.
    f32_normalised: the image points reached solvePnP as CV_32F, so undistortPoints returned f32
    normalised coordinates (every call made from inside solvePnPRansac's loop).
    Returns (R 3x3, t 3, mean reprojection error)."""
    pws = np.asarray(obj, dtype=np.float64).reshape(-1, 3)
    img = np.asarray(img, dtype=np.float64).reshape(-1, 2)
    n = pws.shape[0]
    fu, fv, uc, vc = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    # solvePnPGeneric: undistortPoints (zero distortion -> (u - cx)/fx), then epnp::init_points maps back
    xn, yn = (img[:, 0] - uc) * (1.0 / fu), (img[:, 1] - vc) * (1.0 / fv)     # cvUndistortPoints: (x - cx) * ifx
    if f32_normalised:
        xn, yn = xn.astype(np.float32).astype(np.float64), yn.astype(np.float32).astype(np.float64)
    us = np.stack([xn * fu + uc, yn * fv + vc], 1)

    # choose_control_points
    cws = np.zeros((4, 3))
    cws[0] = pws.sum(0) / n
    PW0 = pws - cws[0]
    dc, U, _ = jacobi_svd(PW0.T @ PW0)
    for i in range(1, 4):
        cws[i] = cws[0] + math.sqrt(dc[i - 1] / n) * U[:, i - 1]
    # compute_barycentric_coordinates
    CC = (cws[1:] - cws[0]).T
    CCi = _svd_inv3(CC)
    al = np.zeros((n, 4))
    al[:, 1:] = (pws - cws[0]) @ CCi.T
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    # fill_M, M^T M, its 4 smallest left singular vectors
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j] * fu
        M[0::2, 3 * j + 2] = al[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = al[:, j] * fv
        M[1::2, 3 * j + 2] = al[:, j] * (vc - us[:, 1])
    _, U12, _ = jacobi_svd(M.T @ M)
    v = [U12[:, 11], U12[:, 10], U12[:, 9], U12[:, 8]]
    # compute_L_6x10, compute_rho
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        for p, (a, b) in enumerate(_PAIRS):
            dv[i, p] = v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3]
    L = np.zeros((6, 10))
    for p in range(6):
        d0, d1, d2, d3 = dv[0, p], dv[1, p], dv[2, p], dv[3, p]
        L[p] = [d0 @ d0, 2 * (d0 @ d1), d1 @ d1, 2 * (d0 @ d2), 2 * (d1 @ d2), d2 @ d2,
                2 * (d0 @ d3), 2 * (d1 @ d3), 2 * (d2 @ d3), d3 @ d3]
    rho = np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in _PAIRS])

    def approx_1():
        b4 = _svd_solve(L[:, [0, 1, 3, 6]], rho)
        if b4[0] < 0:
            b0 = math.sqrt(-b4[0])
            return np.array([b0, -b4[1] / b0, -b4[2] / b0, -b4[3] / b0])
        b0 = math.sqrt(b4[0])
        return np.array([b0, b4[1] / b0, b4[2] / b0, b4[3] / b0])

    def _two(b):
        be = np.zeros(4)
        if b[0] < 0:
            be[0] = math.sqrt(-b[0])
            be[1] = math.sqrt(-b[2]) if b[2] < 0 else 0.0
        else:
            be[0] = math.sqrt(b[0])
            be[1] = math.sqrt(b[2]) if b[2] > 0 else 0.0
        if b[1] < 0:
            be[0] = -be[0]
        return be

    def approx_2():
        return _two(_svd_solve(L[:, [0, 1, 2]], rho))

    def approx_3():
        b5 = _svd_solve(L[:, [0, 1, 2, 3, 4]], rho)
        be = _two(b5)
        be[2] = b5[3] / be[0]
        return be

    def gauss_newton(be):
        be = be.copy()
        for _ in range(5):
            A = np.stack([2 * L[:, 0] * be[0] + L[:, 1] * be[1] + L[:, 3] * be[2] + L[:, 6] * be[3],
                          L[:, 1] * be[0] + 2 * L[:, 2] * be[1] + L[:, 4] * be[2] + L[:, 7] * be[3],
                          L[:, 3] * be[0] + L[:, 4] * be[1] + 2 * L[:, 5] * be[2] + L[:, 8] * be[3],
                          L[:, 6] * be[0] + L[:, 7] * be[1] + L[:, 8] * be[2] + 2 * L[:, 9] * be[3]], 1)
            b = rho - (L[:, 0] * be[0] * be[0] + L[:, 1] * be[0] * be[1] + L[:, 2] * be[1] * be[1] + L[:, 3] * be[0] * be[2]
                       + L[:, 4] * be[1] * be[2] + L[:, 5] * be[2] * be[2] + L[:, 6] * be[0] * be[3] + L[:, 7] * be[1] * be[3]
                       + L[:, 8] * be[2] * be[3] + L[:, 9] * be[3] * be[3])
            be = be + _qr_solve(A, b)
        return be

    def compute_R_and_t(be):
        ccs = np.zeros((4, 3))
        for i in range(4):
            for j in range(4):
                ccs[j] += be[i] * v[i][3 * j:3 * j + 3]
        pcs = al @ ccs
        if pcs[0, 2] < 0.0:                      # solve_for_sign
            ccs, pcs = -ccs, -pcs
        pc0, pw0 = pcs.sum(0) / n, pws.sum(0) / n
        ABt = (pcs - pc0).T @ (pws - pw0)
        _, Ua, Vta = jacobi_svd(ABt)
        R = Ua @ Vta
        if np.linalg.det(R) < 0:
            R[2] = -R[2]
        t = pc0 - R @ pw0
        P = pws @ R.T + t
        ue, ve = uc + fu * P[:, 0] / P[:, 2], vc + fv * P[:, 1] / P[:, 2]
        err = float(np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2).sum() / n)
        return R, t, err

    best = None
    for f in (approx_1, approx_2, approx_3):
        with np.errstate(all="ignore"):
            cand = compute_R_and_t(gauss_newton(f()))
        if best is None or cand[2] < best[2]:
            best = cand
    return best


def solve_pnp_epnp(obj, img, K, f32_normalised=False):
    """cv::solvePnP(..., SOLVEPNP_EPNP): (rvec, tvec) or None when the pose is not finite."""
    R, t, _ = epnp(obj, img, K, f32_normalised)
    if not (np.isfinite(R).all() and np.isfinite(t).all()):
        return None
    return rodrigues_to_vector(R), t


# ----------------------------------------------------------------------------- RANSAC
def project_points(obj, rvec, tvec, K):
    """cv::projectPoints with zero distortion, f64 arithmetic."""
    R = rodrigues_to_matrix(rvec)
    P = np.asarray(obj, dtype=np.float64) @ R.T + np.asarray(tvec, dtype=np.float64)
    z = 1.0 / P[:, 2]
    x, y = P[:, 0] * z, P[:, 1] * z
    return np.stack([x * K[0, 0] + K[0, 2], y * K[1, 1] + K[1, 2]], 1)


def reproj_err_f32(obj32, img32, rvec, tvec, K):
    """PnPRansacCallback::computeError: f32 projected points, f32 squared distance (no FMA)."""
    with np.errstate(all="ignore"):
        pp = project_points(obj32, rvec, tvec, K).astype(np.float32)
        d = img32 - pp
        return d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]


def sample_stream(n, count, model_points=5):
    """the first `count` minimal index sets of the registrator's fixed-seed stream (no checkSubset)."""
    rng = CvRNG()
    out = []
    for _ in range(count):
        idx = []
        for _i in range(model_points):
            v = rng.uniform(0, n)
            while v in idx:
                v = rng.uniform(0, n)
            idx.append(v)
        out.append(idx)
    return np.array(out, dtype=np.int32).reshape(-1, model_points)


def solve_pnp_ransac(obj, img, K, iters=100, thr=8.0, confidence=0.99):
    """cv::solvePnPRansac(obj, img, K, zeros, iters, thr, confidence, flags=EPNP).
    Returns (found, rvec[3], tvec[3], inlier indices int32)."""
    obj32 = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3).astype(np.float32)
    img32 = np.ascontiguousarray(img, dtype=np.float64).reshape(-1, 2).astype(np.float32)
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    n = obj32.shape[0]
    if n < 4:
        raise ValueError("-215: solvePnPRansac needs at least 4 correspondences")
    if n == 4:
        raise NotImplementedError("npoints == 4 switches the kernel to P3P (not on the reference's default path)")
    mp = 5
    objd, imgd = obj32.astype(np.float64), img32.astype(np.float64)
    if n == mp:
        sol = solve_pnp_epnp(objd, imgd, K, True)
        if sol is None:
            return False, np.zeros(3), np.zeros(3), np.zeros(0, np.int32)
        return True, sol[0], sol[1], np.arange(n, dtype=np.int32)
    thr2 = np.float32(float(thr) * float(thr))
    rng = CvRNG()
    niters = iters
    best = None
    best_count = 0
    it = 0
    while it < niters:
        idx = []
        for _ in range(mp):
            v = rng.uniform(0, n)
            while v in idx:
                v = rng.uniform(0, n)
            idx.append(v)
        sol = solve_pnp_epnp(objd[idx], imgd[idx], K, True)
        if sol is not None:
            err = reproj_err_f32(obj32, img32, sol[0], sol[1], K)
            mask = err <= thr2
            good = int(mask.sum())
            if good > max(best_count, mp - 1):
                best, best_count = (sol, mask), good
                niters = ransac_update_num_iters(confidence, (n - good) / n, mp, niters)
        it += 1
    if best is None:
        return False, np.zeros(3), np.zeros(3), np.zeros(0, np.int32)
    (rv, tv), mask = best
    sol = solve_pnp_epnp(objd[mask], imgd[mask], K)
    if sol is None:
        return False, rv, tv, np.zeros(0, np.int32)
    return True, sol[0], sol[1], np.nonzero(mask)[0].astype(np.int32)


def refine_pose_lm(obj, img, K, rvec, tvec, max_iters=100):
    """SOLVEPNP_ITERATIVE's answer: the minimum of the squared reprojection error (calib3d solvePnP ->
    Levenberg-Marquardt over (rvec, tvec); cv2 4.13.0 agrees with the minimiser to ~1e-8).  Gauss-Newton with a
    local rotation update R <- exp([dw]x) R; stops when the step is below 1e-14."""
    obj = np.asarray(obj, np.float64).reshape(-1, 3)
    img = np.asarray(img, np.float64).reshape(-1, 2)
    K = np.asarray(K, np.float64).reshape(3, 3)
    fu, fv, uc, vc = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    R, t = rodrigues_to_matrix(np.asarray(rvec, np.float64)), np.asarray(tvec, np.float64).copy()
    for _ in range(max_iters):
        a = obj @ R.T
        x, y, z = (a + t).T
        ru, rv = fu * x / z + uc - img[:, 0], fv * y / z + vc - img[:, 1]
        ux, uz, vy, vz = fu / z, -fu * x / z ** 2, fv / z, -fv * y / z ** 2
        zero = np.zeros_like(z)
        ju = np.stack([uz * a[:, 1], ux * a[:, 2] - uz * a[:, 0], -ux * a[:, 1], ux, zero, uz], 1)
        jv = np.stack([-vy * a[:, 2] + vz * a[:, 1], -vz * a[:, 0], vy * a[:, 0], zero, vy, vz], 1)
        J = np.concatenate([ju, jv])
        d = np.linalg.solve(J.T @ J, -J.T @ np.concatenate([ru, rv]))
        R = rodrigues_to_matrix(d[:3]) @ R
        t = t + d[3:]
        if np.abs(d[:3]).max() < 1e-14 and np.abs(d[3:]).max() < 1e-14 * max(1.0, np.abs(t).max()):
            break
    return rodrigues_to_vector(R), t


def solve_pnp_ransac_iterative(obj, img, K, iters=100, thr=8.0, confidence=0.99):
    """cv::solvePnPRansac(..., flags=SOLVEPNP_ITERATIVE): the RANSAC stage is the EPnP one (5-point samples); when
    there are more points than the minimal sample the final pose over the inliers is the LM minimum."""
    found, rv, tv, inl = solve_pnp_ransac(obj, img, K, iters, thr, confidence)
    n = np.asarray(obj).reshape(-1, 3).shape[0]
    if not found or n <= 5:
        return found, rv, tv, inl
    obj32 = np.asarray(obj, np.float64).reshape(-1, 3).astype(np.float32).astype(np.float64)
    img32 = np.asarray(img, np.float64).reshape(-1, 2).astype(np.float32).astype(np.float64)
    rv, tv = refine_pose_lm(obj32[inl], img32[inl], K, rv, tv)
    return True, rv, tv, inl


# ----------------------------------------------------------------------------- P3P (SOLVEPNP_P3P)
# OpenCV's p3p.cpp (Gao et al.) solves a quartic for the ratio of two depths and aligns the three
# camera-frame points with Horn's quaternion method.  The restatement below uses the same unknowns
# (Fischler-Bolles form s1 = u s0, s2 = v s0; the quartic's coefficients were derived symbolically,
# tools/derive_p3p.py), OpenCV's Ferrari / Cardano root finder (solve_deg4 / solve_deg3), a triad
# alignment (exact for the congruent triangles P3P produces) and OpenCV's rule for the 4th point: the
# solution with the smallest squared reprojection error in normalised coordinates wins.  Every correct
# P3P returns the same geometric solutions, so poses agree with cv2 to rounding (~1e-9), not bitwise.
def _solve_deg2(a, b, c):
    delta = b * b - 4 * a * c
    if delta < 0:
        return []
    inv_2a = 0.5 / a
    if delta == 0:
        return [-b * inv_2a]
    s = math.sqrt(delta)
    return [(-b + s) * inv_2a, (-b - s) * inv_2a]


def _solve_deg3(a, b, c, d):
    if a == 0:
        if b == 0:
            return [] if c == 0 else [-d / c]
        return _solve_deg2(b, c, d)
    inv_a = 1.0 / a
    b_a, c_a, d_a = inv_a * b, inv_a * c, inv_a * d
    b_a2 = b_a * b_a
    Q = (3 * c_a - b_a2) / 9
    R = (9 * b_a * c_a - 27 * d_a - 2 * b_a * b_a2) / 54
    Q3 = Q * Q * Q
    D = Q3 + R * R
    b_a_3 = (1.0 / 3.0) * b_a
    if Q == 0:
        if R == 0:
            return [-b_a_3] * 3
        return [math.copysign(abs(2 * R) ** (1 / 3.0), R) - b_a_3]
    if D <= 0:
        theta = math.acos(max(-1.0, min(1.0, R / math.sqrt(-Q3))))
        sq = math.sqrt(-Q)
        return [2 * sq * math.cos(theta / 3.0) - b_a_3, 2 * sq * math.cos((theta + 2 * math.pi) / 3.0) - b_a_3,
                2 * sq * math.cos((theta + 4 * math.pi) / 3.0) - b_a_3]
    AD, BD = 0.0, 0.0
    R_abs = abs(R)
    if R_abs > DBL_EPSILON:
        AD = (R_abs + math.sqrt(D)) ** (1 / 3.0)
        AD = AD if R >= 0 else -AD
        BD = -Q / AD
    return [AD + BD - b_a_3]


def _solve_deg4(a, b, c, d, e):
    if a == 0:
        return _solve_deg3(b, c, d, e)
    inv_a = 1.0 / a
    b, c, d, e = b * inv_a, c * inv_a, d * inv_a, e * inv_a
    b2, bc, b3 = b * b, b * c, b * b * b
    r = _solve_deg3(1, -c, d * b - 4 * e, 4 * c * e - d * d - b2 * e)
    if not r:
        return []
    r0 = r[0]
    R2 = 0.25 * b2 - c + r0
    if R2 < 0:
        return []
    R = math.sqrt(R2)
    if R < 10e-12:
        temp = r0 * r0 - 4 * e
        if temp < 0:
            D2 = E2 = -1.0
        else:
            st = math.sqrt(temp)
            D2 = 0.75 * b2 - 2 * c + 2 * st
            E2 = D2 - 4 * st
    else:
        uu = 0.75 * b2 - 2 * c - R2
        vv = 0.25 * (1.0 / R) * (4 * bc - 8 * d - b3)
        D2, E2 = uu + vv, uu - vv
    b_4, R_2 = 0.25 * b, 0.5 * R
    out = []
    if D2 >= 0:
        D = math.sqrt(D2)
        x0 = R_2 + 0.5 * D - b_4
        out += [x0, x0 - D]
    if E2 >= 0:
        E = math.sqrt(E2)
        x2 = -R_2 + 0.5 * E - b_4
        out += [x2, x2 - E]
    return out


def _triad(p0, p1, p2):
    e1 = p1 - p0
    e1 = e1 / math.sqrt(float(e1 @ e1))
    e3 = np.cross(e1, p2 - p0)
    e3 = e3 / math.sqrt(float(e3 @ e3))
    return np.stack([e1, np.cross(e3, e1), e3], 1)          # columns


def p3p_solutions(obj3, xy3):
    """all (R, t) with R X_i + t on the bearing of normalised image point i, i = 0..2"""
    P = np.asarray(obj3, np.float64)
    f = np.c_[np.asarray(xy3, np.float64), np.ones(3)]
    f = f / np.sqrt((f * f).sum(1))[:, None]
    a, b, c = ((P[0] - P[1]) ** 2).sum(), ((P[0] - P[2]) ** 2).sum(), ((P[1] - P[2]) ** 2).sum()
    p, q, r = float(f[0] @ f[1]), float(f[0] @ f[2]), float(f[1] @ f[2])
    if a == 0 or b == 0 or c == 0:
        return []
    k4 = -a * a + 4 * a * b * r * r - 2 * a * b + 2 * a * c - b * b + 2 * b * c - c * c
    k3 = -4 * (-a * a * q * r + 2 * a * b * p * r * r - a * b * p + a * b * q * r + a * c * p + a * c * q * r - b * b * p + 2 * b * c * p - c * c * p)
    k2 = -2 * (2 * a * a * q * q + 2 * a * a * r * r - a * a - 4 * a * b * p * q * r - 2 * a * b * r * r - 4 * a * c * p * q * r - 2 * a * c * q * q
               + 2 * b * b * p * p + b * b - 4 * b * c * p * p - 2 * b * c + 2 * c * c * p * p + c * c)
    k1 = -4 * (-a * a * q * r + a * b * p + a * b * q * r + 2 * a * c * p * q * q - a * c * p + a * c * q * r - b * b * p + 2 * b * c * p - c * c * p)
    k0 = -a * a + 2 * a * b + 4 * a * c * q * q - 2 * a * c - b * b + 2 * b * c - c * c
    sols = []
    for u in _solve_deg4(k4, k3, k2, k1, k0):
        if not (u > 0):
            continue
        den = 2 * a * (q - r * u)
        if den == 0:
            continue
        v = (-a * u * u + a + 2 * b * p * u - b * u * u - b - 2 * c * p * u + c * u * u + c) / den
        w = 1 + u * u - 2 * u * p
        if not (v > 0 and w > 0):
            continue
        s0 = math.sqrt(a / w)
        C = np.stack([s0 * f[0], u * s0 * f[1], v * s0 * f[2]])
        R = _triad(C[0], C[1], C[2]) @ _triad(P[0], P[1], P[2]).T
        t = C[0] - R @ P[0]
        if np.isfinite(R).all() and np.isfinite(t).all():
            sols.append((R, t))
    return sols


def solve_pnp_p3p(obj4, img4, K, f32_normalised=False):
    """cv::solvePnP(4 points, SOLVEPNP_P3P): P3P on the first three points, the fourth picks the solution.
    Returns (rvec, tvec) or None."""
    obj4 = np.asarray(obj4, np.float64).reshape(4, 3)
    img4 = np.asarray(img4, np.float64).reshape(4, 2)
    fu, fv, uc, vc = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    xn, yn = (img4[:, 0] - uc) * (1.0 / fu), (img4[:, 1] - vc) * (1.0 / fv)
    if f32_normalised:
        xn, yn = xn.astype(np.float32).astype(np.float64), yn.astype(np.float32).astype(np.float64)
    xy = np.stack([xn, yn], 1)
    best = None
    for R, t in p3p_solutions(obj4[:3], xy[:3]):
        X = R @ obj4[3] + t
        e = (X[0] / X[2] - xy[3, 0]) ** 2 + (X[1] / X[2] - xy[3, 1]) ** 2
        if best is None or e < best[0]:
            best = (e, R, t)
    if best is None:
        return None
    return rodrigues_to_vector(best[1]), best[2]


def solve_pnp_ransac_p3p(obj, img, K, iters=100, thr=8.0, confidence=0.99):
    """cv::solvePnPRansac(..., flags = SOLVEPNP_P3P): 4-point samples, P3P kernel, final EPnP on the inliers
    (solvepnp.cpp replaces P3P by EPNP for the final solve)."""
    obj32 = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3).astype(np.float32)
    img32 = np.ascontiguousarray(img, dtype=np.float64).reshape(-1, 2).astype(np.float32)
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    n = obj32.shape[0]
    if n < 4:
        raise ValueError("-215: solvePnPRansac needs at least 4 correspondences")
    objd, imgd = obj32.astype(np.float64), img32.astype(np.float64)
    none = (False, np.zeros(3), np.zeros(3), np.zeros(0, np.int32))
    if n == 4:
        sol = solve_pnp_p3p(objd, imgd, K, True)
        return none if sol is None else (True, sol[0], sol[1], np.arange(4, dtype=np.int32))
    thr2 = np.float32(float(thr) * float(thr))
    rng = CvRNG()
    niters, best, best_count, it = iters, None, 0, 0
    while it < niters:
        idx = []
        for _ in range(4):
            v = rng.uniform(0, n)
            while v in idx:
                v = rng.uniform(0, n)
            idx.append(v)
        sol = solve_pnp_p3p(objd[idx], imgd[idx], K, True)
        if sol is not None:
            mask = reproj_err_f32(obj32, img32, sol[0], sol[1], K) <= thr2
            good = int(mask.sum())
            if good > max(best_count, 3):
                best, best_count = (sol, mask), good
                niters = ransac_update_num_iters(confidence, (n - good) / n, 4, niters)
        it += 1
    if best is None:
        return none
    mask = best[1]
    sol = solve_pnp_epnp(objd[mask], imgd[mask], K)
    if sol is None:
        return False, best[0][0], best[0][1], np.zeros(0, np.int32)
    return True, sol[0], sol[1], np.nonzero(mask)[0].astype(np.int32)
