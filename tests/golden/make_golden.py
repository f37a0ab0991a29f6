"""Generates the golden vectors under tests/golden/ by running the reference's OpenCV entry
points (through cv2, the same C++ library the `opencv` crate binds) with the reference's exact
arguments.  Run in the build container:  python tests/golden/make_golden.py [match|ransac|lmeds|rho|akaze|config1|pnp|pnp_iter|warp|l2|all]
The OpenCV version is recorded in every file (parity is defined against that version)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def knn_cv(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)      # lib.rs:101-103
    idx = np.array([[a.trainIdx for a in r] + [-1] * (2 - len(r)) for r in m], dtype=np.int32).reshape(-1, 2)
    dist = np.array([[a.distance for a in r] + [-1] * (2 - len(r)) for r in m], dtype=np.int32).reshape(-1, 2)
    return idx, dist


def cross_cv(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)             # lib.rs:121-123
    return np.array([(a.queryIdx, a.trainIdx, a.distance) for a in m], dtype=np.int32).reshape(-1, 3)


def make_match():
    rng = np.random.default_rng(20240321)
    cases = {}
    # uniform random 61-byte rows (last byte 6 bits) — ragged sizes, not multiples of any tile
    q = rng.integers(0, 256, (333, 61), dtype=np.uint8); q[:, 60] &= 0x3F
    t = rng.integers(0, 256, (1777, 61), dtype=np.uint8); t[:, 60] &= 0x3F
    cases["rand"] = (q, t)
    # tie-heavy: few distinct bits -> many equal distances (tie-break = lowest train index)
    q = np.zeros((257, 61), np.uint8); t = np.zeros((901, 61), np.uint8)
    q[:, :2] = rng.integers(0, 256, (257, 2)); t[:, :2] = rng.integers(0, 256, (901, 2))
    cases["ties"] = (q, t)
    # duplicates planted: exact zero-distance matches and duplicate train rows
    t = rng.integers(0, 256, (640, 61), dtype=np.uint8); t[:, 60] &= 0x3F
    q = t[rng.integers(0, 640, 200)].copy()
    t[100:110] = t[5]                                                  # duplicate rows
    cases["dups"] = (q, t)
    # minimum train size for the ratio test, and a single query
    cases["tiny"] = (rng.integers(0, 256, (1, 61), dtype=np.uint8), rng.integers(0, 256, (2, 61), dtype=np.uint8))
    out = {"opencv_version": np.array(cv2.__version__)}
    for name, (q, t) in cases.items():
        idx, dist = knn_cv(q, t)
        out[f"{name}_q"], out[f"{name}_t"] = q, t
        out[f"{name}_idx"], out[f"{name}_dist"] = idx, dist
        out[f"{name}_cross"] = cross_cv(q, t)
    np.savez_compressed(os.path.join(HERE, "match_golden.npz"), **out)
    print("match_golden.npz:", {k: v.shape for k, v in out.items()})


H_TRUE = np.array([[0.98, -0.12, 60], [0.10, 1.03, -40], [1e-5, -2e-5, 1]])


def ransac_case(n, out_frac, sigma, seed, H=H_TRUE):
    r = np.random.default_rng(seed)
    src = r.uniform(0, 1024, (n, 2)).astype(np.float32)
    p = np.c_[src, np.ones(n)] @ H.T
    dst = (p[:, :2] / p[:, 2:]) + r.normal(0, sigma, (n, 2))
    k = int(n * out_frac)
    dst[:k] = r.uniform(0, 1024, (k, 2))
    return src, dst.astype(np.float32)


RANSAC_CASES = [(50, 0.2, 0.5), (200, 0.4, 0.5), (1000, 0.3, 1.0), (2000, 0.4, 0.5), (500, 0.6, 0.3),
                (100, 0.0, 0.1), (300, 0.5, 1.5), (1500, 0.1, 0.8), (64, 0.7, 0.5), (800, 0.45, 0.7),
                (5, 0.0, 0.0), (4, 0.0, 0.0)]


def make_ransac():
    """cv2.findHomography(src, dst, RANSAC, thr) — mod.rs:243-250 (maxIters 2000, conf 0.995)."""
    out = {"opencv_version": np.array(cv2.__version__), "n_cases": np.array(len(RANSAC_CASES) + 1)}
    cases = [ransac_case(n, of, sg, seed) + (3.0,) for seed, (n, of, sg) in enumerate(RANSAC_CASES)]
    # the reference's own test homography_success (mod.rs:436-472): 10x10 grid onto itself, thr 1
    grid = np.array([(i, j) for i in range(1, 11) for j in range(1, 11)], dtype=np.float32)
    cases.append((grid, grid.copy(), 1.0))
    for i, (src, dst, thr) in enumerate(cases):
        H, mask = cv2.findHomography(src, dst, cv2.RANSAC, thr)
        out[f"c{i}_src"], out[f"c{i}_dst"], out[f"c{i}_thr"] = src, dst, np.array(thr)
        out[f"c{i}_H"], out[f"c{i}_mask"] = H, mask.ravel().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "ransac_golden.npz"), **out)
    print("ransac_golden.npz:", len(cases), "cases")


def make_pnp_iter():
    """cv2.solvePnPRansac(..., flags=SOLVEPNP_ITERATIVE) on the cases of pnp_golden.npz (same inputs, regenerated
    by pnp_case): the RANSAC stage is the EPnP one, the final pose is the Levenberg-Marquardt minimum of the
    reprojection error over the inliers."""
    out = {"opencv_version": np.array(cv2.__version__), "n_cases": np.array(len(PNP_CASES))}
    for i, (n, of, noise, iters, thr, conf, relief) in enumerate(PNP_CASES):
        obj, img = pnp_case(i, n, of, noise, relief)
        ok, r, t, inl = cv2.solvePnPRansac(obj, img, PNP_K, np.zeros((4, 1)), None, None, False, iters, thr, conf, None,
                                           cv2.SOLVEPNP_ITERATIVE)
        out[f"c{i}_checksum"] = np.array([obj.sum(), img.sum()])
        out[f"c{i}_found"] = np.array(bool(ok))
        out[f"c{i}_rvec"], out[f"c{i}_tvec"] = r.ravel(), t.ravel()
        out[f"c{i}_inliers"] = np.zeros(0, np.int32) if inl is None else inl.ravel().astype(np.int32)
        print(f"pnp iterative case {i}: n={n} found={ok} inliers={len(out[f'c{i}_inliers'])}")
    np.savez_compressed(os.path.join(HERE, "pnp_iter_golden.npz"), **out)


def make_rho():
    """cv2.findHomography(src, dst, RHO, thr) — HomographyMethod::RHO, mod.rs:25-31.  RHO is PROSAC: it assumes the pairs
    are ordered by quality, so the outliers ransac_case() puts FIRST make it fail; the cases are shuffled (what a
    matcher's query-ordered output looks like)."""
    out = {"opencv_version": np.array(cv2.__version__), "n_cases": np.array(len(RANSAC_CASES))}
    for i, (n, of, sg) in enumerate(RANSAC_CASES):
        src, dst = ransac_case(n, of, sg, 200 + i)
        p = np.random.default_rng(i).permutation(n)
        src, dst = src[p], dst[p]
        H, mask = cv2.findHomography(src, dst, cv2.RHO, 3.0)
        out[f"c{i}_src"], out[f"c{i}_dst"], out[f"c{i}_thr"] = src, dst, np.array(3.0)
        out[f"c{i}_H"], out[f"c{i}_mask"] = H, mask.ravel().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "rho_golden.npz"), **out)
    print("rho_golden.npz:", len(RANSAC_CASES), "cases")


LMEDS_CASES = [(50, 0.2, 0.5), (200, 0.4, 0.5), (1000, 0.3, 1.0), (2000, 0.4, 2.0), (100, 0.0, 0.1), (300, 0.45, 1.5),
               (1500, 0.1, 3.0), (777, 0.25, 4.0), (5, 0.0, 0.0), (4, 0.0, 0.0)]


def make_lmeds():
    """cv2.findHomography(src, dst, LMEDS, thr) — HomographyMethod::LMEDS, mod.rs:25-31,243-250.  Noise levels up to
    4 px make sigma, the sigma-inlier set and the returned (thr-based) mask differ, so the mask rule is pinned."""
    out = {"opencv_version": np.array(cv2.__version__), "n_cases": np.array(len(LMEDS_CASES) + 1)}
    cases = [ransac_case(n, of, sg, 100 + seed) + (3.0,) for seed, (n, of, sg) in enumerate(LMEDS_CASES)]
    grid = np.array([(i, j) for i in range(1, 11) for j in range(1, 11)], dtype=np.float32)
    cases.append((grid, grid.copy(), 1.0))
    for i, (src, dst, thr) in enumerate(cases):
        H, mask = cv2.findHomography(src, dst, cv2.LMEDS, thr)
        out[f"c{i}_src"], out[f"c{i}_dst"], out[f"c{i}_thr"] = src, dst, np.array(thr)
        out[f"c{i}_H"], out[f"c{i}_mask"] = H, mask.ravel().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "lmeds_golden.npz"), **out)
    print("lmeds_golden.npz:", len(cases), "cases")


def synth(n, seed, m=None):
    """SURVEY 8d synthetic image: multi-scale smoothed noise, min-max normalised to u8 (h=n, w=m)."""
    m = m or n
    rng = np.random.default_rng(seed)
    acc = np.zeros((n, m), np.float64)
    for s in (1, 2, 4, 8, 16, 32):
        acc += cv2.resize(rng.standard_normal((n // s + 1, m // s + 1)), (m, n), interpolation=cv2.INTER_CUBIC) * np.sqrt(s)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return (acc * 255).astype(np.uint8)


def akaze_cv(img, max_points=(1 << 18) - 1):
    """lib.rs:64-79"""
    ak = cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2, max_points)
    kps, desc = ak.detectAndCompute(img, None)
    k = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave, p.class_id) for p in kps],
                 dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                        ("octave", "<i4"), ("class_id", "<i4")])
    return k, (desc if desc is not None else np.zeros((0, 61), np.uint8))


def make_akaze():
    out = {"opencv_version": np.array(cv2.__version__)}
    imgs = {"a": synth(320, 21, 384),          # non-square, 3 octaves
            "b": synth(512, 1)}                # 4 octaves
    rng = np.random.default_rng(5)
    g = synth(200, 22, 260)
    col = np.clip(g[..., None].astype(np.int32) + rng.integers(-40, 40, (200, 260, 3)), 0, 255).astype(np.uint8)
    imgs["c"] = col                            # BGR input (cvtColor path), 2 octaves
    # dense keypoints (contrast-stretched, decimated scene): exercises the suppression corner cases
    # (neighbours at exactly distance r, multi-partner windows) that sparse images never hit
    big = synth(1024, 23).astype(np.float32)
    big = np.clip((big - big.mean()) * 2.2 + 128, 0, 255).astype(np.uint8)
    imgs["d"] = cv2.resize(big, (448, 448), interpolation=cv2.INTER_AREA)
    for name, img in imgs.items():
        k, d = akaze_cv(img)
        out[f"{name}_img"], out[f"{name}_kps"], out[f"{name}_desc"] = img, k, d
    k, d = akaze_cv(imgs["b"], 100)            # max_points path (top-100 by response)
    out["b100_kps"], out["b100_desc"] = k, d
    np.savez_compressed(os.path.join(HERE, "akaze_golden.npz"), **out)
    print("akaze_golden.npz:", {k: v.shape for k, v in out.items()})


def make_config1():
    """BASELINE config 1 at FULL size: g = synth(1024, 0), w = warpPerspective(g, H_TRUE) with the reference's
    warp settings (mod.rs:286-294); extract both (lib.rs:64-79), knnMatch(w -> g, 2) + ratio in {0.3, 0.7, 0.8}
    (lib.rs:101-111), findHomography(RANSAC, 3.0) (mod.rs:243-250).  Expected [probe]: 3163 / 3033 keypoints,
    1347 / 2269 / 2377 matches.  The 1024^2 images exercise octave 3 and its 17/20/24/29-step FED chains."""
    g = synth(1024, 0)
    w = cv2.warpPerspective(g, H_TRUE, (1024, 1024), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=1)
    kg, dg = akaze_cv(g)
    kw, dw = akaze_cv(w)
    idx, dist = knn_cv(dw, dg)
    out = {"opencv_version": np.array(cv2.__version__), "g_img": g, "w_img": w, "g_kps": kg, "g_desc": dg, "w_kps": kw,
           "w_desc": dw, "knn_idx": idx, "knn_dist": dist}
    for ratio in (0.3, 0.7, 0.8):
        keep = dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * np.float32(ratio)
        qi = np.nonzero(keep)[0]
        src = np.stack([kg["x"][idx[qi, 0]], kg["y"][idx[qi, 0]]], 1).astype(np.float32)      # tile points
        dst = np.stack([kw["x"][qi], kw["y"][qi]], 1).astype(np.float32)                      # warped-image points
        H, mask = cv2.findHomography(src, dst, cv2.RANSAC, 3.0)
        tag = f"r{int(ratio * 10)}"
        out[f"{tag}_query"], out[f"{tag}_H"], out[f"{tag}_mask"] = qi.astype(np.int32), H, mask.ravel().astype(np.uint8)
        print(f"config1 ratio {ratio}: {len(qi)} matches, {int(mask.sum())} inliers, H err "
              f"{np.abs(H - H_TRUE).max() / np.abs(H_TRUE).max():.2e}")
    print("config1:", len(kg), "/", len(kw), "keypoints")
    np.savez_compressed(os.path.join(HERE, "config1_golden.npz"), **out)


PNP_K = np.array([[800., 0, 512], [0, 820., 500], [0, 0, 1]])
# (n, outlier fraction, pixel noise sigma, iterations, reprojection threshold, confidence, relief)
PNP_CASES = [(100, 0.3, 0.5, 100, 8.0, 0.99, "cube"), (1000, 0.4, 0.5, 1000, 2.0, 0.99, "cube"),
             (50, 0.0, 0.0, 100, 8.0, 0.99, "cube"), (300, 0.5, 1.0, 500, 3.0, 0.999, "cube"),
             (6, 0.0, 0.1, 100, 8.0, 0.99, "cube"), (2000, 0.2, 0.3, 1000, 2.0, 0.99, "cube"),
             (500, 0.6, 0.5, 1000, 2.0, 0.99, "cube"), (5, 0.0, 0.1, 100, 8.0, 0.99, "cube"),
             (400, 0.3, 0.5, 300, 3.0, 0.99, "terrain"), (64, 0.9, 0.5, 50, 1.0, 0.99, "cube")]


def pnp_case(seed, n, out_frac, noise, relief):
    """n object points seen by a camera 8 units away; `terrain` = mild relief over a plane (DEM-like)."""
    rng = np.random.default_rng(seed)
    rv = rng.normal(0, 0.4, 3)
    tv = np.array([0.3, -0.2, 8.0]) + rng.normal(0, 0.5, 3)
    obj = rng.uniform(-2, 2, (n, 3))
    if relief == "terrain":
        obj[:, 2] = 0.15 * np.sin(obj[:, 0] * 2.0) * np.cos(obj[:, 1] * 1.5) + rng.normal(0, 0.02, n)
    R, _ = cv2.Rodrigues(rv)
    P = obj @ R.T + tv
    img = np.stack([PNP_K[0, 0] * P[:, 0] / P[:, 2] + PNP_K[0, 2], PNP_K[1, 1] * P[:, 1] / P[:, 2] + PNP_K[1, 2]], 1)
    img += rng.normal(0, noise, (n, 2))
    no = int(out_frac * n)
    o = rng.choice(n, no, replace=False)
    img[o] = rng.uniform(0, 1024, (no, 2))
    return obj, img


def make_pnp():
    """cv2.solvePnPRansac(obj, img, K, zeros(4,1), ..., flags=EPNP) — mod.rs:347-361 — plus
    cv2.solvePnP(EPNP) on f64 and f32 point sets, cv2.Rodrigues and cv2.SVDecomp probes."""
    out = {"opencv_version": np.array(cv2.__version__), "n_cases": np.array(len(PNP_CASES)), "K": PNP_K}
    for i, (n, of, noise, iters, thr, conf, relief) in enumerate(PNP_CASES):
        obj, img = pnp_case(i, n, of, noise, relief)
        ok, r, t, inl = cv2.solvePnPRansac(obj, img, PNP_K, np.zeros((4, 1)), None, None, False, iters, thr, conf, None,
                                           cv2.SOLVEPNP_EPNP)
        out[f"c{i}_obj"], out[f"c{i}_img"] = obj, img
        out[f"c{i}_params"] = np.array([iters, thr, conf])
        out[f"c{i}_found"] = np.array(bool(ok))
        out[f"c{i}_rvec"], out[f"c{i}_tvec"] = r.ravel(), t.ravel()
        out[f"c{i}_inliers"] = np.zeros(0, np.int32) if inl is None else inl.ravel().astype(np.int32)
        print(f"pnp case {i}: n={n} found={ok} inliers={len(out[f'c{i}_inliers'])}")
    # the same cases through the P3P kernel (4-point samples; the final pose is still EPnP on the inliers)
    for i, (n, of, noise, iters, thr, conf, relief) in enumerate(PNP_CASES):
        obj, img = pnp_case(i, n, of, noise, relief)
        ok, r, t, inl = cv2.solvePnPRansac(obj, img, PNP_K, np.zeros((4, 1)), None, None, False, iters, thr, conf, None,
                                           cv2.SOLVEPNP_P3P)
        out[f"p{i}_found"] = np.array(bool(ok))
        out[f"p{i}_rvec"], out[f"p{i}_tvec"] = r.ravel(), t.ravel()
        out[f"p{i}_inliers"] = np.zeros(0, np.int32) if inl is None else inl.ravel().astype(np.int32)
    # exactly 4 correspondences: OpenCV runs P3P once whatever the flag is
    for j in range(12):
        obj, img = pnp_case(200 + j, 4, 0.0, 0.3, "cube")
        ok, r, t, inl = cv2.solvePnPRansac(obj, img, PNP_K, np.zeros((4, 1)), None, None, False, 100, 8.0, 0.99, None,
                                           cv2.SOLVEPNP_EPNP)
        out[f"q{j}_obj"], out[f"q{j}_img"], out[f"q{j}_found"] = obj, img, np.array(bool(ok))
        out[f"q{j}_rt"] = np.r_[r.ravel(), t.ravel()]
    out["n_four"] = np.array(12)
    # plain EPnP solves (f64 points and f32 points take different undistortPoints precisions)
    rng = np.random.default_rng(99)
    for j, n in enumerate((6, 7, 12, 40, 300)):
        obj, img = pnp_case(100 + j, n, 0.0, 0.5, "cube")
        ok, r, t = cv2.solvePnP(obj, img, PNP_K, np.zeros((4, 1)), flags=cv2.SOLVEPNP_EPNP)
        o32, i32 = obj.astype(np.float32), img.astype(np.float32)
        ok2, r2, t2 = cv2.solvePnP(o32, i32, PNP_K, np.zeros((4, 1)), flags=cv2.SOLVEPNP_EPNP)
        out[f"e{j}_obj"], out[f"e{j}_img"] = obj, img
        out[f"e{j}_rt64"] = np.r_[r.ravel(), t.ravel()]
        out[f"e{j}_rt32"] = np.r_[r2.ravel(), t2.ravel()]
    out["n_epnp"] = np.array(5)
    # Rodrigues both ways, Jacobi SVD signs on small symmetric matrices
    rv = rng.normal(0, 1.0, (8, 3))
    out["rod_rvec"] = rv
    out["rod_R"] = np.stack([cv2.Rodrigues(v)[0] for v in rv])
    out["rod_back"] = np.stack([cv2.Rodrigues(R)[0].ravel() for R in out["rod_R"]])
    A = rng.normal(size=(6, 7, 3))
    B = np.einsum("kij,kil->kjl", A, A)
    out["svd_in"] = B
    sv = [cv2.SVDecomp(b) for b in B]
    out["svd_w"] = np.stack([s[0].ravel() for s in sv])
    out["svd_u"] = np.stack([s[1] for s in sv])
    out["svd_vt"] = np.stack([s[2] for s in sv])
    np.savez_compressed(os.path.join(HERE, "pnp_golden.npz"), **out)
    print("pnp_golden.npz written")


def make_warp():
    """cv2.warpPerspective(src, M, size, INTER_LINEAR, BORDER_CONSTANT, (1,1,1,1)) — mod.rs:286-294"""
    import synthdata
    out = {"opencv_version": np.array(cv2.__version__)}
    gray = synthdata.synth_image(200, 264, 3)
    bgra = np.stack([gray, gray[::-1], gray[:, ::-1], np.full_like(gray, 255)], -1).copy()
    Hs = [np.eye(3), H_TRUE, np.array([[1.2, 0.3, -50], [-0.2, 0.9, 80], [3e-4, -2e-4, 1.0]]),
          np.array([[0.5, 0, 10.25], [0, 0.5, 3.5], [0, 0, 1.0]]), np.array([[0.9, -0.4, 200.], [0.4, 0.9, -100], [1e-5, 1e-5, 1]]),
          np.array([[1.0, 0, 0.5], [0, 1.0, 0.5], [0, 0, 1.0]])]
    out["gray"], out["bgra"], out["H"] = gray, bgra, np.stack(Hs)
    sizes = [(264, 200), (300, 170), (97, 333)]
    out["sizes"] = np.array(sizes)
    for i, H in enumerate(Hs):
        for j, (w, h) in enumerate(sizes):
            out[f"g{i}_{j}"] = cv2.warpPerspective(gray, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                                   borderValue=(1, 1, 1, 1))
        out[f"c{i}"] = cv2.warpPerspective(bgra, H, (264, 200), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                                           borderValue=(1, 1, 1, 1))
    np.savez_compressed(os.path.join(HERE, "warp_golden.npz"), **out)
    print("warp_golden.npz written")


def make_l2():
    """cv2.BFMatcher(NORM_L2).knnMatch(q, t, 2) on the seeded f32 descriptor sets of synthdata.l2_descriptors
    (north_star's float matcher); only cv2's answers are stored"""
    import synthdata
    out = {"opencv_version": np.array(cv2.__version__)}
    for name in synthdata.L2_CASES:
        q, t = synthdata.l2_descriptors(name)
        m = cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(q, t, 2)
        out[f"{name}_idx"] = np.array([[x.trainIdx for x in r] for r in m], np.int32)
        out[f"{name}_dist"] = np.array([[x.distance for x in r] for r in m], np.float32)
        out[f"{name}_checksum"] = np.array([float(q.astype(np.float64).sum()), float(t.astype(np.float64).sum())])
    np.savez_compressed(os.path.join(HERE, "l2_golden.npz"), **out)
    print("l2_golden.npz written")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("match", "all"):
        make_match()
    if what in ("ransac", "all") and "make_ransac" in globals():
        make_ransac()
    if what in ("lmeds", "all") and "make_lmeds" in globals():
        make_lmeds()
    if what in ("rho", "all"):
        make_rho()
    if what in ("akaze", "all") and "make_akaze" in globals():
        make_akaze()
    if what in ("config1", "all"):
        make_config1()
    if what in ("pnp", "all") and "make_pnp" in globals():
        make_pnp()
    if what in ("pnp_iter", "all") and "make_pnp_iter" in globals():
        make_pnp_iter()
    if what in ("warp", "all") and "make_warp" in globals():
        make_warp()
    if what in ("l2", "all") and "make_l2" in globals():
        make_l2()
