"""CPU: the `__host__ __device__` PnP math the CUDA kernels are built from
(cubesat-apds_b200/csrc/pnp_math.cuh: Jacobi SVD, Rodrigues, EPnP, f32 reprojection error),
compiled for the host by a test-only harness (tests/hostcheck/), agrees with the oracle and with
the cv2 goldens.  This checks the device code's arithmetic without a GPU; the `-m gpu` tests
(tests/test_pnp_gpu.py) check the same code on the device through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pnp_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "pnp_golden.npz"))
K = G["K"]


@pytest.fixture(scope="module")
def hc():
    out = os.path.join(HERE, "_build", "libpnp_hostcheck.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(HERE, "hostcheck", "pnp_hostcheck.cpp"), "-o", out], check=True)
    lib = C.CDLL(out)
    lib.hc_epnp.restype = C.c_double
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def host_epnp(hc, obj32, img32, idx, f32n):
    r, t = np.zeros(3), np.zeros(3)
    idx_p = None if idx is None else ptr(np.ascontiguousarray(idx, np.int32))
    n = len(obj32) if idx is None else len(idx)
    hc.hc_epnp(ptr(obj32), ptr(img32), idx_p, n, ptr(np.ascontiguousarray(K)), int(f32n), ptr(r), ptr(t))
    return r, t


def test_jacobi_svd_signs(hc):
    for B, w, u, vt in zip(G["svd_in"], G["svd_w"], G["svd_u"], G["svd_vt"]):
        W, U, Vt = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3))
        hc.hc_jacobi_svd3(ptr(np.ascontiguousarray(B)), ptr(W), ptr(U), ptr(Vt))
        assert np.abs(W - w).max() < 1e-12 * w.max()
        assert np.abs(U - u).max() < 1e-12 and np.abs(Vt - vt).max() < 1e-12


def test_rodrigues(hc):
    for rv, R, back in zip(G["rod_rvec"], G["rod_R"], G["rod_back"]):
        Rm, r = np.zeros((3, 3)), np.zeros(3)
        hc.hc_rodrigues_to_matrix(ptr(np.ascontiguousarray(rv)), ptr(Rm))
        hc.hc_rodrigues_to_vector(ptr(np.ascontiguousarray(R)), ptr(r))
        assert np.abs(Rm - R).max() < 1e-12 and np.abs(r - back).max() < 1e-12


@pytest.mark.parametrize("j", range(int(G["n_epnp"])))
def test_epnp_f32_points_match_cv2(hc, j):
    """cv2.solvePnP(EPNP) on f32 points (= what every call inside solvePnPRansac's loop sees)"""
    o32 = np.ascontiguousarray(G[f"e{j}_obj"], np.float32)
    i32 = np.ascontiguousarray(G[f"e{j}_img"], np.float32)
    r, t = host_epnp(hc, o32, i32, None, True)
    assert np.abs(np.r_[r, t] - G[f"e{j}_rt32"]).max() < 1e-9


@pytest.mark.parametrize("i", [0, 1, 3, 8])
def test_final_solve_on_golden_inliers(hc, i):
    """the last step of solvePnPRansac: EPnP (f64-normalised) on the f32-rounded inliers"""
    o32 = np.ascontiguousarray(G[f"c{i}_obj"], np.float32)
    i32 = np.ascontiguousarray(G[f"c{i}_img"], np.float32)
    r, t = host_epnp(hc, o32, i32, G[f"c{i}_inliers"], False)
    assert np.abs(r - G[f"c{i}_rvec"]).max() < 1e-9 and np.abs(t - G[f"c{i}_tvec"]).max() < 1e-9


@pytest.mark.parametrize("i", [0, 1, 3])
def test_minimal_samples_score_like_the_oracle(hc, i):
    """5-point hypotheses of the fixed-seed stream ("identical seeded hypothesis sets"): the inlier
    count of every hypothesis RANSAC could accept (> 4 inliers) equals the oracle's.  5-point samples
    leave M^T M a 2-D null space whose basis is rounding noise (oracle header), so samples that
    contain outliers give unrelated garbage poses on both sides — always with <= 4 inliers."""
    obj, img = G[f"c{i}_obj"], G[f"c{i}_img"]
    thr2 = np.float32(float(G[f"c{i}_params"][1]) ** 2)
    o32, i32 = np.ascontiguousarray(obj, np.float32), np.ascontiguousarray(img, np.float32)
    od, idd = o32.astype(np.float64), i32.astype(np.float64)
    acceptable = 0
    for s in po.sample_stream(len(obj), 40):
        r, t = host_epnp(hc, o32, i32, s, True)
        e = np.zeros(len(obj), np.float32)
        hc.hc_reproj_err(ptr(r), ptr(t), ptr(np.ascontiguousarray(K)), ptr(o32), ptr(i32), len(obj), ptr(e))
        assert np.array_equal(e, po.reproj_err_f32(o32, i32, r, t, K))          # same pose -> bit-equal f32 errors
        ro, to = po.solve_pnp_epnp(od[s], idd[s], K, True)
        c_host = int((e <= thr2).sum())
        c_oracle = int((po.reproj_err_f32(o32, i32, ro, to, K) <= thr2).sum())
        assert c_host == c_oracle or max(c_host, c_oracle) <= 4
        acceptable += c_oracle > 4
    assert acceptable >= 1


def test_p3p_host_math_matches_oracle(hc):
    """the device's P3P (quartic + triad) on the host vs the oracle: same success flag, same pose up to the quartic's
    conditioning"""
    hc.hc_p3p.restype = C.c_int
    rng = np.random.default_rng(3)
    close = total = 0
    for _ in range(200):
        rv = rng.normal(0, 0.5, 3)
        tv = np.array([0.3, -0.2, 8.0]) + rng.normal(0, 0.5, 3)
        obj = rng.uniform(-2, 2, (4, 3)).astype(np.float32)
        P = obj.astype(np.float64) @ po.rodrigues_to_matrix(rv).T + tv
        img = (np.stack([K[0, 0] * P[:, 0] / P[:, 2] + K[0, 2], K[1, 1] * P[:, 1] / P[:, 2] + K[1, 2]], 1)
               + rng.normal(0, 0.3, (4, 2))).astype(np.float32)
        r, t = np.zeros(3), np.zeros(3)
        ok = hc.hc_p3p(ptr(obj), ptr(img), None, ptr(np.ascontiguousarray(K)), 1, ptr(r), ptr(t))
        sol = po.solve_pnp_p3p(obj.astype(np.float64), img.astype(np.float64), K, True)
        assert bool(ok) == (sol is not None)
        if ok:
            total += 1
            close += np.abs(np.r_[r, t] - np.r_[sol[0], sol[1]]).max() < 1e-6
    assert total > 150 and close >= 0.95 * total


GI = np.load(os.path.join(HERE, "golden", "pnp_iter_golden.npz"))


@pytest.mark.parametrize("i", range(int(G["n_cases"])))
def test_iterative_refinement_on_golden_inliers(hc, i):
    """pnp_refine (the code the GPU runs after the final EPnP when the method is SOLVEPNP_ITERATIVE), compiled for the
    host: from the EPnP pose of cv2's inlier set to cv2's ITERATIVE pose"""
    if not bool(GI[f"c{i}_found"]) or len(G[f"c{i}_obj"]) <= 5:
        pytest.skip("no pose / minimal point set (OpenCV returns the kernel's pose unrefined)")
    obj32 = np.ascontiguousarray(G[f"c{i}_obj"], np.float32)
    img32 = np.ascontiguousarray(G[f"c{i}_img"], np.float32)
    idx = np.ascontiguousarray(GI[f"c{i}_inliers"], np.int32)
    r, t = host_epnp(hc, obj32, img32, idx, False)
    hc.hc_refine.restype = C.c_double
    rms = hc.hc_refine(ptr(obj32), ptr(img32), ptr(idx), len(idx), ptr(np.ascontiguousarray(K)), ptr(r), ptr(t))
    assert np.isfinite(rms)
    assert np.abs(r - GI[f"c{i}_rvec"]).max() < 1e-6
    assert np.abs(t - GI[f"c{i}_tvec"]).max() < 1e-6 * max(1.0, np.abs(GI[f"c{i}_tvec"]).max())
