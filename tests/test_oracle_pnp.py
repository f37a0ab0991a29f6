"""CPU: the PnP oracle reproduces cv2 4.13.0 `solvePnPRansac(..., EPNP)` (the call made at
homographier/src/homographier/mod.rs:347-361) on the committed golden vectors: inlier index lists
identical, rvec/tvec within 1e-9; plain EPnP solves within 1e-9 for f64 and f32 point sets;
cv::Rodrigues and the Jacobi SVD sign conventions within 1e-12."""
import os

import numpy as np
import pytest

from oracle import pnp_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "pnp_golden.npz"))
N = int(G["n_cases"])
K = G["K"]
POSE_ATOL = 1e-9


@pytest.mark.parametrize("i", range(N))
def test_solve_pnp_ransac_matches_cv2(i):
    iters, thr, conf = G[f"c{i}_params"]
    found, r, t, inl = po.solve_pnp_ransac(G[f"c{i}_obj"], G[f"c{i}_img"], K, int(iters), float(thr), float(conf))
    assert found == bool(G[f"c{i}_found"])
    assert np.array_equal(inl, G[f"c{i}_inliers"])
    if found:
        assert np.abs(r - G[f"c{i}_rvec"]).max() < POSE_ATOL
        assert np.abs(t - G[f"c{i}_tvec"]).max() < POSE_ATOL


@pytest.mark.parametrize("j", range(int(G["n_epnp"])))
def test_epnp_matches_cv2(j):
    obj, img = G[f"e{j}_obj"], G[f"e{j}_img"]
    r, t = po.solve_pnp_epnp(obj, img, K)
    assert np.abs(np.r_[r, t] - G[f"e{j}_rt64"]).max() < POSE_ATOL
    o32, i32 = obj.astype(np.float32).astype(np.float64), img.astype(np.float32).astype(np.float64)
    r, t = po.solve_pnp_epnp(o32, i32, K, f32_normalised=True)
    assert np.abs(np.r_[r, t] - G[f"e{j}_rt32"]).max() < POSE_ATOL


def test_rodrigues_both_ways():
    for rv, R, back in zip(G["rod_rvec"], G["rod_R"], G["rod_back"]):
        assert np.abs(po.rodrigues_to_matrix(rv) - R).max() < 1e-12
        assert np.abs(po.rodrigues_to_vector(R) - back).max() < 1e-12


def test_jacobi_svd_signs():
    for B, w, u, vt in zip(G["svd_in"], G["svd_w"], G["svd_u"], G["svd_vt"]):
        W, U, Vt = po.jacobi_svd(B)
        assert np.abs(W - w).max() < 1e-12 * w.max()
        assert np.abs(U - u).max() < 1e-12 and np.abs(Vt - vt).max() < 1e-12


def test_too_few_points():
    """reference test pnp_solver_ransac_no_work_lthan_3_points (mod.rs:627-638): error, not None"""
    with pytest.raises(ValueError):
        po.solve_pnp_ransac(np.array([[1., 2, 3], [4, 5, 6]]), np.array([[1., 2], [4, 5]]), np.zeros((3, 3)), 50, 2.0, 0.99)


def test_sample_stream_is_the_homography_stream_with_5_points():
    s = po.sample_stream(100, 4)
    assert s.shape == (4, 5) and all(len(set(r)) == 5 for r in s.tolist())


@pytest.mark.parametrize("i", range(N))
def test_solve_pnp_ransac_p3p_matches_cv2(i):
    """flags = SOLVEPNP_P3P (the method the reference's own `pnp_solver_works` test passes, mod.rs:640-681)"""
    iters, thr, conf = G[f"c{i}_params"]
    found, r, t, inl = po.solve_pnp_ransac_p3p(G[f"c{i}_obj"], G[f"c{i}_img"], K, int(iters), float(thr), float(conf))
    assert found == bool(G[f"p{i}_found"])
    assert np.array_equal(inl, G[f"p{i}_inliers"])
    if found:
        assert np.abs(r - G[f"p{i}_rvec"]).max() < 1e-6 and np.abs(t - G[f"p{i}_tvec"]).max() < 1e-6


def _reproj(obj, img, rt):
    return np.abs(po.project_points(obj, rt[:3], rt[3:], K) - img).max(axis=1)


def test_four_points_run_p3p_once():
    """n == 4: one P3P solve.  Both implementations must return a pose that reproduces the first three points exactly
    and the fourth equally well (near-double roots of the quartic make the pose itself ill-conditioned in ~4 % of cases)"""
    same = 0
    for j in range(int(G["n_four"])):
        obj, img = G[f"q{j}_obj"], G[f"q{j}_img"]
        found, r, t, inl = po.solve_pnp_ransac_p3p(obj, img, K)
        assert found == bool(G[f"q{j}_found"])
        if not found:
            continue
        o32, i32 = obj.astype(np.float32).astype(np.float64), img.astype(np.float32).astype(np.float64)
        e_ours, e_cv = _reproj(o32, i32, np.r_[r, t]), _reproj(o32, i32, G[f"q{j}_rt"])
        assert e_ours[:3].max() < 1e-3 and e_cv[:3].max() < 1e-3
        assert abs(e_ours[3] - e_cv[3]) <= 0.05 * max(e_cv[3], 1e-3)
        same += np.abs(np.r_[r, t] - G[f"q{j}_rt"]).max() < 1e-6
    assert same >= int(G["n_four"]) - 2


# ---- SOLVEPNP_ITERATIVE ------------------------------------------------------------------------------------
GI = np.load(os.path.join(os.path.dirname(__file__), "golden", "pnp_iter_golden.npz"))


@pytest.mark.parametrize("i", range(N))
def test_solve_pnp_ransac_iterative_matches_cv2(i):
    """same RANSAC stage (identical inlier lists), final pose = minimum of the reprojection error over the inliers;
    cv2's own LM stops within ~1e-8 of it"""
    assert np.allclose([G[f"c{i}_obj"].sum(), G[f"c{i}_img"].sum()], GI[f"c{i}_checksum"])
    iters, thr, conf = G[f"c{i}_params"]
    found, r, t, inl = po.solve_pnp_ransac_iterative(G[f"c{i}_obj"], G[f"c{i}_img"], K, int(iters), float(thr), float(conf))
    assert found == bool(GI[f"c{i}_found"])
    assert np.array_equal(inl, GI[f"c{i}_inliers"])
    if found:
        assert np.abs(r - GI[f"c{i}_rvec"]).max() < 1e-6
        assert np.abs(t - GI[f"c{i}_tvec"]).max() < 1e-6 * max(1.0, np.abs(GI[f"c{i}_tvec"]).max())
