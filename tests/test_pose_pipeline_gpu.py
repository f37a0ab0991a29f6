"""GPU: the pose stage composed behind the homography (north_star stage 3 "yields the attitude estimate"; BASELINE
config 5): dunk_register_frames_pose == the stage-by-stage calls through the mirrored reference API
(get_world_coordinates elevationdb.rs:64-104 -> pnp_solver_ransac mod.rs:320-369), == cv2.solvePnPRansac run live
on the same correspondences, and the recovered pose is the camera that rendered the frame."""
import numpy as np
import pytest

import synthdata

pytestmark = pytest.mark.gpu
S = 2048


@pytest.fixture(scope="module")
def world(dunk, ctx):
    fd = dunk.feature_database
    scene = synthdata.synth_scene(S, seed=11)
    tiles = np.stack([scene[r * 1024:(r + 1) * 1024, c * 1024:(c + 1) * 1024] for r in range(2) for c in range(2)])
    db = fd.DescriptorDatabase(ctx, capacity=60000)
    db.append_tiles(tiles, x_off=[0, 1024, 0, 1024], y_off=[0, 0, 1024, 1024], scale=1.0, image_ids=[1, 2, 3, 4])
    Hs, Rs, ts, resid = synthdata.config5_views(3, S, 4242)
    hg = dunk.homographier
    frames = np.stack([hg.warp_image_perspective(hg.Cmat(scene, np.uint8), H, (1024, 1024), ctx).mat for H in Hs])
    gt_e, heights = synthdata.scene_dem(S)
    geo = fd.Geotransform(synthdata.scene_geotransform(), gt_e, heights, ctx)
    origin = synthdata.scene_origin(S)
    pose = fd.PoseStage(geo, synthdata.CAMERA_K, origin, 1000, 3.0, 0.99, 1)
    yield {"db": db, "frames": frames, "Hs": Hs, "Rs": Rs, "ts": ts, "geo": geo, "origin": origin, "pose": pose, "resid": resid}
    db.close()
    geo.close()


def stagewise(dunk, ctx, w, frame):
    """the same composition through the mirrored reference API, one call per reference function"""
    fe, hg = dunk.feature_extraction, dunk.homographier
    d, k, _ = w["db"].read_rows(0, len(w["db"]))
    q = fe.akaze_keypoint_descriptor_extraction_def(frame, None, ctx)
    m = fe.get_knn_matches(q.descriptors, d, 2, 0.8, ctx)
    p_q, p_r = fe.get_points_from_matches(q.keypoints, k, m)
    H, mask = hg.find_homography_mat(p_q, p_r, hg.HomographyMethod.RANSAC, 3.0, ctx)
    inl = mask.mat.ravel() > 0
    xyz, miss = w["geo"].world_coordinates(p_r[inl, 0].astype(np.float64), p_r[inl, 1].astype(np.float64))
    assert miss == 0
    obj = xyz - w["origin"]
    img = p_q[inl].astype(np.float64)
    sol = hg.pnp_solver_ransac((obj, img), synthdata.CAMERA_K, 1000, 3.0, 0.99, None, hg.SolvePnPMethod.SOLVEPNP_EPNP, ctx)
    return H, int(inl.sum()), obj, img, sol


def test_fused_pose_equals_stagewise_and_cv2(dunk, ctx, world):
    import cv2
    w = world
    res, poses = w["db"].register_frames(w["frames"], ratio=0.8, reproj_threshold=3.0, pose=w["pose"])
    plain = w["db"].register_frames(w["frames"], ratio=0.8, reproj_threshold=3.0)
    assert plain.tobytes() == res.tobytes()                       # the pose stage does not disturb stages 1-3
    for i in range(len(w["frames"])):
        H, n_inl, obj, img, sol = stagewise(dunk, ctx, w, w["frames"][i])
        assert res["found"][i] == 1 and res["inliers"][i] == n_inl and n_inl > 300
        assert np.array_equal(res["H"][i].reshape(3, 3), H.mat)
        assert poses["found"][i] == 1 and sol is not None
        d_r = np.abs(poses["rvec"][i] - sol.rvec.mat.ravel()).max()
        d_t = np.abs(poses["tvec"][i] - sol.tvec.mat.ravel()).max()
        # cv2 live on the same correspondences (the reference's call, mod.rs:347-361)
        ok, rv, tv, inl = cv2.solvePnPRansac(obj, img, synthdata.CAMERA_K, np.zeros((4, 1)), None, None, False, 1000, 3.0, 0.99, None,
                                             cv2.SOLVEPNP_EPNP)
        # the geometry is ill-conditioned on purpose (1.2 deg field of view from 500 km: a rotation of 1e-6 rad trades against
        # 0.5 m of lateral position), so the comparison with cv2 is stated as what it means physically
        rot_cv, pos_cv = synthdata.pose_errors([poses["rvec"][i]], [poses["tvec"][i]], [1], [cv2.Rodrigues(rv)[0]], [tv.ravel()])
        print(f"frame {i}: inliers fused {poses['inliers'][i]} / stagewise {len(sol.inliers.mat)} / cv2 {len(inl)}; fused - stagewise "
              f"|drvec| {d_r:.2e} |dtvec| {d_t:.2e}; fused vs cv2: rotation {rot_cv[0]:.2e} deg, camera centre {pos_cv[0]:.2e} m")
        assert poses["inliers"][i] == len(sol.inliers.mat)
        # same kernels on the same f32 inputs: identical pose
        assert d_r <= 1e-12 and d_t <= 1e-6
        assert ok and np.array_equal(inl.ravel(), sol.inliers.mat.ravel())
        assert rot_cv[0] < 1e-4 and pos_cv[0] < 1.0           # < 1e-4 degrees, < 1 m of 500 km
    # the pose is the camera that rendered the frame (narrow field of view: rotation / position coupled, see bench notes)
    rot, pos = synthdata.pose_errors(poses["rvec"], poses["tvec"], poses["found"], w["Rs"], w["ts"])
    print("rotation error deg", rot, "camera centre error m", pos, "homography fit residual px", w["resid"])
    assert (rot < 8.0).all() and (pos < 80e3).all()
    for r, H in zip(res, w["Hs"]):
        Hi = np.linalg.inv(H); Hi /= Hi[2, 2]
        assert np.abs(r["H"].reshape(3, 3) - Hi).max() / np.abs(Hi).max() < 5e-3


def test_pose_stage_degenerate_inputs(dunk, ctx, world):
    """a flat frame has no homography -> no pose; the batch's other frames are unaffected; bad configs fail loudly"""
    w = world
    flat = np.full((1024, 1024), 90, np.uint8)
    res, poses = w["db"].register_frames(np.stack([flat, w["frames"][0]]), pose=w["pose"])
    assert res["found"][0] == 0 and poses["found"][0] == 0 and (poses["rvec"][0] == 0).all()
    assert res["found"][1] == 1 and poses["found"][1] == 1
    fd = dunk.feature_database
    bad = fd.PoseStage(w["geo"], synthdata.CAMERA_K, w["origin"], 100, 3.0, 0.99, method=7)
    with pytest.raises(dunk.DunkError) as e:
        w["db"].register_frames(w["frames"][:1], pose=bad)
    assert e.value.code == dunk._lib.ERR_BAD_ARG


def test_shard_group_world1_equals_unsharded(dunk, ctx, world):
    """the sharded step with a one-rank group (no NCCL) == the plain pipeline, record for record"""
    w = world
    fd = dunk.feature_database
    g = fd.ShardGroup(ctx, 0, 1)
    shard = g.balance(w["db"])
    assert len(shard) == len(w["db"]) == g.total_rows and g.base(0) == 0 and g.base(1) == len(shard)
    assert shard.rows().tobytes() == w["db"].rows().tobytes()
    ref, ref_p = w["db"].register_frames(w["frames"], pose=w["pose"])
    got, got_p = g.register_frames(shard, w["frames"], pose=w["pose"])
    assert ref.tobytes() == got.tobytes() and ref_p.tobytes() == got_p.tobytes()
    q = w["db"].read_descriptors(5, 300)
    assert g.match(shard, q, 0.9, 0).tobytes() == w["db"].match(q, 0.9).tobytes()
    shard.close()
    g.close()
