"""GPU parity: warp_image_perspective through the C ABI is bit-exact with cv2 4.13.0
`warpPerspective(INTER_LINEAR, BORDER_CONSTANT, Scalar(1,1,1,1))` goldens and with the oracle."""
import os

import numpy as np
import pytest

from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "warp_golden.npz"))


@pytest.mark.parametrize("i", range(6))
def test_warp_gray_bit_exact(dunk, ctx, i):
    hg = dunk.homographier
    for j, (w, h) in enumerate(G["sizes"]):
        got = hg.warp_image_perspective(hg.Cmat(G["gray"], np.uint8), hg.Cmat(G["H"][i], np.float64), (int(w), int(h)), ctx)
        assert np.array_equal(got.mat, G[f"g{i}_{j}"])


@pytest.mark.parametrize("i", range(6))
def test_warp_bgra_bit_exact(dunk, ctx, i):
    hg = dunk.homographier
    got = hg.warp_image_perspective(hg.Cmat(G["bgra"], np.uint8), hg.Cmat(G["H"][i], np.float64), None, ctx)
    assert np.array_equal(got.mat, G[f"c{i}"])


def test_reference_test_warp_image_empty(dunk, ctx):
    """mod.rs:682-707"""
    hg = dunk.homographier
    img = hg.raster_to_mat(np.arange(4 * 4 * 4, dtype=np.uint8).reshape(-1, 4), 4, 4)
    out = hg.warp_image_perspective(img, hg.Cmat(np.eye(3), np.float64), None, ctx)
    for r in range(4):
        for c in range(4):
            assert np.array_equal(img.at_2d(r, c), out.at_2d(r, c))


def test_full_size_frame_vs_oracle_and_errors(dunk, ctx):
    import synthdata
    hg = dunk.homographier
    scene = synthdata.synth_image(1024, 1024, 5)
    H = synthdata.H_CONFIG1
    got = hg.warp_image_perspective(hg.Cmat(scene, np.uint8), H, (1024, 1024), ctx)
    assert np.array_equal(got.mat, wo.warp_perspective(scene, H, 1024, 1024, 1))
    with pytest.raises(hg.MatError):
        hg.warp_image_perspective(hg.Cmat(scene, np.uint8), np.eye(4), None, ctx)
    # a singular matrix maps everything to the source origin's neighbourhood, like OpenCV (zero inverse)
    out = hg.warp_image_perspective(hg.Cmat(scene, np.uint8), np.zeros((3, 3)), (64, 64), ctx)
    assert np.array_equal(out.mat, wo.warp_perspective(scene, np.zeros((3, 3)), 64, 64, 1))


def test_batch_dev_cuts_frames_from_a_scene(dunk, ctx):
    import ctypes as C
    import torch
    import synthdata
    from cubesat_apds_b200._lib import check, load
    scene = synthdata.synth_image(700, 900, 9)
    Hs = np.stack([synthdata.window_homography(100.0 + 30 * k, 50.0 + 20 * k, 40 + k) for k in range(5)])
    dev = torch.device("cuda", ctx.device)
    s_dev = torch.from_numpy(scene).to(dev)
    out = torch.empty(5 * 256 * 256, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    slot = ctx.reserve_slot()
    try:
        check(load().dunk_warp_perspective_batch_dev(ctx.handle, slot, s_dev.data_ptr(), 700, 900, 1, 900,
                                                     Hs.ctypes.data_as(C.c_void_p), 5, 256, 256, None, out.data_ptr()))
        ctx.sync(slot)
    finally:
        ctx.release_slot(slot)
    got = out.cpu().numpy().reshape(5, 256, 256)
    for k in range(5):
        assert np.array_equal(got[k], wo.warp_perspective(scene, Hs[k], 256, 256, 1))
