"""GPU parity: CUDA AKAZE (through the C ABI) vs the oracle restatement and the cv2 4.13.0 goldens.
Tolerances are stated in tests/akaze_compare.py."""
import os

import numpy as np
import pytest

from oracle import akaze_oracle as ao
from tests.akaze_compare import assert_parity, compare

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "akaze_golden.npz"))


def test_scale_space_planes_vs_oracle(dunk, ctx):
    """per-stage intermediates: Lt, Lx, Ly, Ldet of every level and the k-contrast."""
    img = G["a_img"]
    lv, kc = ao.build_scale_space(img)
    for i, e in enumerate(lv):
        Lt, Lx, Ly, Ldet, k, nl = dunk._extract.debug_level(img, i, ctx)
        assert nl == len(lv)
        assert abs(k - kc) <= 1e-6 * kc
        for name, got in (("Lt", Lt), ("Lx", Lx), ("Ly", Ly)):
            ref = e[name]
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (i, name)
        b = e["border"]
        if b * 2 + 2 < min(Ldet.shape):
            ref = e["Ldet"][b - 1:-b + 1, b - 1:-b + 1]
            got = Ldet[b - 1:-b + 1, b - 1:-b + 1]
            assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, (i, "Ldet")


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_extract_vs_cv2_golden(dunk, ctx, name):
    r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(G[f"{name}_img"], None, ctx)
    rep = compare(G[f"{name}_kps"], G[f"{name}_desc"], r.keypoints, r.descriptors)
    print(name, rep)
    assert_parity(rep)
    if rep["recall"] == 1.0 and rep["precision"] == 1.0:
        assert np.array_equal(r.keypoints["class_id"], G[f"{name}_kps"]["class_id"])   # OpenCV output order


def test_extract_vs_oracle_fresh_image(dunk, ctx):
    rng = np.random.default_rng(77)
    base = rng.standard_normal((40, 52))
    img = np.kron(base, np.ones((8, 8)))
    img = img + 0.3 * rng.standard_normal(img.shape)
    # smooth a little so that extrema are well defined
    for _ in range(3):
        img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5
    img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    kps, desc = ao.detect_and_compute(img)
    r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(img, None, ctx)
    rep = compare(kps, desc, r.keypoints, r.descriptors)
    print(rep)
    assert len(kps) > 50
    assert_parity(rep)


def test_extract_odd_size_bgr_vs_oracle(dunk, ctx):
    """517 x 333 BGR (no extent is a multiple of any tile size: every kernel runs partial and border tiles)"""
    import synthdata
    g = synthdata.synth_image(333, 517, 9)
    img = np.stack([g, np.roll(g, 3, 1), np.roll(g, -2, 0)], -1).copy()
    kps, desc = ao.detect_and_compute(img)
    r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(img, None, ctx)
    rep = compare(kps, desc, r.keypoints, r.descriptors)
    print(rep)
    assert len(kps) > 100
    assert_parity(rep)


@pytest.mark.parametrize("h,w,channels", [(210, 386, 4), (128, 130, 1), (200, 770, 1)])
def test_extract_even_awkward_shapes_vs_oracle(dunk, ctx, h, w, channels):
    """even widths take the register-window kernels (spans of 64 columns, bands of <= 128 rows): shapes whose last span
    and band are partial, a level just above the FED cascade's size threshold (386 x 210), a level just above the
    kernels' minimum (130 x 128), and a wide flat one; checked against the oracle like the odd-size case"""
    import synthdata
    g = synthdata.synth_image(h, w, 21 + w)
    img = g if channels == 1 else np.dstack([g, np.roll(g, 2, 1), np.roll(g, -3, 0), np.full_like(g, 255)]).copy()
    kps, desc = ao.detect_and_compute(img)
    r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(img, None, ctx)
    rep = compare(kps, desc, r.keypoints, r.descriptors)
    print(rep)
    assert len(kps) >= 5
    assert_parity(rep)


def test_gray_bgr_bgra_identical(dunk, ctx):
    g = G["a_img"][:200, :240]
    fe = dunk.feature_extraction
    a = fe.akaze_keypoint_descriptor_extraction_def(g, None, ctx)
    b = fe.akaze_keypoint_descriptor_extraction_def(np.dstack([g, g, g]), None, ctx)
    c = fe.akaze_keypoint_descriptor_extraction_def(np.dstack([g, g, g, np.full_like(g, 255)]), None, ctx)
    assert len(a.keypoints) > 20
    assert np.array_equal(a.keypoints, b.keypoints) and np.array_equal(a.descriptors, c.descriptors)


def test_max_points(dunk, ctx):
    r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(G["b_img"], 100, ctx)
    ref = G["b100_kps"]
    assert len(r.keypoints) == 100
    assert np.allclose(np.sort(r.keypoints["response"]), np.sort(ref["response"]), rtol=1e-3)
    rep = compare(ref, G["b100_desc"], r.keypoints, r.descriptors)
    assert rep["recall"] >= 0.99


def test_batch_equals_single(dunk, ctx):
    imgs = np.stack([G["a_img"], G["a_img"][::-1].copy(), G["a_img"][:, ::-1].copy()])
    batch = dunk._extract.extract_batch(imgs, ctx=ctx)
    for i in range(3):
        single = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(imgs[i], None, ctx)
        assert np.array_equal(batch[i].keypoints, single.keypoints)
        assert np.array_equal(batch[i].descriptors, single.descriptors)


def test_determinism_and_to_db_type(dunk, ctx):
    fe = dunk.feature_extraction
    a = fe.akaze_keypoint_descriptor_extraction_def(G["c_img"], None, ctx)
    b = fe.akaze_keypoint_descriptor_extraction_def(G["c_img"], None, ctx)
    assert np.array_equal(a.keypoints, b.keypoints) and np.array_equal(a.descriptors, b.descriptors)
    rows = a.to_db_type(42)
    assert len(rows) == len(a.keypoints) and rows[0].image_id == 42 and len(rows[0].descriptor) == 61
    assert (a.descriptors[:, 60] <= 63).all()


def test_errors(dunk, ctx):
    fe = dunk.feature_extraction
    with pytest.raises(dunk.DunkError) as e:
        fe.akaze_keypoint_descriptor_extraction_def(np.zeros((0, 0), np.uint8), None, ctx)
    assert e.value.code == -215
    with pytest.raises(dunk.DunkError):
        fe.akaze_keypoint_descriptor_extraction_def(np.zeros((64, 64), np.float32), None, ctx)
    # blank image: no keypoints, no error
    r = fe.akaze_keypoint_descriptor_extraction_def(np.full((128, 160), 7, np.uint8), None, ctx)
    assert len(r.keypoints) == 0 and r.descriptors.shape == (0, 61)
