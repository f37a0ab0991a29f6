"""GPU: BASELINE.json's full-size configurations through size-independent properties (the oracle is too slow at
these sizes): config 4 (10980 x 10980 scene -> LoD tiles -> HBM descriptor DB) and config 2 (256 frames, extraction
only).  Parity itself is established at oracle-sized inputs in the other test files."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config4_scene_to_database(dunk, ctx):
    import synthdata
    fd = dunk.feature_database
    n, lods = 10980, 4
    g = synthdata.synth_scene(n, seed=11).astype(np.float32)
    r = g * np.float32(0.0011) + np.float32(0.002)
    gr = g * np.float32(0.0010) + np.float32(0.003)
    b = g * np.float32(0.0009) + np.float32(0.001)
    mm = (0.002, 0.2825, 0.003, 0.258, 0.001, 0.2305)
    db = fd.DescriptorDatabase(ctx, capacity=3_000_000)
    n_tiles, (tw, th) = db.build_from_bands(r, gr, b, mm, lods, "area")
    # preprocessor/src/main.rs:212-216: tile = scene / 2^(lods-1) (integer division), remainder dropped
    assert (tw, th) == (n // 8, n // 8) == (1372, 1372)
    assert n_tiles == 64 + 16 + 4 + 1
    rows = len(db)
    assert rows > 100_000
    # ref_image rows: extents follow main.rs:283-289
    for lod, count in ((0, 64), (1, 16), (2, 4), (3, 1)):
        ids = db.find_images_from_lod(lod)
        assert len(ids) == count
        im = db.read_image_from_id(ids[-1])
        s = tw << lod
        assert im["x_end"] - im["x_start"] == s - 1 and im["x_start"] % s == 0 and im["level_of_detail"] == lod
    # every keypoint lies inside its tile's extent in scene pixels (main.rs:300-301 mapping)
    sub = db.select(level_of_detail=2)
    k = sub.rows()
    sub.close()
    assert len(k) > 1000 and (np.diff(k["response"]) <= 0).all()
    for iid in np.unique(k["image_id"]):
        im = db.read_image_from_id(int(iid))
        sel = k[k["image_id"] == iid]
        assert (sel["x_coord"] >= im["x_start"]).all() and (sel["x_coord"] <= im["x_end"] + 1).all()
        assert (sel["y_coord"] >= im["y_start"]).all() and (sel["y_coord"] <= im["y_end"] + 1).all()
    # a keyed read by coordinates returns exactly the LoD-0 rows inside the box
    box = (2000.0, 3000.0, 2600.0, 3500.0)
    inside = db.read_keypoints_from_coordinates(*box, 0)
    assert len(inside) > 0
    assert (inside["x_coord"] >= 2000).all() and (inside["x_coord"] <= 2600).all()
    assert (inside["y_coord"] >= 3000).all() and (inside["y_coord"] <= 3500).all()
    # matching a tile's own descriptors against the DB finds them at distance 0
    probe = db.read_keypoints_from_image_id(int(db.find_images_from_lod(1)[3]))[:500]
    top = db.knn2(probe["descriptor"])
    assert (top["d1"] == 0).all()
    db.close()


def test_config2_batch_extraction_is_deterministic_and_batch_invariant(dunk, ctx):
    import synthdata
    from cubesat_apds_b200 import _extract
    frames = np.stack([synthdata.synth_image(1024, 1024, 100 + i) for i in range(4)])
    batch = np.ascontiguousarray(frames[np.arange(256) % 4])
    out = _extract.extract_batch(batch, dunk._lib.MAX_POINTS, ctx)
    counts = np.array([len(o.keypoints) for o in out])
    assert (counts > 1500).all()
    for i in range(4, 256):
        assert counts[i] == counts[i % 4]
        assert out[i].descriptors.tobytes() == out[i % 4].descriptors.tobytes()
        assert out[i].keypoints.tobytes() == out[i % 4].keypoints.tobytes()
    single = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(frames[2], None, ctx)
    assert single.descriptors.tobytes() == out[2].descriptors.tobytes()
