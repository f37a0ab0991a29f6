"""GPU: feature_database keyed reads / ref_image table / dump + load vs the SQL-semantics oracle
(oracle/db_oracle.py): result sets identical (row ids, order, every column bit-exact)."""
import os

import numpy as np
import pytest

from oracle import db_oracle as do

pytestmark = pytest.mark.gpu


def make_db(dunk, ctx, n=20000, n_images=12, seed=3):
    rng = np.random.default_rng(seed)
    fd = dunk.feature_database
    db = fd.DescriptorDatabase(ctx, capacity=n + 16)
    images = []
    for k in range(n_images):
        lod = k % 3
        x0, y0 = int(rng.integers(0, 4000)), int(rng.integers(0, 4000))
        iid = db.create_image(x0, y0, x0 + 1024, y0 + 1024, lod)
        images.append((iid, x0, y0, x0 + 1024, y0 + 1024, lod))
    desc = rng.integers(0, 256, (n, 61), dtype=np.uint8)
    desc[:, 60] &= 0x3F
    kps = np.zeros(n, dtype=dunk._lib.KEYPOINT_DTYPE)
    kps["x"] = rng.uniform(0, 5000, n).astype(np.float32)
    kps["y"] = rng.uniform(0, 5000, n).astype(np.float32)
    kps["size"] = 4.8
    kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    kps["response"] = rng.choice(rng.uniform(1e-3, 0.2, n // 4), n).astype(np.float32)     # many exact ties
    kps["octave"] = rng.integers(0, 4, n)
    kps["class_id"] = rng.integers(0, 16, n)
    image_id = rng.integers(1, n_images + 2, n).astype(np.int32)                          # some ids have no ref_image row
    db.append(desc, kps, image_id)
    return db, images, desc, kps, image_id


def check_rows(rows, idx, desc, kps, image_id):
    assert np.array_equal(rows["id"], idx + 1)
    assert np.array_equal(rows["descriptor"], desc[idx])
    assert np.array_equal(rows["image_id"], image_id[idx])
    assert np.array_equal(rows["x_coord"], kps["x"][idx]) and np.array_equal(rows["y_coord"], kps["y"][idx])
    for name in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(rows[name], kps[name][idx])
    assert (np.diff(rows["response"]) <= 0).all()


def test_keyed_reads_match_sql_semantics(dunk, ctx):
    db, images, desc, kps, image_id = make_db(dunk, ctx)
    lod_of = [im[5] for im in images]
    x, y, r = kps["x"], kps["y"], kps["response"]
    rows = db.read_keypoints_from_image_id(5)
    check_rows(rows, do.select_rows(x, y, r, image_id, lod_of, f_image_id=5), desc, kps, image_id)
    rows = db.read_keypoints_from_lod(1)
    check_rows(rows, do.select_rows(x, y, r, image_id, lod_of, f_lod=1), desc, kps, image_id)
    box = (1000.3, 499.5, 3000.2, 2500.7)
    rows = db.read_keypoints_from_coordinates(*box, 2)
    idx = do.select_rows(x, y, r, image_id, lod_of, f_lod=2, box=box)
    assert len(idx) > 0
    check_rows(rows, idx, desc, kps, image_id)
    assert len(db.read_keypoints_from_image_id(999)) == 0
    # LIMIT keeps the strongest responses
    sub = db.select(level_of_detail=0, limit=100)
    check_rows(sub.rows(), do.select_rows(x, y, r, image_id, lod_of, f_lod=0, limit=100), desc, kps, image_id)
    sub.close()
    one = db.read_keypoint_from_id(17)
    assert one["id"] == 17 and np.array_equal(one["descriptor"], desc[16])
    with pytest.raises(dunk.feature_database.NotFound):
        db.read_keypoint_from_id(len(db) + 1)
    db.close()


def test_image_table(dunk, ctx):
    db, images, *_ = make_db(dunk, ctx, n=64)
    for im in images:
        got = db.read_image_from_id(im[0])
        assert tuple(int(v) for v in got.tolist()) == im
    with pytest.raises(dunk.feature_database.NotFound):
        db.read_image_from_id(len(images) + 1)
    for lod in range(3):
        assert db.find_images_from_lod(lod) == do.find_images(images, lod)
        box = (1500, 1500, 3000, 3000)
        assert db.find_images_from_dimensions(*box, lod) == do.find_images(images, lod, box)
    db.close()


def test_selected_rows_can_be_matched_against(dunk, ctx):
    """a keyed read is a shard: 2-NN against it equals the oracle on the same rows"""
    from oracle import match_oracle as mo
    db, images, desc, kps, image_id = make_db(dunk, ctx, n=6000)
    lod_of = [im[5] for im in images]
    sub = db.select(level_of_detail=1)
    idx = do.select_rows(kps["x"], kps["y"], kps["response"], image_id, lod_of, f_lod=1)
    q = desc[::37][:100].copy()
    q[:, 3] ^= 0x55
    got = sub.knn2(q)
    oi, od = mo.knn2(q, desc[idx])
    assert np.array_equal(got["i1"], oi[:, 0]) and np.array_equal(got["d1"], od[:, 0])
    assert np.array_equal(got["i2"], oi[:, 1]) and np.array_equal(got["d2"], od[:, 1])
    sub.close(); db.close()


def test_dump_and_load_round_trip(dunk, ctx, tmp_path):
    db, images, desc, kps, image_id = make_db(dunk, ctx, n=5000)
    path = os.path.join(tmp_path, "shard.dunkdb")
    db.save(path)
    assert os.path.getsize(path) >= 5000 * (64 + 28 + 4)
    db2 = dunk.feature_database.DescriptorDatabase.load(path, ctx, min_capacity=6000)
    assert len(db2) == len(db)
    a, b = db.rows(), db2.rows()
    assert a.tobytes() == b.tobytes()
    assert db2.find_images_from_lod(1) == db.find_images_from_lod(1)
    q = desc[:50]
    assert db.knn2(q).tobytes() == db2.knn2(q).tobytes()
    db2.append(desc[:10], kps[:10], image_id[:10])           # loaded shards stay appendable
    assert len(db2) == len(db) + 10
    with pytest.raises(dunk._lib.DunkError):
        dunk.feature_database.DescriptorDatabase.load(os.path.join(tmp_path, "missing.dunkdb"), ctx)
    db.close(); db2.close()
