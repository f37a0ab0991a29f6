"""CPU: pins oracle/db_oracle.py against a real SQL engine.  The reference's keyed reads are diesel queries on
Postgres (feature_database/src/keypointdb.rs:38-90, imagedb.rs:39-66) over the tables of its migrations
(migrations/2024-03-21-110256_image/up.sql, 2024-03-21-110413_keypoint/up.sql).  No Postgres here, but the statements
are plain SQL-92 — inner join on the foreign key, comparisons, ORDER BY ... DESC, LIMIT — so sqlite3 (stdlib) executes
the same tables and the same queries (column types real / integer as in the migrations; `floor` / `ceil` of the bounds
are applied by the Rust code before binding, keypointdb.rs:80-83).  Rows with equal `response` may come back in any
order from either engine, so ties are compared as sets per response value."""
import sqlite3

import numpy as np
import pytest

from oracle import db_oracle as do

SCHEMA = """
CREATE TABLE ref_image (id INTEGER PRIMARY KEY AUTOINCREMENT, x_start integer NOT NULL, y_start integer NOT NULL,
                        x_end integer NOT NULL, y_end integer NOT NULL, level_of_detail integer NOT NULL);
CREATE TABLE keypoint (id INTEGER PRIMARY KEY AUTOINCREMENT, x_coord real NOT NULL, y_coord real NOT NULL, size real NOT NULL,
                       angle real NOT NULL, response real NOT NULL, octave integer NOT NULL, class_id integer NOT NULL,
                       descriptor blob NOT NULL, image_id integer NOT NULL REFERENCES ref_image (id));
"""
LIMIT = 2 ** 18 - 1


@pytest.fixture(scope="module")
def tables():
    rng = np.random.default_rng(12)
    n_img, n = 40, 6000
    images = [(i + 1, int(rng.integers(0, 4000)), int(rng.integers(0, 4000)), 0, 0, int(rng.integers(0, 4))) for i in range(n_img)]
    images = [(i, xs, ys, xs + int(rng.integers(1, 2000)), ys + int(rng.integers(1, 2000)), lod) for i, xs, ys, _, _, lod in images]
    x = rng.uniform(0, 6000, n).astype(np.float32)
    y = rng.uniform(0, 6000, n).astype(np.float32)
    resp = rng.choice(np.linspace(0.001, 0.2, 300).astype(np.float32), n)          # many exact ties
    image_id = rng.integers(1, n_img + 1, n).astype(np.int32)
    con = sqlite3.connect(":memory:")
    con.executescript(SCHEMA)
    con.executemany("INSERT INTO ref_image (x_start, y_start, x_end, y_end, level_of_detail) VALUES (?,?,?,?,?)", [im[1:] for im in images])
    con.executemany("INSERT INTO keypoint (x_coord, y_coord, size, angle, response, octave, class_id, descriptor, image_id) "
                    "VALUES (?,?,?,?,?,?,?,?,?)",
                    [(float(x[i]), float(y[i]), 4.8, 0.0, float(resp[i]), 0, 0, b"\0" * 61, int(image_id[i])) for i in range(n)])
    return con, images, x, y, resp, image_id


def same_result(sql_rows, oracle_idx, resp):
    """identical multiset per response value, and the response column is non-increasing in both"""
    sql_ids = np.array([r[0] for r in sql_rows], dtype=np.int64) - 1           # SERIAL ids are 1-based
    assert len(sql_ids) == len(oracle_idx)
    if len(sql_ids) == 0:
        return
    assert (np.diff(resp[sql_ids]) <= 0).all() and (np.diff(resp[oracle_idx]) <= 0).all()
    assert np.array_equal(resp[sql_ids], resp[oracle_idx])
    for v in np.unique(resp[oracle_idx]):
        assert set(sql_ids[resp[sql_ids] == v]) == set(oracle_idx[resp[oracle_idx] == v])


def test_read_keypoints_from_image_id(tables):
    con, images, x, y, resp, image_id = tables
    lod = np.array([im[5] for im in images])
    for iid in (1, 7, 40, 41):
        rows = con.execute("SELECT keypoint.id FROM keypoint WHERE image_id = ? ORDER BY response DESC LIMIT ?", (iid, LIMIT)).fetchall()
        same_result(rows, do.select_rows(x, y, resp, image_id, lod, f_image_id=iid), resp)


def test_read_keypoints_from_lod_and_coordinates(tables):
    con, images, x, y, resp, image_id = tables
    lod = np.array([im[5] for im in images])
    for level in range(5):
        rows = con.execute("SELECT keypoint.id FROM keypoint INNER JOIN ref_image ON keypoint.image_id = ref_image.id "
                           "WHERE ref_image.level_of_detail = ? ORDER BY keypoint.response DESC LIMIT ?", (level, LIMIT)).fetchall()
        same_result(rows, do.select_rows(x, y, resp, image_id, lod, f_lod=level), resp)
    for box in ((1000.3, 2000.7, 3000.2, 2500.9), (0.0, 0.0, 6000.0, 6000.0), (10.5, 10.5, 10.6, 10.6)):
        fl = [float(np.floor(np.float32(box[0]))), float(np.floor(np.float32(box[1]))), float(np.ceil(np.float32(box[2]))),
              float(np.ceil(np.float32(box[3])))]                                   # keypointdb.rs:80-83 binds floor / ceil
        rows = con.execute("SELECT keypoint.id FROM keypoint INNER JOIN ref_image ON keypoint.image_id = ref_image.id "
                           "WHERE ref_image.level_of_detail = ? AND x_coord >= ? AND x_coord <= ? AND y_coord >= ? AND y_coord <= ? "
                           "ORDER BY keypoint.response DESC LIMIT ?", (1, fl[0], fl[2], fl[1], fl[3], LIMIT)).fetchall()
        same_result(rows, do.select_rows(x, y, resp, image_id, lod, f_lod=1, box=box), resp)


def test_limit_applies_after_the_ordering(tables):
    con, images, x, y, resp, image_id = tables
    lod = np.array([im[5] for im in images])
    rows = con.execute("SELECT keypoint.id FROM keypoint INNER JOIN ref_image ON keypoint.image_id = ref_image.id "
                       "WHERE ref_image.level_of_detail = ? ORDER BY keypoint.response DESC LIMIT ?", (2, 25)).fetchall()
    got = do.select_rows(x, y, resp, image_id, lod, f_lod=2, limit=25)
    assert len(rows) == len(got) == 25
    sql_ids = np.array([r[0] for r in rows]) - 1
    assert np.array_equal(resp[sql_ids], resp[got])            # the same 25 response values (membership of the last tie may differ)


def test_find_images(tables):
    con, images, *_ = tables
    for level in range(4):
        rows = con.execute("SELECT id FROM ref_image WHERE level_of_detail = ?", (level,)).fetchall()
        assert sorted(r[0] for r in rows) == sorted(do.find_images(images, level))
        box = (500, 700, 2500, 2600)
        rows = con.execute("SELECT id FROM ref_image WHERE x_end >= ? AND x_start <= ? AND y_end >= ? AND y_start <= ? AND level_of_detail = ?",
                           (box[0], box[2], box[1], box[3], level)).fetchall()
        assert sorted(r[0] for r in rows) == sorted(do.find_images(images, level, box))
