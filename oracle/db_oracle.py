"""Oracle for the feature_database keyed reads — TEST INFRASTRUCTURE ONLY.

Restates the SQL the reference issues through diesel (feature_database/src/keypointdb.rs:38-90,
src/imagedb.rs:39-66) over in-memory numpy columns: WHERE clause, inner join on ref_image for the
level of detail, ORDER BY response DESC (ties: ascending row id — Postgres leaves them unspecified),
LIMIT 2^18 - 1.  The reference holds no golden vectors for this path (its DB tests need a live Postgres).
PINNED instead against a real SQL engine: tests/test_oracle_db_sql.py creates the tables of the reference's
migrations (ref_image, keypoint) in sqlite3 and runs the same statements (join on the foreign key, bounds
floor()/ceil()-ed as keypointdb.rs:80-83 binds them, ORDER BY response DESC, LIMIT); result sets are identical
(rows of equal response compared as sets — SQL leaves their order unspecified, the oracle and the device use
ascending row id).
"""
import numpy as np

LIMIT = 2 ** 18 - 1


def select_rows(x, y, response, image_id, image_lod, f_image_id=-1, f_lod=-1, box=None, limit=LIMIT):
    """returns the 0-based row indices of the result set, in result order.
    image_lod[k] = level_of_detail of image id k + 1."""
    n = len(x)
    ok = np.ones(n, bool)
    if f_image_id >= 0:
        ok &= image_id == f_image_id
    if f_lod >= 0:
        valid = (image_id >= 1) & (image_id <= len(image_lod))
        lod = np.full(n, -1, np.int64)
        lod[valid] = np.asarray(image_lod)[image_id[valid] - 1]
        ok &= lod == f_lod
    if box is not None:
        x0, y0, x1, y1 = box
        ok &= (x >= np.floor(np.float32(x0))) & (x <= np.ceil(np.float32(x1))) & (y >= np.floor(np.float32(y0))) & (y <= np.ceil(np.float32(y1)))
    idx = np.nonzero(ok)[0]
    order = np.argsort(-response[idx].astype(np.float64), kind="stable")
    return idx[order][:limit]


def find_images(images, lod, box=None):
    """images: rows (id, x_start, y_start, x_end, y_end, level_of_detail)"""
    out = []
    for im in images:
        if im[5] != lod:
            continue
        if box is not None:
            xs, ys, xe, ye = box
            if not (im[3] >= xs and im[1] <= xe and im[4] >= ys and im[2] <= ye):
                continue
        out.append(int(im[0]))
    return out
