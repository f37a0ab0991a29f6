"""GPU: the device LoD tiler + DB build (dunk_db_build_from_bands) equals the composition of its
separately verified parts: oracle-resampled + oracle-converted BGRA tiles pushed through
dunk_db_append_tiles give the same rows (descriptors, keypoints in scene coordinates, image ids) and
the same ref_image table.  Resampled pixel values: parity with GDAL unpinned (oracle/lod_oracle.py)."""
import numpy as np
import pytest

from oracle import lod_oracle as lo

pytestmark = pytest.mark.gpu


def scene_bands(n=1024, seed=4):
    import synthdata
    g = synthdata.synth_image(n, n, seed).astype(np.float32)
    r = g * np.float32(0.0011) + np.float32(0.002)
    gg = np.roll(g, 5, 0) * np.float32(0.0010) + np.float32(0.003)
    b = np.roll(g, -7, 1) * np.float32(0.0009) + np.float32(0.001)
    return r, gg, b, (0.002, 0.2825, 0.003, 0.258, 0.001, 0.2305)


@pytest.mark.parametrize("resample", ["area", "lanczos"])
def test_build_from_bands_equals_tilewise_build(dunk, ctx, resample):
    fd = dunk.feature_database
    r, g, b, mm = scene_bands()
    lods = 3
    db = fd.DescriptorDatabase(ctx, capacity=200000)
    n_tiles, (tw, th) = db.build_from_bands(r, g, b, mm, lods, resample)
    assert (tw, th) == (256, 256) and n_tiles == 16 + 4 + 1
    ref = fd.DescriptorDatabase(ctx, capacity=200000)
    images = []
    for lod, col, row, x0, y0, s, tile in lo.lod_tiles(r, g, b, mm, lods, resample):
        iid = ref.create_image(x0, y0, x0 + tw * s - 1, y0 + th * s - 1, lod)
        images.append((iid, x0, y0, x0 + tw * s - 1, y0 + th * s - 1, lod))
        ref.append_tiles(tile[None], [x0], [y0], [s], [iid])
    assert len(db) == len(ref) > 100
    a, c = db.rows(), ref.rows()
    assert a.tobytes() == c.tobytes()
    for im in images:
        assert tuple(int(v) for v in db.read_image_from_id(im[0]).tolist()) == im
    # keyed read through the LoD join
    top = db.read_keypoints_from_lod(lods - 1)
    assert len(top) > 0 and set(top["image_id"].tolist()) == {n_tiles}
    db.close(); ref.close()
