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


def test_build_from_device_bands_and_clear(dunk, ctx):
    """dunk_db_build_from_bands_dev (bands already in HBM) gives the rows of the host-band call; dunk_db_clear
    empties keypoint + ref_image and a rebuild into the same buffers reproduces them."""
    import ctypes as C
    import torch
    from cubesat_apds_b200._lib import check, load
    fd = dunk.feature_database
    r, g, b, mm = scene_bands()
    db = fd.DescriptorDatabase(ctx, capacity=200000)
    n_tiles, _ = db.build_from_bands(r, g, b, mm, 3)
    want = db.rows().tobytes()
    images = [db.read_image_from_id(i + 1).tolist() for i in range(n_tiles)]
    db.clear()
    assert len(db) == 0
    with pytest.raises(fd.NotFound):
        db.read_image_from_id(1)
    dev = [torch.from_numpy(x).cuda() for x in (r, g, b)]
    m = np.ascontiguousarray(mm, dtype=np.float64)
    n, tw, th = C.c_int(0), C.c_int(0), C.c_int(0)
    check(load().dunk_db_build_from_bands_dev(db.handle, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), 1024, 1024,
                                              m.ctypes.data, 3, 0, 0, C.byref(n), C.byref(tw), C.byref(th)))
    assert n.value == n_tiles and (tw.value, th.value) == (256, 256)
    assert db.rows().tobytes() == want
    assert [db.read_image_from_id(i + 1).tolist() for i in range(n_tiles)] == images
    db.close()
