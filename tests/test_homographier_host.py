"""CPU: host-side mirror of homographier's checked-Mat helpers (the reference's own unit tests,
homographier/src/homographier/mod.rs:474-625, restated)."""
import numpy as np
import pytest


def test_cmat_init_empty_is_error(dunk):
    hg = dunk.homographier
    with pytest.raises(hg.MatError):                      # cmat_init mod.rs:474-477
        hg.Cmat(np.zeros((0, 0), np.uint8))


def test_cmat_from_slice(dunk):
    hg = dunk.homographier
    c = hg.Cmat.from_2d_slice([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float64)   # mod.rs:514-553
    assert c.mat.shape == (2, 3) and c.at_2d(1, 2) == 6.0
    with pytest.raises(hg.MatError):
        hg.Cmat.from_2d_slice([[1.0, 2.0], [3.0]], np.float64)
    with pytest.raises(hg.MatError):
        hg.Cmat.from_2d_slice([], np.float64)


def test_raster_to_mat_works(dunk):
    """mod.rs:555-603: BGRA order + row-major."""
    hg = dunk.homographier
    size = 4
    px = np.ones((size * size, 4), np.uint8)
    for i in range(size * size):
        row, col = i // size + 1, i % size + 1
        px[i] = (1, col, row, 1)                          # RGBA: g = col, b = row
    m = hg.raster_to_mat(px, size, size)
    assert m.mat.shape == (size, size, 4)
    for r in range(size):
        for c in range(size):
            b, g, rr, a = m.mat[r, c]
            assert (b, g, rr, a) == (r + 1, c + 1, 1, 1)
    with pytest.raises(hg.MatError) as e:
        hg.raster_to_mat(px[:-1], size, size)
    assert e.value.kind == "Unknown"


def test_cmat_at_2d_out_of_range(dunk):
    """mod.rs:605-625: StsOutOfRange = -211."""
    hg = dunk.homographier
    c = hg.Cmat.zeros(3, 3)
    with pytest.raises(hg.MatError) as e:
        c.at_2d(4, 4)
    assert e.value.code == -211
    with pytest.raises(hg.MatError) as e:
        c.at_2d(3, 0)
    assert e.value.code == -211
