"""GPU: the register-window scale-space kernels (k_gray_gauss9_reg, k_contrast_modg_reg, k_prep_level_reg,
k_fed_reg, k_hessian_reg) give the same BITS as the tiled shared-memory kernels they replaced, which stay in
the library for odd / tiny levels and can be forced with DUNK_*_OLD=1.  The switches are read once per process,
so tools/ab_kernels.py runs the two generations in two child processes and compares every plane of every level,
the keypoints and the descriptors of five images (gray 1024^2, 700x520, 1372^2, BGRA and BGR 640x480)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_window_kernels_equal_tiled_kernels_bit_for_bit():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_kernels.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    last = [ln for ln in out.stdout.splitlines() if ln.startswith("keys ")]
    assert last, out.stdout[-2000:]
    n_keys, n_diff = int(last[-1].split()[1]), int(last[-1].split()[3])
    assert n_keys >= 300 and n_diff == 0, out.stdout[-2000:]
