"""A/B: planes of every level, keypoints and descriptors with the register-window k_hessian_reg / k_prep_level_reg / k_fed_reg / k_gray_gauss9_reg / k_contrast_modg_reg
kernels vs their predecessors (DUNK_HESSIAN_OLD=1 DUNK_PREP_OLD=1 DUNK_FED_OLD=1 DUNK_LEVEL0_OLD=1), compared bit for bit.
Usage: python tools/ab_kernels.py  (spawns itself twice)"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dump(path):
    import cubesat_apds_b200 as dunk
    import synthdata
    ctx = dunk.Context(0, 2)
    out = {}
    rng = np.random.default_rng(1)
    g640 = synthdata.synth_image(480, 640, 7)
    bgra = np.stack([g640, np.roll(g640, 3, 0), np.roll(g640, 5, 1), np.full_like(g640, 255)], axis=-1)
    bgr = np.ascontiguousarray(bgra[..., :3])
    for name, img in (("s1024", synthdata.synth_image(1024, 1024, 3)), ("s700x520", synthdata.synth_image(520, 700, 5)),
                      ("s1372", synthdata.synth_image(1372, 1372, 9)), ("bgra640", bgra), ("bgr640", bgr)):
        n = 1
        i = 0
        while i < n:
            Lt, Lx, Ly, Ldet, k, n = dunk._extract.debug_level(img, i, ctx)
            out[f"{name}_{i}_Lx"], out[f"{name}_{i}_Ly"], out[f"{name}_{i}_Ldet"], out[f"{name}_{i}_Lt"] = Lx, Ly, Ldet, Lt
            i += 1
        r = dunk.feature_extraction.akaze_keypoint_descriptor_extraction_def(img, None, ctx)
        out[f"{name}_kps"], out[f"{name}_desc"] = r.keypoints, r.descriptors
    np.savez(path, **out)
    os._exit(0)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        dump(sys.argv[1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    a, b = os.path.join(ROOT, "gpurun_out", "ab_new.npz"), os.path.join(ROOT, "gpurun_out", "ab_old.npz")
    subprocess.check_call([sys.executable, __file__, a], env={**os.environ})
    subprocess.check_call([sys.executable, __file__, b], env={**os.environ, "DUNK_HESSIAN_OLD": "1", "DUNK_PREP_OLD": "1", "DUNK_FED_OLD": "1", "DUNK_LEVEL0_OLD": "1"})
    A, B = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        same = A[k].tobytes() == B[k].tobytes()
        if not same:
            bad += 1
            if A[k].dtype == np.float32 and A[k].shape == B[k].shape:
                d = np.abs(A[k] - B[k])
                print("DIFF", k, A[k].shape, "max", d.max(), "count", int((d > 0).sum()), "first", np.argwhere(d > 0)[:3].tolist())
            else:
                print("DIFF", k, A[k].shape, B[k].shape)
    print("keys", len(A.files), "different", bad)
    os.remove(a); os.remove(b)
