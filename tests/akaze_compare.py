"""Shared keypoint/descriptor comparison for the AKAZE parity tests (oracle and GPU).

Tolerances (north_star: "keypoint coordinates and responses within a stated tolerance"):
  * keypoints are matched 1:1 by class_id and position within POS_TOL = 0.5 px; the tests
    require recall and precision >= MIN_RECALL over the cv2 set;
  * for matched pairs: position error <= POS_STRICT px, response relative error <= RESP_RTOL,
    size / octave identical, orientation within ANGLE_TOL degrees for >= ANGLE_FRAC of the pairs
    (a borderline sample can move the dominant-orientation window discontinuously);
  * descriptors: mean Hamming distance over matched pairs <= DESC_MEAN_BITS (of 486 bits) and
    >= DESC_EXACT_FRAC of them bit-identical.
"""
import numpy as np

POS_TOL = 0.5
POS_STRICT = 0.05
RESP_RTOL = 1e-3
ANGLE_TOL = 0.1
ANGLE_FRAC = 0.98
MIN_RECALL = 0.99
DESC_MEAN_BITS = 2.0
DESC_EXACT_FRAC = 0.90


def match_keypoints(ref, got):
    """greedy 1:1 matching by class_id and nearest position (< POS_TOL). Returns index pairs."""
    pairs = []
    used = np.zeros(len(got), bool)
    for cid in np.unique(ref["class_id"]):
        ri = np.nonzero(ref["class_id"] == cid)[0]
        gi = np.nonzero(got["class_id"] == cid)[0]
        if len(gi) == 0:
            continue
        gx, gy = got["x"][gi], got["y"][gi]
        for i in ri:
            d = np.hypot(gx - ref["x"][i], gy - ref["y"][i])
            d[used[gi]] = np.inf
            j = int(np.argmin(d))
            if d[j] < POS_TOL:
                used[gi[j]] = True
                pairs.append((i, gi[j]))
    return np.array(pairs, dtype=np.int64).reshape(-1, 2)


def compare(ref_kps, ref_desc, got_kps, got_desc):
    p = match_keypoints(ref_kps, got_kps)
    n_ref, n_got = len(ref_kps), len(got_kps)
    rep = {"n_ref": n_ref, "n_got": n_got, "matched": len(p),
           "recall": len(p) / max(n_ref, 1), "precision": len(p) / max(n_got, 1)}
    if len(p) == 0:
        return rep
    r, g = ref_kps[p[:, 0]], got_kps[p[:, 1]]
    rep["pos_err_max"] = float(np.hypot(r["x"] - g["x"], r["y"] - g["y"]).max())
    rep["resp_rel_max"] = float((np.abs(r["response"] - g["response"]) / np.abs(r["response"])).max())
    rep["size_equal"] = bool(np.allclose(r["size"], g["size"], rtol=1e-6))
    rep["octave_equal"] = bool((r["octave"] == g["octave"]).all())
    da = np.abs(r["angle"] - g["angle"])
    da = np.minimum(da, 360 - da)
    rep["angle_frac_ok"] = float((da <= ANGLE_TOL).mean())
    rep["angle_err_median"] = float(np.median(da))
    if ref_desc is not None and got_desc is not None:
        hd = np.unpackbits(ref_desc[p[:, 0]] ^ got_desc[p[:, 1]], axis=1).sum(1)
        rep["desc_mean_bits"] = float(hd.mean())
        rep["desc_exact_frac"] = float((hd == 0).mean())
        rep["desc_max_bits"] = int(hd.max())
    return rep


def assert_parity(rep, need_desc=True):
    assert rep["recall"] >= MIN_RECALL and rep["precision"] >= MIN_RECALL, rep
    assert rep["pos_err_max"] <= POS_STRICT, rep
    assert rep["resp_rel_max"] <= RESP_RTOL, rep
    assert rep["size_equal"] and rep["octave_equal"], rep
    assert rep["angle_frac_ok"] >= ANGLE_FRAC, rep
    if need_desc:
        assert rep["desc_mean_bits"] <= DESC_MEAN_BITS and rep["desc_exact_frac"] >= DESC_EXACT_FRAC, rep
