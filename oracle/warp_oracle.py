"""Oracle for `warp_image_perspective` — TEST INFRASTRUCTURE ONLY.

Restates `cv::warpPerspective(src, dst, M, size, INTER_LINEAR, BORDER_CONSTANT, Scalar(1,1,1,1))` as the
reference calls it (homographier/src/homographier/mod.rs:271-300) for 8-bit images.  The arithmetic
lives in OpenCV's imgproc (`imgwarp.cpp` WarpPerspectiveInvoker + remapBilinear; not vendored in the
reference): M is inverted in f64, source coordinates are computed per 64 x 16 block in f64
(X0 + M0*x1 with X0 taken at the block's first column), scaled by 32 / W and rounded to a 1/32-pixel
grid (cvRound), and the four neighbours are blended with 15-bit fixed-point weights; neighbours outside
the source take the border value.  Pinned against cv2 4.13.0 in tests/golden/warp_golden.npz.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15


def invert3(M):
    """cv::invert for a 3 x 3 f64 matrix (closed form, core/src/lapack.cpp)"""
    m = np.asarray(M, dtype=np.float64).reshape(3, 3)
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if d == 0:
        return None
    d = 1.0 / d
    t = np.empty(9)
    t[0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[3] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[4] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[5] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[6] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[7] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[8] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t.reshape(3, 3)


def block_size(width, height):
    bh0 = min(16, height)
    bw0 = min(1024 // bh0, width)
    bh0 = min(1024 // bw0, height)
    return bw0, bh0


def warp_perspective(src, M, out_w=None, out_h=None, border_value=1):
    """src: H x W (x C) u8.  Returns the warped image (out_h x out_w (x C)) u8."""
    src = np.asarray(src)
    assert src.dtype == np.uint8
    sq = src.ndim == 2
    s = src[..., None] if sq else src
    sh, sw, ch = s.shape
    out_w = out_w or sw
    out_h = out_h or sh
    Mi = invert3(M)
    if Mi is None:
        Mi = np.zeros((3, 3))
    m = Mi.ravel()
    bw, bh = block_size(out_w, out_h)
    xs = np.arange(out_w)
    x_blk = (xs // bw) * bw                     # first column of the block
    x1 = (xs - x_blk).astype(np.float64)
    ys = np.arange(out_h, dtype=np.float64)[:, None]
    xb = x_blk.astype(np.float64)[None, :]
    X0 = (m[0] * xb + m[1] * ys) + m[2]
    Y0 = (m[3] * xb + m[4] * ys) + m[5]
    W0 = (m[6] * xb + m[7] * ys) + m[8]
    W = W0 + m[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
    fX = np.clip((X0 + m[0] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + m[3] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    ax = X & (INTER_TAB_SIZE - 1)
    ay = Y & (INTER_TAB_SIZE - 1)
    w = [(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32]
    out = np.empty((out_h, out_w, ch), np.uint8)
    for c in range(ch):
        acc = np.zeros((out_h, out_w), np.int64)
        for k, (dx, dy) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
            px, py = sx + dx, sy + dy
            inside = (px >= 0) & (px < sw) & (py >= 0) & (py < sh)
            v = np.where(inside, s[np.clip(py, 0, sh - 1), np.clip(px, 0, sw - 1), c].astype(np.int64), border_value)
            acc += v * w[k]
        out[..., c] = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)
    return out[..., 0] if sq else out
