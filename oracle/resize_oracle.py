"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy restatement of the frame resize the reference's callers do
before the model: `cv2.resize(img, dsize=(224, 224), interpolation=cv2.INTER_CUBIC)` on uint8 BGR frames
(run_inference.py:79-80, :91-92; dota.py:347-348).

The algorithm is OpenCV's (the reference calls it through the un-pinned `opencv-python` dependency, not vendored):
imgproc/src/resize.cpp, 8-bit path — per destination coordinate  fx = (d + 0.5) * scale - 0.5  in float32, the four
Catmull-Rom-like taps with A = -0.75 evaluated in float32 (`interpolateCubic`), converted to 11-bit fixed point
(`saturate_cast<short>(c * 2048)`, round half to even), source taps clamped at the borders, horizontal pass into int32,
vertical pass, `(v + 2^21) >> 22` and saturation.

PINNED against cv2 4.13 in the build container (oracle/make_golden.py::gen_resize, fixture tests/golden/resize_cubic.npz)
with two caveats that are properties of OpenCV, not of this restatement: (1) OpenCV's SIMD vertical pass works in
float32 and differs from its own scalar fixed-point arithmetic by 1 LSB on ~7e-6 of the pixels; (2) the pip build
dispatches 8-bit cubic resize to Intel IPP by default (`cv2.ipp.useIPP()`), whose result differs from OpenCV's own C++
path by 1 LSB on ~4 % of the pixels — the fixture is generated with `cv2.ipp.setUseIPP(False)` and records both rates.
"""
import numpy as np

COEF_BITS = 11  # INTER_RESIZE_COEF_BITS


def cubic_taps(n_dst, n_src):
    """(first source index - of the tap with weight c[1] -, int32 weights [n_dst, 4]) per destination coordinate."""
    scale = n_src / n_dst  # double, as `scale_x = 1. / inv_scale_x`
    ofs = np.zeros(n_dst, np.int32)
    w = np.zeros((n_dst, 4), np.int32)
    A = np.float32(-0.75)
    one = np.float32(1.0)
    for d in range(n_dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        x = np.float32(fx - np.float32(sx))
        c = np.zeros(4, np.float32)
        c[0] = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
        c[1] = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
        c[2] = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
        c[3] = one - c[0] - c[1] - c[2]
        ofs[d] = sx
        w[d] = np.rint(c * np.float32(1 << COEF_BITS)).astype(np.int32)
    return ofs, w


def resize_cubic_u8(img, dst_h, dst_w):
    """img uint8 [H, W, C] -> uint8 [dst_h, dst_w, C]."""
    sh, sw, _ = img.shape
    xo, xa = cubic_taps(dst_w, sw)
    yo, ya = cubic_taps(dst_h, sh)
    src = img.astype(np.int64)
    cols = np.clip(xo[:, None] + np.arange(-1, 3)[None, :], 0, sw - 1)       # [dst_w, 4]
    hpass = (src[:, cols, :] * xa[None, :, :, None]).sum(2)                   # [sh, dst_w, C] int
    rows = np.clip(yo[:, None] + np.arange(-1, 3)[None, :], 0, sh - 1)       # [dst_h, 4]
    v = (hpass[rows] * ya[:, :, None, None]).sum(1)                           # [dst_h, dst_w, C]
    out = (v + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS)
    return np.clip(out, 0, 255).astype(np.uint8)


def synthetic_frame(h, w, seed):
    """uint8 BGR frame with smooth structure plus noise (so that taps with negative weights overshoot at edges)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(xx / 17.0)[..., None] * np.cos(yy / 23.0)[..., None] * np.array([1.0, 0.7, -0.8])
    img = base + rng.integers(-60, 60, (h, w, 3))
    img[h // 3: h // 3 + 9, :, :] = 255   # hard edges: overshoot / saturation
    img[:, w // 2: w // 2 + 5, :] = 0
    return np.clip(img, 0, 255).astype(np.uint8)
