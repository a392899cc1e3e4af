"""CPU oracle for the frame preprocessing  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of /root/reference/fingerprint.py:186-214 (``_preprocess_frames``): ``cv2.resize(..., INTER_AREA)`` to a
short side of 64 with the reference's int() truncation of the long side, centre crop, and (optionally) the float / 255 CHW
tensor. The resize itself belongs to a third-party dependency, **opencv-python** (reference pin: uv.lock), whose published
algorithm (imgproc/src/resize.cpp: computeResizeAreaTab + ResizeArea_Invoker, ResizeAreaFast_Invoker, ResizeAreaFastVec) is
restated here; it is pinned against the installed cv2 (4.13) by tests/golden/preprocess.npz (tests/golden/make_golden_preprocess.py)
and, where cv2 is importable, live in tests/test_preprocess.py. Frames with a side below 64 px are UP-scaled: cv::resize then
leaves the area code and runs its 8-bit bilinear kernels (HResizeLinear / VResizeLinear, 11-bit fixed-point coefficients) with
the INTER_AREA coefficient rule; restated in ``_resize_linear_area``.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

FRAME = 64


def target_size(h: int, w: int) -> Tuple[int, int]:
    """(new_h, new_w) of fingerprint.py:190-196."""
    if h < w:
        return FRAME, int(w * FRAME / h)
    return int(h * FRAME / w), FRAME


def _area_tab(ssize: int, dsize: int, scale: float) -> List[Tuple[int, int, np.float32]]:
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((sx1 - 1, dx, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((sx, dx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((sx2, dx, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _short(v: np.float32) -> int:
    """saturate_cast<short>(float): cvRound (half to even), clamped."""
    return max(-32768, min(32767, int(np.rint(np.float32(v)))))


def linear_area_tab(ssize: int, dsize: int) -> Tuple[List[int], List[Tuple[int, int]]]:
    """Source index and the two 11-bit coefficients per destination index: cv::resize's table loop with area_mode = true
    (sx = floor(dx * scale), fx = (dx + 1) - (sx + 1) * inv_scale, clipped to [0, 1) by dropping the integer part; at the
    last source sample the pair degenerates to (2048, 0))."""
    inv_scale = dsize / ssize
    scale = 1.0 / inv_scale
    ofs, coef = [], []
    for d in range(dsize):
        s = math.floor(d * scale)
        f = np.float32((d + 1) - (s + 1) * inv_scale)
        f = np.float32(0.0) if f <= 0 else np.float32(f - np.floor(f))
        if s >= ssize - 1:
            f, s = np.float32(0.0), ssize - 1
        ofs.append(s)
        coef.append((_short((np.float32(1.0) - f) * np.float32(2048)), _short(f * np.float32(2048))))
    return ofs, coef


def _resize_linear_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """The up-scaling branch of cv2.resize(..., INTER_AREA) on uint8: horizontal pass in int32 (S[sx] * a0 + S[sx + 1] * a1),
    vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 (VResizeLinear<uchar, int, short, ...>)."""
    sh, sw, cn = src.shape
    xo, xa = linear_area_tab(sw, dw)
    yo, yb = linear_area_tab(sh, dh)
    S = src.astype(np.int64)
    H = np.zeros((sh, dw, cn), np.int64)
    for d in range(dw):
        s0, s1 = xo[d], min(xo[d] + 1, sw - 1)
        H[:, d] = S[:, s0] * xa[d][0] + S[:, s1] * xa[d][1]
    out = np.zeros((dh, dw, cn), np.uint8)
    for d in range(dh):
        s0, s1 = yo[d], min(yo[d] + 1, sh - 1)
        b0, b1 = yb[d]
        v = (((b0 * (H[s0] >> 4)) >> 16) + ((b1 * (H[s1] >> 4)) >> 16) + 2) >> 2
        out[d] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 H x W x C."""
    sh, sw, cn = src.shape
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    if scale_x < 1.0 or scale_y < 1.0:
        return _resize_linear_area(src, dw, dh)
    isx, isy = int(round(scale_x)), int(round(scale_y))
    eps = np.finfo(np.float64).eps
    if abs(scale_x - isx) < eps and abs(scale_y - isy) < eps:          # integer factors: box sums
        blk = src[: dh * isy, : dw * isx].reshape(dh, isy, dw, isx, cn).astype(np.int64).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((blk + 2) >> 2).astype(np.uint8)
        v = blk.astype(np.float32) * np.float32(1.0 / (isx * isy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xtab, ytab = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)
    xs = np.array([t[0] for t in xtab])
    xd = np.array([t[1] for t in xtab])
    xa = np.array([t[2] for t in xtab], np.float32)
    dst = np.zeros((dh, dw, cn), np.uint8)
    prev, acc = -1, None
    for sy, dy, beta in ytab:
        S = src[sy].astype(np.float32)
        buf = np.zeros((dw, cn), np.float32)
        for k in range(len(xtab)):                                      # order matters: fp32, one entry at a time
            buf[xd[k]] = buf[xd[k]] + S[xs[k]] * xa[k]
        if dy != prev:
            if prev >= 0:
                dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            prev, acc = dy, beta * buf
        else:
            acc = acc + beta * buf
    dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return dst


def preprocess_frame(frame: np.ndarray) -> np.ndarray:
    """One frame of fingerprint.py:189-207: (H, W, 3) uint8 -> (64, 64, 3) uint8."""
    h, w = frame.shape[:2]
    nh, nw = target_size(h, w)
    r = resize_area(frame, nw, nh)
    sh, sw = (nh - FRAME) // 2, (nw - FRAME) // 2
    return r[sh : sh + FRAME, sw : sw + FRAME]


def preprocess_frames(frames) -> np.ndarray:
    """(T, 64, 64, 3) uint8; ``/ 255`` and the CHW permute (fingerprint.py:210-212) are left to the caller."""
    return np.stack([preprocess_frame(f) for f in frames])


# seeded frame sets for the golden file / the GPU parity test: (name, T, H, W)
PREPROCESS_CASES = [
    ("1080p", 2, 1080, 1920), ("480p", 2, 480, 854), ("portrait", 2, 640, 360), ("square_300", 2, 300, 300),
    ("int_3x3", 2, 192, 192), ("int_2x2", 2, 128, 128), ("int_2x2_wide", 1, 128, 200), ("int_4x3", 1, 192, 256),
    ("identity", 1, 64, 64), ("barely", 1, 65, 70), ("odd", 3, 211, 397), ("uhd", 1, 2160, 3840),
    # a side below 64 px: INTER_AREA up-scales through the fixed-point bilinear kernels
    ("tiny_48x50", 2, 48, 50), ("tiny_20x33", 1, 20, 33), ("tiny_portrait", 1, 90, 40), ("tiny_63x64", 1, 63, 64),
    ("tiny_7x5", 1, 7, 5), ("tiny_wide", 1, 16, 300), ("tiny_square", 1, 32, 32),
]


def make_frames(name: str, t: int, h: int, w: int) -> np.ndarray:
    rng = np.random.default_rng(abs(hash((h, w, t))) % (2**31) if False else (h * 10007 + w * 31 + t))
    base = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    # smooth gradient mixed in so that neighbouring pixels correlate like real frames (exercises the .5 rounding cases less
    # uniformly than white noise alone)
    gy, gx = np.mgrid[0:h, 0:w]
    grad = ((gy * 255 // max(h - 1, 1)) ^ (gx * 255 // max(w - 1, 1))).astype(np.uint8)[None, :, :, None]
    return np.where(rng.random((t, h, w, 1)) < 0.5, base, grad).astype(np.uint8)
