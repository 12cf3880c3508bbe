"""CPU ORACLE (test infrastructure, not product code) — recognition crop preprocessing.

numpy restatement of what SuryaOCRProcessor does to one line crop (surya/common/surya/processor/__init__.py:140-230): scale_to_fit's
cv2.resize(INTER_LANCZOS4), _process_and_tile's cv2.resize(INTER_CUBIC) to a multiple of 28, _image_processor, merge-block-major tiles.
The resampling itself lives in a third-party dependency of the reference, OpenCV (opencv-python, 4.13.0 in this image); its resize for
float32 images is restated here from the published algorithm (modules/imgproc/src/resize.cpp: resizeGeneric_, interpolateLanczos4,
interpolateCubic) and PINNED against cv2 itself in tests/test_preproc_cpu.py:
  * destination index d samples source coordinate f = (d + 0.5) * scale - 0.5, scale = 1 / (dst / src); taps floor(f) - k/2 + 1 ...
    floor(f) + k/2 with clamped indices (border replicate); k = 8 (Lanczos4), 4 (cubic);
  * INTER_LANCZOS4 takes the generic path: float32 coordinate, weights from interpolateLanczos4;
  * INTER_CUBIC on float32 takes the IPP path in this build: coordinate and Keys weights (A = -0.75) in double (the non-IPP generic
    path uses a float32 coordinate and differs by up to 0.02 on the 0..255 scale for 2000-pixel-wide crops);
  * horizontal pass into float32 rows, then the vertical pass.
Float summation order inside OpenCV's SIMD kernels is not specified, so the pin is to float32 rounding (<= 5e-4 on the 0..255 scale).
Only tests/ import this module.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

IMAGE_MEAN = np.array((0.485, 0.456, 0.406), dtype=np.float32)
IMAGE_STD = np.array((0.229, 0.224, 0.225), dtype=np.float32)


def lanczos_weights(x: np.float32) -> np.ndarray:
    """interpolateLanczos4 (imgproc/src/resize.cpp)."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    c = np.zeros(8, np.float32)
    if x < np.finfo(np.float32).eps:
        c[3] = 1
        return c
    y0 = -(float(x) + 3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    total = np.float32(0)
    for i in range(8):
        y = -(float(x) + 3 - i) * math.pi * 0.25
        c[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        total = np.float32(total + c[i])
    return (c * (np.float32(1) / total)).astype(np.float32)


def cubic_weights(x: float) -> np.ndarray:
    """Keys cubic, A = -0.75 (interpolateCubic), evaluated in double, stored as float32."""
    A = -0.75
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    return np.array([c0, c1, c2, 1.0 - c0 - c1 - c2]).astype(np.float32)


def axis_table(src: int, dst: int, mode: str) -> Tuple[np.ndarray, np.ndarray]:
    k = 8 if mode == "lanczos" else 4
    scale = 1.0 / (dst / src)
    idx = np.empty((dst, k), np.int64)
    w = np.empty((dst, k), np.float32)
    for d in range(dst):
        if mode == "lanczos":
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(math.floor(f))
            w[d] = lanczos_weights(np.float32(f - np.float32(s)))
        else:
            f = (d + 0.5) * scale - 0.5
            s = math.floor(f)
            w[d] = cubic_weights(f - s)
        idx[d] = np.clip(np.arange(s - k // 2 + 1, s + k // 2 + 1), 0, src - 1)
    return idx, w


def resize(img: np.ndarray, dw: int, dh: int, mode: str) -> np.ndarray:
    """cv2.resize(img float32 HWC, (dw, dh), INTER_LANCZOS4 | INTER_CUBIC) — separable, float32 accumulation tap by tap."""
    img = np.asarray(img, np.float32)
    h, w = img.shape[:2]
    xi, xw = axis_table(w, dw, mode)
    yi, yw = axis_table(h, dh, mode)
    rows = np.zeros((h, dw, img.shape[2]), np.float32)
    for j in range(xi.shape[1]):
        rows = rows + img[:, xi[:, j], :] * xw[None, :, j, None]
    out = np.zeros((dh, dw, img.shape[2]), np.float32)
    for j in range(yi.shape[1]):
        out = out + rows[yi[:, j]] * yw[:, j, None, None]
    return out


def fit_size(h: int, w: int, max_size=(1024, 256), min_size=(168, 168)) -> Tuple[int, int]:
    """Output (h, w) of scale_to_fit (processor/__init__.py:148-174)."""
    cur, mx, mn = w * h, max_size[0] * max_size[1], min_size[0] * min_size[1]
    if cur > mx:
        s = (mx / cur) ** 0.5
        return math.floor(h * s), math.floor(w * s)
    if cur < mn:
        s = (mn / cur) ** 0.5
        return math.ceil(h * s), math.ceil(w * s)
    return h, w


def process_crop(crop_u8: np.ndarray, patch: int = 14, merge: int = 2):
    """uint8 crop -> (tiles fp32 [gh * gw, 3 * patch * patch], (1, gh, gw)): scale_to_fit + _process_and_tile."""
    img = np.asarray(crop_u8, np.float32)
    h, w = img.shape[:2]
    nh, nw = fit_size(h, w)
    if (nh, nw) != (h, w):
        img = resize(img, nw, nh, "lanczos")
    factor = patch * merge
    hb, wb = math.ceil(nh / factor) * factor, math.ceil(nw / factor) * factor
    if (hb, wb) != (nh, nw):
        img = resize(img, wb, hb, "cubic")
    img = (img.astype(np.float64) * (1 / 255.0)).astype(np.float32)
    img = (img - IMAGE_MEAN) / IMAGE_STD
    gh, gw = hb // patch, wb // patch
    x = img.transpose(2, 0, 1).reshape(3, gh // merge, merge, patch, gw // merge, merge, patch).transpose(1, 4, 2, 5, 0, 3, 6)
    return np.ascontiguousarray(x).reshape(gh * gw, 3 * patch * patch), (1, gh, gw)
