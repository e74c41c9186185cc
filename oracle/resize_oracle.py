"""CPU restatement of the image resize the reference's loader applies before ToTensor — TEST INFRASTRUCTURE ONLY.

`GeneralDataset.image_transform` (src/data_loader/GeneralDataset.py:38-59) is torchvision `transforms.Resize((S, S))` on a
PIL image, i.e. Pillow's `Image.resize(..., BILINEAR)`: a separable, antialiased triangle filter evaluated in 8-bit
fixed point (Pillow 12.2 `src/libImaging/Resample.c`: `precompute_coeffs`, `normalize_coeffs_8bpc`,
`ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`; horizontal pass first, uint8 intermediate).
Pillow is not part of /root/reference (requirements.txt pins it as a dependency), so the algorithm is restated here and
pinned bit-exactly against Pillow itself in this container (tests/test_resize.py, tests/golden/resize_u8.npz).
"""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: coefficients are int32 with 22 fractional bits


def bilinear_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0) over the full image box.
    Returns (xmin int32 [out], count int32 [out], kk int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)  # C cast: truncation toward zero
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(ksize, np.float64)
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            t = -t if t < 0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
        tot = 0.0
        for x in range(n):
            tot += w[x]
        for x in range(n):
            if tot != 0.0:
                w[x] /= tot
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8bpc pass along `axis`: out = clip8((2^21 + sum pixel * kk) >> 22)."""
    xmin, cnt, kk = bilinear_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), np.uint8)
    for xx in range(out_size):
        n = cnt[xx]
        acc = (src[..., xmin[xx]:xmin[xx] + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_bilinear_u8(img_hwc: np.ndarray, size: int) -> np.ndarray:
    """uint8 [H, W, C] → uint8 [size, size, C], bit-identical to PIL.Image.resize((size, size), BILINEAR)."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    h, w, _ = img_hwc.shape
    out = img_hwc
    if w != size:  # ImagingResample: horizontal pass first, skipped when the width already matches
        out = _pass(out, size, axis=1)
    if h != size:
        out = _pass(out, size, axis=0)
    return out
