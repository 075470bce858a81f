"""TEST INFRASTRUCTURE -- CPU restatement of the reference's per-tile finalisation (RoiBuilder.py:193-210):

    ToPILImage -> Pad(100) -> RandomCrop(roi) -> Resize(S) -> RandomHorizontalFlip -> RandomVerticalFlip
    -> ToTensor -> Normalize(.5, .5)                                   (`img_finalize`, training)
    ToPILImage -> Resize(S) -> ToTensor -> Normalize(.5, .5)           (`img_finalize_flat`, validation)

The arithmetic that matters is Pillow's antialiased bilinear resampling of 8-bit images (third-party dependency of the
reference: torchvision.transforms.Resize on a PIL image -> PIL.Image.resize(BILINEAR) -> libImaging/Resample.c,
ImagingResample with the 8bpc fixed-point path; Pillow is not pinned by the reference, the restatement is checked
against the Pillow of this image in tests/test_ingest.py):
  * precompute_coeffs: per output position a window [xmin, xmin + xmax) of triangle-filter weights, support =
    max(1, in/out), normalised to sum 1 in double precision;
  * normalize_coeffs_8bpc: weights to fixed point with PRECISION_BITS = 22, round half away from zero;
  * horizontal pass over every input row, then vertical pass, each accumulating in int32 from 1 << 21 and clipping
    (acc >> 22) to [0, 255]: the intermediate image is 8-bit.
Only tests/, tools/ and bench.py's CPU legs may import this module."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """(bounds int32 [out,2] = (xmin, count), coeffs int32 [out, ksize]) exactly as Pillow's Resample.c computes them
    for the whole-image box."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64),
                     (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    return bounds, fixed


def _pass(img, bounds, coef, axis):
    """one resampling pass of an 8-bit image [H, W, C] along `axis` (0 = vertical, 1 = horizontal)"""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o, (lo, cnt) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(cnt):
            acc += src[lo + k] * int(coef[o, k])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img: np.ndarray, size: int) -> np.ndarray:
    """PIL.Image.resize((size, size), BILINEAR) of an 8-bit HWC image (square or not); equal size: a copy"""
    h, w = img.shape[:2]
    if h == size and w == size:
        return img.copy()
    if w != size:
        bx, cx = pil_bilinear_coeffs(w, size)
        img = _pass(img, bx, cx, 1)
    if h != size:
        by, cy = pil_bilinear_coeffs(h, size)
        img = _pass(img, by, cy, 0)
    return img


def finalize_tile(roi: np.ndarray, size: int, crop=None, pad: int = 100, hflip=False, vflip=False) -> np.ndarray:
    """One tile of `img_finalize` (crop = (top, left) inside the padded image) or `img_finalize_flat` (crop None):
    uint8 HWC [R,R,3] -> uint8 CHW [3,size,size].  ToTensor + Normalize are left to the consumer:
    (v / 255 - 0.5) / 0.5, what the stem's 8-bit load applies."""
    img = roi
    if crop is not None:
        r = roi.shape[0]
        padded = np.zeros((r + 2 * pad, roi.shape[1] + 2 * pad, roi.shape[2]), dtype=np.uint8)    # Pad(100): fill 0
        padded[pad:pad + r, pad:pad + roi.shape[1]] = roi
        top, left = crop
        img = padded[top:top + r, left:left + roi.shape[1]]                                    # RandomCrop(roi_size)
    img = resize_u8(np.ascontiguousarray(img), size)
    if hflip:
        img = img[:, ::-1]
    if vflip:
        img = img[::-1]
    return np.ascontiguousarray(img.transpose(2, 0, 1))


def draw_params(n_tiles: int, roi: int, pad: int = 100):
    """The random draws of `img_finalize` for n_tiles tiles in torchvision's own order and from torch's global CPU
    generator: per tile RandomCrop.get_params (top, then left: torch.randint(0, 2 * pad + 1)), then the horizontal
    and the vertical flip (torch.rand(1) < 0.5 each).  Returns (crops int32 [n,2] (top, left), flips uint8 [n])."""
    import torch
    crops = np.zeros((n_tiles, 2), dtype=np.int32)
    flips = np.zeros(n_tiles, dtype=np.uint8)
    for t in range(n_tiles):
        i = int(torch.randint(0, 2 * pad + 1, size=(1,)).item())
        j = int(torch.randint(0, 2 * pad + 1, size=(1,)).item())
        hf = bool(torch.rand(1) < 0.5)
        vf = bool(torch.rand(1) < 0.5)
        crops[t] = (i, j)
        flips[t] = (1 if hf else 0) | (2 if vf else 0)
    return crops, flips
