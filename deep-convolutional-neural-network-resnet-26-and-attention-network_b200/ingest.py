"""GPU tile ingest: the reference's per-tile `img_finalize` / `img_finalize_flat`
(RoiBuilder.py:193-210) on the cached 8-bit tiles of a slide, as one kernel launch per bag.

    ingest = TileIngest(roi_size=1200, resolution=224)          # update_resolution_and_buffer(resolution)
    bag_u8 = ingest(rois_u8_cuda, train=True)                    # [T,1200,1200,3] uint8 HWC -> [T,3,224,224] uint8
    out = classifier(bag_u8, label)                              # ToTensor + Normalize(.5,.5) happen in the stem's load

The resampling is Pillow's antialiased bilinear filter in its own 8-bit fixed-point arithmetic (the weights are
computed here on the host, in double precision, exactly as libImaging/Resample.c does); the random crop offsets and
flips are drawn from torch's global CPU generator in torchvision's own order, so a seeded run reproduces the
reference's augmentations bit for bit.  No CPU fallback: the tiles must be on a CUDA device."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2      # Pillow: 8bpc resampling weights in fixed point


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` for the triangle filter over a whole row of `in_size`
    pixels: (bounds int32 [out, 2] = (first input pixel, count), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                          # bilinear: filter support 1.0, stretched when shrinking
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)        # C cast: truncation (values >= -0.5)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    k = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((k + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where((arg < 1.0) & (k < xmax[:, None]), 1.0 - arg, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):                         # the C loop's summation order
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    fixed = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64),
                     (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return np.ascontiguousarray(bounds), np.ascontiguousarray(fixed)


def draw_augmentations(n_tiles: int, pad: int = 100):
    """Per tile, in torchvision's order: RandomCrop.get_params (top, left: torch.randint(0, 2 pad + 1)), then the
    horizontal and the vertical flip (torch.rand(1) < 0.5).  Returns (crops int32 [n,2], flips uint8 [n])."""
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


class TileIngest:
    """The transforms of one `RoiBuilder` (roi_size, resolution) on the GPU."""

    def __init__(self, roi_size: int, resolution: int, pad: int = 100):
        self.roi, self.side, self.pad = int(roi_size), int(resolution), int(pad)
        self.bounds_host, coef = pil_bilinear_coeffs(self.roi, self.side)
        self.ksize = int(coef.shape[1])
        self._coef_host = coef
        self._dev = {}

    def _tables(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.bounds_host).to(device), torch.from_numpy(self._coef_host).to(device))
        return self._dev[key]

    def __call__(self, rois: torch.Tensor, train: bool = True, crops=None, flips=None, out=None) -> torch.Tensor:
        """rois: uint8 CUDA tensor [T, roi, roi, 3].  train=True: pad + random crop + flips (`img_finalize`), drawn here
        unless `crops` (int [T,2] = top, left) / `flips` (uint8 [T]) are given; train=False: resize only
        (`img_finalize_flat`).  Returns uint8 [T, 3, side, side] on the same device."""
        if not rois.is_cuda:
            raise RuntimeError("TileIngest needs the tile cache on a CUDA device (no CPU fallback)")
        if rois.dtype != torch.uint8 or rois.dim() != 4 or rois.shape[1] != self.roi or rois.shape[2] != self.roi \
                or rois.shape[3] != 3:
            raise ValueError(f"expected uint8 tiles [T,{self.roi},{self.roi},3], got {tuple(rois.shape)} {rois.dtype}")
        rois = rois.contiguous()
        dev, T = rois.device, int(rois.shape[0])
        bounds, coef = self._tables(dev)
        cr = fl = None
        if train:
            if crops is None or flips is None:
                dc, df = draw_augmentations(T, self.pad)
                crops = dc if crops is None else crops
                flips = df if flips is None else flips
            cr = torch.as_tensor(np.asarray(crops), dtype=torch.int32).reshape(T, 2).contiguous().to(dev)
            fl = torch.as_tensor(np.asarray(flips), dtype=torch.uint8).reshape(T).contiguous().to(dev)
        if out is None:
            out = torch.empty((T, 3, self.side, self.side), dtype=torch.uint8, device=dev)
        P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(_lib.load().mil_ingest_tiles_u8(P(rois), T, self.roi, P(cr), self.pad, P(fl), self.side, P(bounds),
                                                       P(coef), self.ksize,
                                                       self.bounds_host.ctypes.data_as(C.c_void_p), P(out), st),
                       "mil_ingest_tiles_u8")
        return out
