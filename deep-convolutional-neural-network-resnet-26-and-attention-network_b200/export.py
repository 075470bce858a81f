"""Attention-map export: what the reference's `write_map` / `visualize` do with the model
outputs (gbm/classify.py:207-225, gbm/classify_combined.py:151-165): min-max scale an attention tensor and write one
`x y weight` line per tile into `prediction-AGMIL-*.{name}.dla` files that the heat-map viewer reads."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def minmax_normalize(t: torch.Tensor):
    """(t - t.min()) / (t.max() - t.min()) over the whole tensor, on the device (one reduction + one scaling pass).
    Returns (scaled tensor, (min, max) as a device tensor)."""
    if not t.is_cuda:
        raise RuntimeError("minmax_normalize: CUDA tensor expected (no CPU fallback)")
    x = t.detach().to(torch.float32).contiguous()
    out = torch.empty_like(x)
    mm = torch.empty(2, dtype=torch.float32, device=x.device)
    P = lambda a: C.c_void_p(a.data_ptr())
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mil_minmax_normalize(P(x), P(out), x.numel(), P(mm),
                                                    C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "mil_minmax_normalize")
    return out, mm


def top_tiles(attention: torch.Tensor, k: int = 8) -> torch.Tensor:
    """Indices of the k most attended tiles of every attention map ([K, N] -> [K, k]), most attended first."""
    return torch.topk(attention.detach(), min(k, attention.shape[1]), dim=1).indices


def write_dla(path: str, raster, weights) -> None:
    """One `x y weight` line per tile; raster[i] = (row, col) as in the reference, written as `col row weight`
    (gbm/classify.py:211-212)."""
    raster = np.asarray(raster)
    weights = np.asarray(weights, dtype=np.float64).reshape(-1)
    if raster.shape[0] != weights.shape[0]:
        raise ValueError(f"write_dla: {raster.shape[0]} coordinates for {weights.shape[0]} weights")
    with open(path, "w") as f:
        for (r, c), w in zip(raster[:, :2], weights):
            f.write(f"{c} {r} {w}\n")


def export_attention_maps(output: dict, raster, out_dir: str, name: str) -> list:
    """Writes, for the K = 3 attention maps of one slide,
        prediction-AGMIL-ATTN{k}.{name}.dla   min-max scaled `Aterm[k]`   (plt.Normalize()(attn), classify.py:209)
        prediction-AGMIL-ACTF{k}.{name}.dla   `wROIs[k]` = attention x instance code, unscaled (classify.py:214-224)
    and returns the paths.  One device->host copy per tensor, after the scaling ran on the device."""
    os.makedirs(out_dir, exist_ok=True)
    scaled, _ = minmax_normalize(output["Aterm"])
    scaled = scaled.cpu().numpy()
    act = output["wROIs"].detach().float().cpu().numpy()
    paths = []
    for k in range(scaled.shape[0]):
        for tag, arr in (("ATTN", scaled), ("ACTF", act)):
            p = os.path.join(out_dir, f"prediction-AGMIL-{tag}{k + 1}.{name}.dla")
            write_dla(p, raster, arr[k])
            paths.append(p)
    return paths
