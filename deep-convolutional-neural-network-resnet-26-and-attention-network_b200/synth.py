"""Deterministic synthetic bags (numpy only, version-stable): test and benchmark inputs.

SURVEY.md section 8d: tiles must be structurally diverse -- i.i.d. noise tiles make the bag-wide
BatchNorm1d of the reference (gbm/model.py:105-109) ill conditioned.  Each tile gets its own
per-channel mean, contrast and a low-frequency field plus a little fine texture, clamped to the
[-1, 1] range the reference's transforms produce (RoiBuilder.py:201-202).

Everything is derived from closed-form sinusoids with parameters drawn from PCG64.random(), so the
same (seed, n, s) gives the same bag on any machine; the goldens under tests/golden/ depend on it.
"""
from __future__ import annotations

import numpy as np


def make_bag(n_tiles: int, side: int, seed: int = 1) -> np.ndarray:
    """fp32 NCHW [n_tiles, 3, side, side] in [-1, 1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random((n_tiles, 3, 16))                     # per (tile, channel) parameters
    yy, xx = np.meshgrid(np.arange(side, dtype=np.float64) / side,
                         np.arange(side, dtype=np.float64) / side, indexing="ij")
    out = np.empty((n_tiles, 3, side, side), dtype=np.float32)
    for n in range(n_tiles):
        for c in range(3):
            q = u[n, c]
            mean = -0.8 + 1.6 * q[0]
            contrast = 0.1 + 0.9 * q[1]
            f1x, f1y = 0.5 + 3.5 * q[2], 0.5 + 3.5 * q[3]
            f2x, f2y = 0.5 + 3.5 * q[4], 0.5 + 3.5 * q[5]
            field = (np.sin(2 * np.pi * (f1x * xx + f1y * yy + q[6]))
                     + np.sin(2 * np.pi * (f2x * xx - f2y * yy + q[7]))) * 0.5
            hx, hy = 17.0 + 40.0 * q[8], 13.0 + 40.0 * q[9]
            fine = np.sin(2 * np.pi * (hx * xx + q[10])) * np.sin(2 * np.pi * (hy * yy + q[11]))
            img = mean + contrast * (0.6 * field + 0.15 * fine)
            out[n, c] = np.clip(img, -1.0, 1.0).astype(np.float32)
    return out


def make_drop_mask(n_tiles: int, seed: int = 2, p: float = 0.25) -> np.ndarray:
    """{0,1} keep mask [n_tiles, 80] for an injected, reproducible Dropout(0.25)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.random((n_tiles, 80)) >= p).astype(np.float32)
