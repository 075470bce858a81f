"""Shared helpers for the parity tests (tests/ may use oracle/)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_weights():
    z = np.load(os.path.join(GOLDEN, "weights.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))):
        z = np.load(f)
        meta = json.loads(bytes(z["meta"]).decode())
        out.append((meta, {k: z[k] for k in z.files if k != "meta"}))
    return out


def relerr(a, b, atol=1e-7):
    """Normwise relative error max|a-b| / max|b| (SURVEY.md section 8d parity gate).
    Differences below `atol` count as zero (quantities that are analytically ~0, e.g. Aterm_mu of a
    2-tile bag or d loss / d buffer.classifier.bias, are pure rounding noise)."""
    a = torch.as_tensor(a, dtype=torch.float64).flatten().cpu()
    b = torch.as_tensor(b, dtype=torch.float64).flatten().cpu()
    den = b.abs().max().item()
    diff = (a - b).abs().max().item() if a.numel() else 0.0
    if diff <= atol:
        return 0.0
    return diff / (den if den > 0 else 1.0)
