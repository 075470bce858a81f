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


def perturbed_weights(seed=0, conv_bias=True):
    """The reference's init moved off its symmetric point.  At the init itself the three attention maps are
    identical (weight_mask = .25 each), every bias is zero and sum_k dLoss/dM_k = 0, so the instance-branch gradient
    and -- through the bag-wide BatchNorm1d, whose backward removes the bag mean -- every bias gradient of the network
    are sums that cancel almost completely: their VALUE is rounding noise in any precision (the reference's own fp32
    gradients sit `gnoise` away from fp64 there).  Kernel cross-checks therefore run on weights where the gradients
    are real numbers: distinct mask logits, non-zero biases, BN affine away from (1, 0), head weights rescaled."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = golden_weights()
    for k, v in sd.items():
        if k == "weight_mask":
            sd[k] = torch.tensor([-1.0, 0.25, 1.0])
        elif k == "context.bn.weight":
            sd[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        elif k == "context.bn.bias":
            sd[k] = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith(".bias"):
            if conv_bias or not k.startswith("cnn."):
                sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        elif k.startswith(("attention.", "buffer.")):
            sd[k] = v * (1.0 + 0.3 * torch.randn(v.shape, generator=g))
    return sd


def wide_golden_cases():
    """Golden cases of the wide (alt_resnet.py) parameterisation: tests/golden/make_wide_golden.py."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "wide_*.npz"))):
        z = np.load(f)
        meta = json.loads(bytes(z["meta"]).decode())
        out.append((meta, {k: z[k] for k in z.files if k != "meta"}))
    return out


def wide_case_inputs(meta, rec):
    """(params, bag, Y, class weights, training kwargs) of a wide golden case, regenerated deterministically."""
    from oracle import synth, wide_oracle
    p = wide_oracle.init_params(meta["wseed"], meta["layers"])
    p["weight_mask"] = torch.tensor(meta["wm"])
    bag = torch.from_numpy(synth.make_bag(meta["n"], meta["side"], seed=meta.get("seed", 1)))
    cw = None if meta["cw"] is None else torch.tensor(meta["cw"])
    kw = {}
    if meta["training"]:
        idx = torch.from_numpy(rec["extra.indices"])
        kw = dict(training=True, indices=idx, drop_mask=torch.from_numpy(synth.make_drop_mask(len(idx), seed=2)))
    return p, bag, torch.tensor([meta["Y"]]), cw, kw


def grad_sample(t, n=2048):
    """The strided sample make_wide_golden.py stores for the large gradient tensors."""
    f = torch.as_tensor(t).flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n]
