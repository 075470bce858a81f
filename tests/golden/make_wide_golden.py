"""Generate the golden vectors of the WIDE extractor parameterisation from the UNMODIFIED reference classes (run in the
build container only).

    python tests/golden/make_wide_golden.py

The reference's `Attention` (gbm/model.py:114-264) with its `cnn` swapped for the reference's own
`alt_resnet.ResNet(alt_resnet.BasicBlock, layers, num_classes=80)` (alt_resnet.py:70-145, imported through the stub
package of oracle/ref_shim.py) runs on the deterministic bags of oracle/synth.py with the deterministic weights of
oracle/wide_oracle.init_params (11 M parameters are not committed: the generator is part of the fixture).  Stored per
case: the 13 outputs, (sum, abs-sum, l2) digests of all gradients, the small gradients in full and a strided sample of
the large ones.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim, synth, wide_oracle  # noqa: E402
from tests.golden.make_golden import _InjectedDropout  # noqa: E402

CASES = [
    dict(name="eval_6x64_r18", layers=[2, 2, 2, 2], n=6, side=64, wm=[0.25, -0.5, -1.0], Y=1, cw=None, training=False, wseed=3),
    dict(name="eval_4x96_l1111_cw", layers=[1, 1, 1, 1], n=4, side=96, wm=[-1.0, 0.25, 1.0], Y=2, cw=[0.5, 1.0, 2.0],
         training=False, wseed=4),
    dict(name="eval_3x100_odd_r18", layers=[2, 2, 2, 2], n=3, side=100, wm=[0.25, 0.25, 0.25], Y=0, cw=None, training=False, wseed=5),
    dict(name="train_20x64_injected_r18", layers=[2, 2, 2, 2], n=20, side=64, wm=[0.25, -1.0, 0.25], Y=1, cw=None,
         training=True, wseed=3),
]
SAMPLE = 2048


def sample(t: torch.Tensor) -> torch.Tensor:
    f = t.flatten()
    step = max(1, f.numel() // SAMPLE)
    return f[::step][:SAMPLE]


def run_case(case):
    net = ref_shim.build_reference_wide(case["layers"], seed=0)
    params = wide_oracle.init_params(case["wseed"], case["layers"])
    with torch.no_grad():
        params["weight_mask"] = torch.tensor(case["wm"])
    missing = net.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(net.state_dict().keys()) == list(params.keys()), "state-dict order differs from wide_oracle.param_shapes"
    bag = torch.from_numpy(synth.make_bag(case["n"], case["side"], seed=case.get("seed", 1)))
    Y = torch.tensor([case["Y"]])
    net.loss.weight = None if case["cw"] is None else torch.tensor(case["cw"])
    extra = {}
    if case["training"]:
        net.train()
        torch.manual_seed(2)
        idx = torch.randperm(case["n"])[: int(case["n"] * 0.2)]      # replay of gbm/model.py:193
        mask = torch.from_numpy(synth.make_drop_mask(len(idx), seed=2))
        net.context.do = _InjectedDropout(mask)
        torch.manual_seed(2)
        extra["indices"] = idx.numpy()
    else:
        net.eval()
    out = net(bag, Y)
    out["loss"].backward()
    rec = {f"out.{k}": v.detach().numpy() for k, v in out.items()}
    for k, prm in net.named_parameters():
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        rec[f"gdigest.{k}"] = np.array([g.sum().item(), g.abs().sum().item(), g.norm().item()], dtype=np.float64)
        if g.numel() <= 3200:
            rec[f"grad.{k}"] = g.numpy()
        else:
            rec[f"gsample.{k}"] = sample(g).numpy()
    rec.update({f"extra.{k}": v for k, v in extra.items()})
    rec["meta"] = np.frombuffer(json.dumps(case).encode(), dtype=np.uint8)
    return rec


def main():
    assert ref_shim.reference_available() and ref_shim.alt_resnet_available(), "needs the reference checkout"
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    for case in CASES:
        rec = run_case(case)
        np.savez_compressed(os.path.join(HERE, f"wide_{case['name']}.npz"), **rec)
        print(case["name"], "loss", float(rec["out.loss"]), "y_pred", rec["out.y_pred"].ravel())


if __name__ == "__main__":
    main()
