"""Pins oracle/mil_oracle.py against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py; reference gbm/model.py:189-264, nnBlocks.py:47-189)."""
import numpy as np
import pytest
import torch

from oracle import mil_oracle, ref_shim, synth
from tests.helpers import golden_cases, golden_weights, relerr

CASES = golden_cases()
TOL = 2e-5          # fp32 restatement vs fp32 reference: summation-order noise only


def _run_oracle(meta, rec):
    p = golden_weights()
    p["weight_mask"] = torch.tensor(meta["wm"])
    bag = torch.from_numpy(synth.make_bag(meta["n"], meta["side"], seed=meta.get("seed", 1)))
    cw = None if meta["cw"] is None else torch.tensor(meta["cw"])
    kw = {}
    if meta["training"]:
        idx = torch.from_numpy(rec["extra.indices"])
        kw = dict(training=True, indices=idx,
                  drop_mask=torch.from_numpy(synth.make_drop_mask(len(idx), seed=2)))
    return mil_oracle.forward_backward(p, bag, torch.tensor([meta["Y"]]), cw, **kw)


def test_golden_files_present():
    assert len(CASES) >= 6
    assert len(golden_weights()) == 65
    assert list(golden_weights().keys()) == list(mil_oracle.param_shapes().keys())
    for k, shp in mil_oracle.param_shapes().items():
        assert tuple(golden_weights()[k].shape) == shp


@pytest.mark.parametrize("meta,rec", CASES, ids=[c[0]["name"] for c in CASES])
def test_oracle_matches_reference_golden(meta, rec):
    out, grads = _run_oracle(meta, rec)
    for k in ("Aterm", "wROIs", "Bterm", "Mterm", "Fterm", "Aterm_mu", "Aterm_var", "loss", "l2", "KLD",
              "y_pred", "error"):
        assert tuple(out[k].shape) == tuple(rec[f"out.{k}"].shape), k
        assert relerr(out[k], rec[f"out.{k}"]) < TOL, k
    assert int(out["y_pred_hat"]) == int(rec["out.y_pred_hat"])
    for name, g in grads.items():
        dig = rec[f"gdigest.{name}"]
        scale = max(dig[1], 1e-12)
        assert abs(g.double().sum().item() - dig[0]) <= 1e-4 * scale + 2e-6, name
        assert abs(g.double().abs().sum().item() - dig[1]) <= 1e-4 * scale + 2e-6, name
        if f"grad.{name}" in rec and np.abs(rec[f"grad.{name}"]).max() > 1e-5:
            # (d loss / d buffer.classifier.bias is analytically 0: sum_k dM_k = 0 -> rounding noise only)
            assert relerr(g, rec[f"grad.{name}"]) < 5e-5, name


def test_single_tile_raises_like_reference():
    p = golden_weights()
    with pytest.raises(ValueError):
        mil_oracle.head_forward(p, torch.zeros(1, 80), torch.tensor([1]))


def test_synth_is_deterministic_and_diverse():
    a = synth.make_bag(4, 32, seed=1)
    b = synth.make_bag(4, 32, seed=1)
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert a.min() >= -1 and a.max() <= 1
    assert np.std(a.mean(axis=(2, 3))) > 0.2


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not on this box")
def test_oracle_matches_live_reference():
    """Where the reference checkout exists: fresh run of the real thing vs the restatement."""
    net = ref_shim.build_reference(seed=0).eval()
    p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    bag = torch.from_numpy(synth.make_bag(6, 80, seed=7))
    Y = torch.tensor([2])
    ref = net(bag, Y)
    ref["loss"].backward()
    out, grads = mil_oracle.forward_backward(p, bag, Y)
    for k in ("Aterm", "Mterm", "Fterm", "loss", "y_pred", "Aterm_var", "Aterm_mu", "KLD", "l2"):
        assert relerr(out[k], ref[k].detach()) < TOL, k
    for k, prm in net.named_parameters():
        if prm.grad.abs().max() > 1e-5:
            assert relerr(grads[k], prm.grad) < 5e-5, k
