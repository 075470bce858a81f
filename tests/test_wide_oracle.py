"""Pins oracle/wide_oracle.py (the alt_resnet.py parameterisation of the extractor) against the
golden vectors produced by the UNMODIFIED reference classes (tests/golden/make_wide_golden.py: gbm/model.py's Attention
with alt_resnet.py's ResNet as `cnn`), and the package's parameter holders against alt_resnet.py itself."""
import importlib

import numpy as np
import pytest
import torch

from oracle import ref_shim, wide_oracle
from tests.helpers import grad_sample, relerr, wide_case_inputs, wide_golden_cases

CASES = wide_golden_cases()
TOL = 2e-5
PKG = "deep-convolutional-neural-network-resnet-26-and-attention-network_b200"


def test_wide_golden_files_present():
    assert len(CASES) >= 4
    assert len(wide_oracle.param_shapes()) == 33


@pytest.mark.parametrize("meta,rec", CASES, ids=[c[0]["name"] for c in CASES])
def test_wide_oracle_matches_reference_golden(meta, rec):
    p, bag, Y, cw, kw = wide_case_inputs(meta, rec)
    out, grads = wide_oracle.forward_backward(p, bag, Y, meta["layers"], cw, **kw)
    for k in ("Aterm", "wROIs", "Bterm", "Mterm", "Fterm", "Aterm_mu", "Aterm_var", "loss", "l2", "KLD", "y_pred", "error"):
        assert tuple(out[k].shape) == tuple(rec[f"out.{k}"].shape), k
        assert relerr(out[k], rec[f"out.{k}"]) < TOL, k
    assert int(out["y_pred_hat"]) == int(rec["out.y_pred_hat"])
    for name, g in grads.items():
        dig = rec[f"gdigest.{name}"]
        scale = max(dig[1], 1e-12)
        assert abs(g.double().abs().sum().item() - dig[1]) <= 1e-4 * scale + 2e-6, name
        if f"grad.{name}" in rec and np.abs(rec[f"grad.{name}"]).max() > 1e-5:
            assert relerr(g, rec[f"grad.{name}"]) < 5e-5, name
        if f"gsample.{name}" in rec and np.abs(rec[f"gsample.{name}"]).max() > 1e-5:
            assert relerr(grad_sample(g), rec[f"gsample.{name}"]) < 5e-5, name


def test_wide_module_tree_matches_the_oracle_table():
    """CPU-only: construction, state-dict keys / shapes / order, and the library's own table (no compute call)."""
    mil = importlib.import_module(PKG)
    torch.manual_seed(0)
    net = mil.WideAttention(n_classes=3)
    sd = net.state_dict()
    shapes = wide_oracle.param_shapes()
    assert list(sd.keys()) == list(shapes.keys())
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == shp, k
    table = net._param_table                      # straight from libmil_b200.so (mil_wide_param_info)
    assert [t[0] for t in table] == list(shapes.keys())
    assert table[-1][2] + 1 == sum(int(np.prod(s)) for s in shapes.values())
    net34 = mil.WideAttention(n_classes=3, layers=(3, 4, 6, 3))
    assert [t[0] for t in net34._param_table] == list(wide_oracle.param_shapes((3, 4, 6, 3)).keys())
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 3, 64, 64))            # CPU tensor: no fallback


@pytest.mark.skipif(not (ref_shim.reference_available() and ref_shim.alt_resnet_available()),
                    reason="reference checkout not on this box")
def test_holder_reproduces_alt_resnet_init_and_oracle_matches_live_reference():
    mil = importlib.import_module(PKG)
    alt = ref_shim.load_reference_alt_resnet()
    torch.manual_seed(5)
    ref = alt.ResNet(alt.BasicBlock, [2, 2, 2, 2], num_classes=80)
    torch.manual_seed(5)
    mine = mil.AltResNet(layers=(2, 2, 2, 2), num_classes=80)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k          # same modules in the same order => the same RNG stream
    # the restatement against a fresh run of the real classes
    x = torch.randn(3, 3, 72, 72)
    p = {"cnn.module." + k: v for k, v in a.items()}
    assert relerr(wide_oracle.alt_resnet_forward(p, x), ref(x).detach()) < TOL
