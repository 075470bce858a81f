"""CPU oracle for the WIDE extractor parameterisation (alt_resnet.py) in front of the MIL head.  TEST INFRASTRUCTURE ONLY.

A plain PyTorch-CPU restatement of the reference's ``alt_resnet.py`` network used as tile feature extractor of
``gbm/model.py``'s ``Attention`` -- i.e. ``self.cnn = nn.DataParallel(alt_resnet.ResNet(alt_resnet.BasicBlock, layers,
num_classes=80))``.  It is the checker the CUDA path (csrc/mil_wide_*.cu, wide.py) is compared with; only ``tests/``,
``tools/`` and ``bench.py``'s CPU legs may import it.

Parity status: PINNED.  ``tests/golden/make_wide_golden.py`` imports the UNMODIFIED ``alt_resnet.py`` (its relative
``from .utils import load_state_dict_from_url`` satisfied by a stub package, oracle/ref_shim.py) together with the
unmodified ``gbm/model.py``, swaps the reference ``Attention``'s ``cnn`` for the reference's own alt ResNet, runs it on
the deterministic synthetic bags of ``oracle/synth.py`` and stores outputs + gradient digests in
``tests/golden/wide_*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against those vectors.

Reference lines followed:
  * conv3x3 / conv1x1 (no bias) ............ alt_resnet.py:24-32
  * BasicBlock ............................. alt_resnet.py:35-67
  * ResNet stem / layers / fc (with bias) .. alt_resnet.py:70-145
  * head ................................... oracle/mil_oracle.py (gbm/model.py:200-264)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from . import mil_oracle
from .mil_oracle import _RoundOperand, _RoundStored

WIDTHS = (64, 128, 256, 512)   # alt_resnet.py:87-90
RESNET18 = (2, 2, 2, 2)        # alt_resnet.py:157-165


def param_shapes(layers: Sequence[int] = RESNET18, widths: Sequence[int] = WIDTHS) -> "OrderedDict[str, tuple]":
    """State dict of Attention with cnn = DataParallel(alt_resnet.ResNet(BasicBlock, layers, num_classes=80)), in order."""
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    sh["weight_mask"] = (3,)
    sh["cnn.module.conv1.weight"] = (widths[0], 3, 7, 7)                  # alt_resnet.py:81 (bias=False)
    inpl = widths[0]
    for li, (w, nb) in enumerate(zip(widths, layers), start=1):
        for b in range(nb):
            cin = inpl if b == 0 else w
            p = f"cnn.module.layer{li}.{b}"
            sh[p + ".conv1.weight"] = (w, cin, 3, 3)
            sh[p + ".conv2.weight"] = (w, w, 3, 3)
            if b == 0 and (li > 1 or cin != w):                             # alt_resnet.py:113-116
                sh[p + ".downsample.0.weight"] = (w, cin, 1, 1)
        inpl = w
    sh["cnn.module.fc.weight"] = (80, widths[3])                          # alt_resnet.py:90
    sh["cnn.module.fc.bias"] = (80,)
    for k, v in mil_oracle.param_shapes().items():
        if not k.startswith("cnn.") and k != "weight_mask":
            sh[k] = v
    return sh


def init_params(seed: int = 0, layers: Sequence[int] = RESNET18, widths: Sequence[int] = WIDTHS,
                dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic weights with the init DISTRIBUTIONS of Attention.reset_params (gbm/model.py:161-181): the golden
    cases are generated from these (11 M parameters are not committed), so the generator is part of the fixture."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    slope = mil_oracle.SLOPE
    for name, shape in param_shapes(layers, widths).items():
        if name == "weight_mask":
            t = torch.full(shape, 0.25)
        elif name == "cnn.module.fc.bias":
            t = torch.randn(shape, generator=g) * 0.05      # non-zero so that the bias path is exercised
        elif name.endswith(".bias"):
            t = torch.zeros(shape)
        elif name == "context.bn.weight":
            t = torch.ones(shape)
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (math.sqrt(2.0 / (1 + slope ** 2)) / math.sqrt(fan_out))
        else:
            fan_in, fan_out = shape[1], shape[0]
            if name.startswith("attention"):
                std = (5.0 / 3.0) / math.sqrt(fan_in)
            elif "classifier" in name:
                std = math.sqrt(2.0 / (fan_in + fan_out))
            else:
                std = math.sqrt(2.0 / (1 + slope ** 2)) / math.sqrt(fan_in)
            t = torch.randn(shape, generator=g) * std
        out[name] = t.to(dtype)
    return out


def alt_resnet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, layers: Sequence[int] = RESNET18,
                       prefix: str = "cnn.module.", slope: float = 0.0, emulate_bf16: str = "") -> torch.Tensor:
    """x [N,3,S,S] -> features [N,80] (alt_resnet.py:125-139).  emulate_bf16 as in mil_oracle.resnet26_forward: "act"
    rounds every STORED activation map to bf16 (and the gradient flowing back through it), "act+w" also the conv
    weights and the input tiles (tensor-core operands)."""
    act = (lambda t: F.relu(t)) if slope == 0.0 else (lambda t: F.leaky_relu(t, slope))
    ra = _RoundStored.apply if emulate_bf16 else (lambda t: t)
    rw = _RoundOperand.apply if emulate_bf16 == "act+w" else (lambda t: t)
    y = F.conv2d(rw(x), rw(p[prefix + "conv1.weight"]), None, stride=2, padding=3)        # :127
    y = ra(F.max_pool2d(rw(act(y)), kernel_size=3, stride=2, padding=1))                   # :128-129
    for li in range(1, 5):
        for b in range(layers[li - 1]):
            q = f"{prefix}layer{li}.{b}"
            stride = 2 if (b == 0 and li > 1) else 1
            h = ra(act(F.conv2d(y, rw(p[q + ".conv1.weight"]), None, stride=stride, padding=1)))   # :55-56
            z = F.conv2d(h, rw(p[q + ".conv2.weight"]), None, stride=1, padding=1)                 # :58
            ident = y
            if q + ".downsample.0.weight" in p:
                ident = ra(F.conv2d(y, rw(p[q + ".downsample.0.weight"]), None, stride=2))         # :60-61
            y = ra(act(z + ident))                                                                  # :63-64
    pooled = torch.flatten(F.adaptive_avg_pool2d(y, (1, 1)), 1)                           # :136-137
    return pooled @ p[prefix + "fc.weight"].t() + p[prefix + "fc.bias"]                    # :138


def attention_forward(p, bag, Y, layers: Sequence[int] = RESNET18, class_weights=None, training: bool = False,
                      indices: Optional[torch.Tensor] = None, drop_mask: Optional[torch.Tensor] = None,
                      slope: float = 0.0, emulate_bf16: str = ""):
    """Attention.forward (gbm/model.py:189-264) with the alt ResNet as `cnn`."""
    x = bag.detach()
    if training:
        if indices is None:
            indices = mil_oracle.subsample_indices(bag.shape[0])
        x = x[indices]
    H = alt_resnet_forward(p, x, layers, slope=slope, emulate_bf16=emulate_bf16)
    return mil_oracle.head_forward(p, H, Y, class_weights, drop_mask if training else None)


def forward_backward(p, bag, Y, layers: Sequence[int] = RESNET18, class_weights=None, training: bool = False,
                     indices=None, drop_mask=None, slope: float = 0.0, emulate_bf16: str = ""):
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out = attention_forward(q, bag, Y, layers, class_weights, training, indices, drop_mask, slope, emulate_bf16)
    out["loss"].backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in q.items())
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads
