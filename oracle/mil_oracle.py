"""CPU oracle for the ResNet-26 + attention-MIL hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-CPU (fp32 or fp64) *restatement* of the algorithm in the
reference's ``gbm/model.py`` and ``nnBlocks.py``.  It is the checker the CUDA path is compared
with; it is never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the UNMODIFIED reference source
(under the three-piece shim in ``oracle/ref_shim.py``) in the build container, runs it on the
deterministic synthetic bags of ``oracle/synth.py`` and stores its outputs + gradients in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against those
vectors.  (The reference itself ships no tests or golden vectors: SURVEY.md section 4.)

Reference lines followed (paths relative to the reference checkout):
  * ResNet stem / layers / tail ............ gbm/model.py:14-61
  * BasicResBlock .......................... nnBlocks.py:157-189
  * ContextLayer (bag BatchNorm1d, dropout). gbm/model.py:89-111
  * Attention.forward ...................... gbm/model.py:189-264
  * smooth_one_hot / CE with probabilities . nnBlocks.py:47-138
  * parameter init ......................... gbm/model.py:161-181
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SLOPE = 0.1            # LeakyReLU slope everywhere (gbm/model.py:25, nnBlocks.py:171)
WIDTHS = (20, 40, 60, 80)   # gbm/model.py:27-30
BLOCKS = (3, 3, 3, 3)       # gbm/model.py:133
L, D, K, O = 80, 40, 3, 1   # gbm/model.py:120-123
BN_EPS = 1e-5               # nn.BatchNorm1d default (gbm/model.py:105)
DROP_P = 0.25               # gbm/model.py:107
SMOOTHING = 0.25            # gbm/model.py:128
SUBSAMPLE = 0.2             # gbm/model.py:193


def param_shapes() -> "OrderedDict[str, tuple]":
    """The 65 tensors of the reference state dict, in its own order."""
    sh: "OrderedDict[str, tuple]" = OrderedDict()
    sh["weight_mask"] = (3,)
    sh["cnn.module.conv1.weight"] = (20, 3, 7, 7)
    sh["cnn.module.conv1.bias"] = (20,)
    inpl = 20
    for li, (w, nb) in enumerate(zip(WIDTHS, BLOCKS), start=1):
        for b in range(nb):
            cin = inpl if b == 0 else w
            p = f"cnn.module.layer{li}.{b}"
            sh[p + ".conv1.weight"] = (w, cin, 3, 3)
            sh[p + ".conv1.bias"] = (w,)
            sh[p + ".conv2.weight"] = (w, w, 3, 3)
            sh[p + ".conv2.bias"] = (w,)
            if b == 0 and li > 1:
                sh[p + ".downsample.0.weight"] = (w, cin, 1, 1)
        inpl = w
    sh["cnn.module.fc.weight"] = (80, 80)
    sh["context.bn.weight"] = (80,)
    sh["context.bn.bias"] = (80,)
    sh["attention.lin1.weight"] = (40, 80)
    sh["attention.lin1.bias"] = (40,)
    sh["attention.lin2.weight"] = (3, 40)
    sh["attention.lin2.bias"] = (3,)
    sh["buffer.lin1.weight"] = (40, 80)
    sh["buffer.lin1.bias"] = (40,)
    sh["buffer.classifier.weight"] = (1, 40)
    sh["buffer.classifier.bias"] = (1,)
    return sh


def init_params(seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Same init *distributions* as Attention.reset_params (gbm/model.py:161-181).

    (Not the same random stream as the reference constructor; golden tests use the reference's
    own weights stored in tests/golden/weights.npz.)
    """
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_shapes().items():
        if name == "weight_mask":
            t = torch.full(shape, 0.25)
        elif name.endswith(".bias"):
            t = torch.zeros(shape)
        elif name == "context.bn.weight":
            t = torch.ones(shape)
        elif len(shape) == 4:          # conv: kaiming normal, fan_out, leaky_relu(0.1)
            fan_out = shape[0] * shape[2] * shape[3]
            std = math.sqrt(2.0 / (1 + SLOPE ** 2)) / math.sqrt(fan_out)
            t = torch.randn(shape, generator=g) * std
        else:                          # linear
            fan_in, fan_out = shape[1], shape[0]
            if name.startswith("attention"):
                std = (5.0 / 3.0) / math.sqrt(fan_in)            # kaiming, tanh gain
            elif "classifier" in name:
                std = math.sqrt(2.0 / (fan_in + fan_out))         # xavier normal
            else:
                std = math.sqrt(2.0 / (1 + SLOPE ** 2)) / math.sqrt(fan_in)
            t = torch.randn(shape, generator=g) * std
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------------
# feature extractor (gbm/model.py:50-61, nnBlocks.py:175-189)
# --------------------------------------------------------------------------------------------
def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).to(t.dtype)


class _RoundOperand(torch.autograd.Function):
    """bf16 rounding of a tensor-core OPERAND that has an fp32 master (conv weights, the input tiles, the conv map
    the pool compares): rounded in the forward pass, gradient passed through unchanged (the CUDA path accumulates
    weight gradients in fp32)."""

    @staticmethod
    def forward(ctx, t):
        return _bf16(t)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundStored(torch.autograd.Function):
    """bf16 rounding of a STORED activation map: the CUDA bf16 mode also stores the gradient maps that flow back
    through these points in bf16, so the gradient is rounded on the way back as well."""

    @staticmethod
    def forward(ctx, t):
        return _bf16(t)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


def resnet26_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "cnn.module.",
                     taps: Optional[dict] = None, emulate_bf16: str = "") -> torch.Tensor:
    """x [N,3,S,S] -> H [N,80].  `taps` (optional dict) receives every block output.

    emulate_bf16: "" = the reference arithmetic.  "act" = additionally round every STORED activation
    (stem output, each block's inner activation and output) to bf16, which is where the CUDA bf16 mode
    rounds; "act+w" = also round the 3x3 / 1x1 conv weights to bf16 (tensor-core operands).  Used by the
    tests to separate kernel defects (must match the emulation tightly) from the conditioning of the
    head with respect to bf16 features (the distance between the emulation and the fp32 reference)."""
    lr = lambda t: F.leaky_relu(t, SLOPE)
    ra = _RoundStored.apply if emulate_bf16 else (lambda t: t)
    rw = _RoundOperand.apply if emulate_bf16 == "act+w" else (lambda t: t)
    # "act+w": the stem runs on the tensor cores too -> bf16 input and conv1 weights, bf16 conv map before the pool
    y = F.conv2d(rw(x), rw(p[prefix + "conv1.weight"]), p[prefix + "conv1.bias"], stride=2, padding=3)
    y = ra(F.max_pool2d(rw(lr(y)), kernel_size=3, stride=2, padding=1))
    if taps is not None:
        taps["stem"] = y
    for li in range(1, 5):
        for b in range(3):
            q = f"{prefix}layer{li}.{b}"
            stride = 2 if (b == 0 and li > 1) else 1
            h = ra(lr(F.conv2d(y, rw(p[q + ".conv1.weight"]), p[q + ".conv1.bias"], stride=stride, padding=1)))
            z = F.conv2d(h, rw(p[q + ".conv2.weight"]), p[q + ".conv2.bias"], stride=1, padding=1)
            ident = y
            if q + ".downsample.0.weight" in p:
                ident = ra(F.conv2d(y, rw(p[q + ".downsample.0.weight"]), None, stride=2))
            y = ra(lr(z + ident))
            if taps is not None:
                taps[f"layer{li}.{b}.y1"] = h
                taps[f"layer{li}.{b}"] = y
    pooled = y.mean(dim=(2, 3))                              # AdaptiveAvgPool2d((1,1)) + flatten
    return pooled @ p[prefix + "fc.weight"].t()              # fc has no bias (gbm/model.py:32)


# --------------------------------------------------------------------------------------------
# MIL head (gbm/model.py:200-264), SURVEY.md appendix A
# --------------------------------------------------------------------------------------------
def smooth_target(Y: torch.Tensor, classes: int = 3, smoothing: float = SMOOTHING, dtype=torch.float32):
    """nnBlocks.py:71-85."""
    t = torch.full((Y.shape[0], classes), smoothing / (classes - 1), dtype=dtype)
    t.scatter_(1, Y.long().view(-1, 1), 1.0 - smoothing)
    return t


def head_forward(p: Dict[str, torch.Tensor], H: torch.Tensor, Y: torch.Tensor,
                 class_weights: Optional[torch.Tensor] = None,
                 drop_mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """H [N,80] -> the reference's 13-key output dict (nothing detached here).

    drop_mask: None => no dropout (eval).  Otherwise a {0,1} keep mask [N,80]; kept values are
    scaled by 1/(1-p) exactly as nn.Dropout does in training mode (gbm/model.py:107,110).
    """
    N = H.shape[0]
    if N < 2:
        # nn.BatchNorm1d(track_running_stats=False) refuses a single row in train AND eval.
        raise ValueError("Expected more than 1 value per channel when training, got input size "
                         f"{tuple(H.shape)}")
    KLD = 0.5 * H.pow(2).mean()                                           # :201
    mu = H.mean(dim=0)
    var = (H - mu).pow(2).mean(dim=0)                                     # biased, batch stats always
    xhat = (H - mu) / torch.sqrt(var + BN_EPS)
    Hz = xhat * p["context.bn.weight"] + p["context.bn.bias"]             # :109
    Hm = F.leaky_relu(H, SLOPE)                                           # :110
    if drop_mask is not None:
        Hm = Hm * drop_mask.to(H.dtype) / (1.0 - DROP_P)
    raw = torch.tanh(Hz @ p["attention.lin1.weight"].t() + p["attention.lin1.bias"]) \
        @ p["attention.lin2.weight"].t() + p["attention.lin2.bias"]        # :209  [N,3]
    act = F.softplus(raw)                                                 # :211
    w = p["weight_mask"]
    g = torch.sigmoid(-10.0 * w) * act + torch.sigmoid(10.0 * w)          # :212
    S = g.abs().sum(dim=0).clamp_min(1e-12)                               # F.normalize(p=1, dim=0)
    A = (g / S).t()                                                       # :213-214  [3,N]
    rn = raw / raw.pow(2).sum(dim=0).sqrt().clamp_min(1e-12)              # :216
    off = 1.0 - torch.eye(3, dtype=H.dtype)
    Aterm_var = ((rn.t() @ rn) * off).mean()                              # :218
    Aterm_mu = 0.5 * raw.mean(dim=0).pow(2).sum()                         # :219
    U = F.leaky_relu(Hm @ p["buffer.lin1.weight"].t() + p["buffer.lin1.bias"], SLOPE)
    B = U @ p["buffer.classifier.weight"].t() + p["buffer.classifier.bias"]  # :223 [N,1]
    M = A @ B                                                             # :227 [3,1]
    wROIs = A * B.view(N)                                                 # :228
    logit = M.view(1, 3)                                                  # :229,:233
    y_pred = F.softmax(logit, dim=1)                                      # :235
    Yl = Y.long().view(-1)
    y_hat = torch.argmax(y_pred).long()                                   # :240
    tgt = smooth_target(Yl, 3, SMOOTHING, H.dtype)
    logp = F.log_softmax(logit, dim=1)
    cw = torch.ones(3, dtype=H.dtype) if class_weights is None else class_weights.to(H.dtype)
    loss = (-(tgt * cw.view(1, 3) * logp).sum(dim=1)).mean()              # nnBlocks.py:121-133
    error = 1 - y_hat.eq(Yl).float()                                      # :242
    l2 = torch.stack([p["buffer.lin1.weight"].norm(), p["buffer.classifier.weight"].norm()]).mean()  # :246
    return {"Aterm": A, "wROIs": wROIs, "Bterm": B, "Mterm": M, "Fterm": H, "Aterm_mu": Aterm_mu,
            "Aterm_var": Aterm_var, "loss": loss, "l2": l2, "KLD": KLD, "y_pred": y_pred,
            "y_pred_hat": y_hat, "error": error}


def subsample_indices(n_bag: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Train-mode 20 % subsample (gbm/model.py:193): randperm on the CPU generator."""
    return torch.randperm(n_bag, generator=generator)[: int(n_bag * SUBSAMPLE)]


def attention_forward(p: Dict[str, torch.Tensor], bag: torch.Tensor, Y: torch.Tensor,
                      class_weights: Optional[torch.Tensor] = None, training: bool = False,
                      indices: Optional[torch.Tensor] = None,
                      drop_mask: Optional[torch.Tensor] = None, emulate_bf16: str = "") -> Dict[str, torch.Tensor]:
    """Whole Attention.forward (gbm/model.py:189-264).  In training mode pass `indices`
    (the subsample) and `drop_mask` explicitly so the run is reproducible."""
    x = bag.detach()
    if training:
        if indices is None:
            indices = subsample_indices(bag.shape[0])
        x = x[indices]
    H = resnet26_forward(p, x, emulate_bf16=emulate_bf16)
    return head_forward(p, H, Y, class_weights, drop_mask if training else None)


def forward_backward(p: Dict[str, torch.Tensor], bag: torch.Tensor, Y: torch.Tensor,
                     class_weights: Optional[torch.Tensor] = None, training: bool = False,
                     indices: Optional[torch.Tensor] = None,
                     drop_mask: Optional[torch.Tensor] = None, emulate_bf16: str = ""):
    """Returns (outputs dict, grads dict of d loss / d param) using autograd on the restatement.
    emulate_bf16="act+w": forward AND backward with the CUDA bf16 mode's rounding points (see resnet26_forward,
    _RoundOperand, _RoundStored) -- the tight gate for the bf16 gradients."""
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out = attention_forward(q, bag, Y, class_weights, training, indices, drop_mask, emulate_bf16)
    out["loss"].backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in q.items())
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads
