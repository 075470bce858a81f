"""Training-loop glue for the drop-in.

The reference trains with `optim.Adam(classifier.parameters(), lr=2e-4)` and steps every 5 bags
(gbm/classify_combined.py:519, :446-454): 65 parameter tensors => ~400 small kernels per step.  `FusedAdam` keeps
the 640 967 parameters and their gradients in two flat device buffers (state-dict order, the layout the extractor's
backward already writes) and does the whole step in ONE launch of `mil_adam_step`; gradient accumulation over
several bags is autograd's ordinary in-place accumulation into the flat gradient views.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


def flatten_parameters(model):
    """Re-home every parameter (and its .grad) of an `Attention` module in two flat fp32 buffers, in the library's
    state-dict order.  Call it after `.cuda()`; `state_dict()` / `load_state_dict()` keep working (same names, the
    tensors are views).  Returns (flat_params, flat_grads)."""
    table = model._param_table
    params = dict(model.named_parameters())
    dev = next(iter(params.values())).device
    if dev.type != "cuda":
        raise RuntimeError("flatten_parameters: move the module to a CUDA device first (no CPU fallback)")
    total = max(off + params[name].numel() for name, _, off in table)     # the model's own table (ResNet-26 or wide)
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    gflat = torch.zeros(total, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for name, shape, off in table:
            p = params[name]
            n = p.numel()
            flat[off:off + n].copy_(p.detach().reshape(-1))
            if p.grad is not None:
                gflat[off:off + n].copy_(p.grad.reshape(-1))
            p.data = flat[off:off + n].view(shape)
            p.grad = gflat[off:off + n].view(shape)
    model._gflat = gflat      # the backward pass accumulates straight into it (model._MilFunction.backward)
    return flat, gflat


class FusedAdam(torch.optim.Optimizer):
    """Adam with torch.optim.Adam's arithmetic (L2 weight decay, no amsgrad) as one kernel over the flat buffers.
    `param_groups[0]['lr']` is read at every step, so LR schedules written against torch optimizers (the reference's
    `SetStage`, gbm/classify_combined.py:110-138) keep working."""

    def __init__(self, model, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self._flat, self._gflat = flatten_parameters(model)
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._m = torch.zeros_like(self._flat)
        self._v = torch.zeros_like(self._flat)
        self._t = 0
        self._hyper = None        # device float[6]: set by graph.GraphedStep (the captured step reads its scalars there)

    def hyper_values(self, t):
        """{step_size, beta1, beta2, bc2_sqrt, eps, weight_decay} of step number t (1-based), as mil_adam_step takes them."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        return [g["lr"] / (1.0 - b1 ** t), b1, b2, math.sqrt(1.0 - b2 ** t), g["eps"], g["weight_decay"]]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        P = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(self._flat.device).cuda_stream)
        if self._hyper is not None:
            # inside (the capture of) a graphed step: the scalars live in device memory, GraphedStep refreshes them
            # and counts the steps
            with torch.cuda.device(self._flat.device):
                _lib.check(_lib.load().mil_adam_step_dev(P(self._flat), P(self._gflat), P(self._m), P(self._v),
                                                         self._flat.numel(), P(self._hyper), st), "mil_adam_step_dev")
            return loss
        self._t += 1
        step_size, b1, b2, bc2_sqrt, eps, wd = self.hyper_values(self._t)
        with torch.cuda.device(self._flat.device):
            _lib.check(_lib.load().mil_adam_step(P(self._flat), P(self._gflat), P(self._m), P(self._v),
                                                 self._flat.numel(), step_size, b1, b2, bc2_sqrt, eps, wd, st),
                       "mil_adam_step")
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """Keeps the flat gradient views in place (set_to_none would detach them from the flat buffer)."""
        self._gflat.zero_()

    def state_dict(self):
        return {"t": self._t, "exp_avg": self._m.clone(), "exp_avg_sq": self._v.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._t = int(sd["t"])
        self._m.copy_(sd["exp_avg"])
        self._v.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


# ---- the training driver's glue (gbm/classify_combined.py:110-138, 468-474, 521-535) ------------------------------
STAGE_SCHEDULE = (0, 10, 150, 250, 340)      # epochs: warm-up | main | check | freeze | stop
BASE_LR = 0.0002


def set_stage(optimizer, model, epoch: int, test: bool = False, base_lr: float = BASE_LR, schedule=STAGE_SCHEDULE,
              verbose: bool = True):
    """The reference's `SetStage` (gbm/classify_combined.py:110-138): learning rate and train / eval mode of the epoch.
    Warm-up: base_lr / (10 - epoch), train;  main: base_lr, train;  check: base_lr / 2;  freeze: base_lr / 10 (both:
    eval when `test`, else train).  Past the schedule the reference saves `..._FINAL.model` and exits: here the stage is
    returned as "Stop" and the caller decides (save_checkpoint(..., final=True)).  Returns (stage, lr).
    Works with any optimizer that reads `param_groups[...]['lr']` at step time (FusedAdam does)."""
    def set_lr(lr):
        for g in optimizer.param_groups:
            g["lr"] = lr
    stage, lr = None, None
    if schedule[0] <= epoch < schedule[1]:
        stage, lr = "Warmup", base_lr / (schedule[1] - epoch)
        set_lr(lr)
        model.train()
    if schedule[1] <= epoch < schedule[2]:
        stage, lr = "Main", base_lr
        set_lr(lr)
        model.train()
    if schedule[2] <= epoch < schedule[3]:
        stage, lr = "Check", base_lr / 2.0
        set_lr(lr)
        model.eval() if test else model.train()
    if schedule[3] <= epoch < schedule[4]:
        stage, lr = "Freeze", base_lr / 10.0
        set_lr(lr)
        model.eval() if test else model.train()
    if epoch > schedule[4]:
        stage, lr = "Stop", 0.0
    if stage is None:                      # epoch == schedule[4]: the reference falls through every branch
        stage, lr = "Hold", optimizer.param_groups[0]["lr"]
    if verbose:
        print("Stage = [{0}], lr = [{1}], training mode = [{2}]".format(stage, lr, model.training))
    return stage, lr


def save_checkpoint(path, classifier, optimizer=None, final: bool = False):
    """The reference's checkpoint file (gbm/classify_combined.py:468-474; :137 for the final one without the optimizer):
    torch.save({'classifier': state_dict, 'optimizer': state_dict}).  The classifier's keys are the reference's own
    (cnn.module.*, context.bn.*, ...), so the file loads into the reference model and vice versa."""
    blob = {"classifier": {k: v.detach().cpu().clone() for k, v in classifier.state_dict().items()}}
    if optimizer is not None and not final:
        blob["optimizer"] = optimizer.state_dict()
    torch.save(blob, path)
    return path


def load_checkpoint(path, classifier, optimizer=None, transfer: bool = False, map_location="cpu"):
    """`--ckpt` of the reference (gbm/classify_combined.py:521-535): the full classifier (strict=False), or with
    `transfer` only the extractor's convolutions (keys containing 'cnn' and 'conv').  The optimizer state is restored
    when the file has one and an optimizer is given (the reference never reloads it)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["classifier"]
    if transfer:
        sd = {k: v for k, v in sd.items() if "cnn" in k and "conv" in k}
    missing = classifier.load_state_dict(sd, strict=False)
    if optimizer is not None and "optimizer" in ckpt and not transfer:
        optimizer.load_state_dict(ckpt["optimizer"])
    return missing
