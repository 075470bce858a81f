"""`WideAttention`: the same MIL module with the reference's `alt_resnet.py` network as tile feature extractor
(SURVEY.md section 8f, N4).

alt_resnet.py is torchvision's ResNet with the BatchNorm layers stripped: conv 7x7/2 (3 -> 64, no bias), ReLU, max-pool,
four layers of `BasicBlock`s (conv3x3 -> ReLU -> conv3x3 -> += identity | conv1x1/2 -> ReLU, alt_resnet.py:35-67) with
widths 64 / 128 / 256 / 512 (:87-90), average pool, `fc` WITH bias (:90).  What the reference would write to use it,

    self.cnn = nn.DataParallel(alt_resnet.ResNet(alt_resnet.BasicBlock, [2, 2, 2, 2], num_classes=self.L))

is `WideAttention(n_classes, class_weights, layers=(2, 2, 2, 2))` here: the same module tree (=> the same state-dict
keys `cnn.module.layer2.0.downsample.0.weight`, ... and the same RNG stream at construction), the same
`forward(full_input, Y) -> dict`, the same head kernels.  The extractor runs on the wide-channel tcgen05 kernels
(csrc/mil_wide_conv.cu, mil_wide_wgrad.cu, mil_wide_net.cu), bf16 activations with fp32 accumulation.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib
from .distributed import BagGroup
from .model import (Attention, _Lease, _check_param, _ptr, _stream)


class _WideDesc(C.Structure):
    _fields_ = [("layers", C.c_int * 4), ("widths", C.c_int * 4), ("stem", C.c_int), ("features", C.c_int),
                ("slope", C.c_float)]


# ------------------------------------------------------------------------------------------------------
# parameter holders mirroring alt_resnet.py's module tree (construction order = RNG order)
# ------------------------------------------------------------------------------------------------------
def _conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)      # alt_resnet.py:24-27


def _conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)                 # alt_resnet.py:30-32


class AltBasicBlock(nn.Module):
    """Parameter holder with alt_resnet.BasicBlock's attribute names (alt_resnet.py:35-50)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter holder: the B200 path runs the whole extractor in libmil_b200.so")


class AltResNet(nn.Module):
    """Parameter holder for alt_resnet.ResNet (alt_resnet.py:70-123), including its kaiming_normal_ pass (:93-95)."""

    def __init__(self, layers=(2, 2, 2, 2), num_classes=80, widths=(64, 128, 256, 512)):
        super().__init__()
        self.inplanes = widths[0]
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(widths[0], layers[0])
        self.layer2 = self._make_layer(widths[1], layers[1], stride=2)
        self.layer3 = self._make_layer(widths[2], layers[2], stride=2)
        self.layer4 = self._make_layer(widths[3], layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(widths[3], num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(_conv1x1(self.inplanes, planes, stride))
        layers = [AltBasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(AltBasicBlock(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter holder: use WideAttention.forward")


# ------------------------------------------------------------------------------------------------------
# the autograd function
# ------------------------------------------------------------------------------------------------------
class _WideFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, bag, Y, idx, drop, n_global, *params):
        with torch.cuda.device(bag.device):
            return _WideFunction._forward(ctx, owner, bag, Y, idx, drop, n_global, *params)

    @staticmethod
    def _forward(ctx, owner, bag, Y, idx, drop, n_global, *params):
        lib = _lib.load()
        dev = bag.device
        for nm, p in zip(owner._param_names, params):
            _check_param(nm, p)
            if p.device != dev:
                raise RuntimeError(f"parameter {nm} is on {p.device} but the bag is on {dev}")
        pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        pph = owner._head_pointer_array(params)
        n = int(idx.numel()) if idx is not None else int(bag.shape[0])
        side = int(bag.shape[2])
        group: BagGroup = owner.bag_group
        if n_global < 2:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                             f"torch.Size([{n_global}, 80])")
        need_grad = any(ctx.needs_input_grad[6:])
        desc = owner._desc
        nbytes = int(lib.mil_wide_workspace_bytes(C.byref(desc), n, side))
        if nbytes == 0:
            _lib.check(1, "mil_wide_workspace_bytes")
        ws = owner._pool.acquire((n, side, "wide", dev.index), nbytes, dev)
        lease = _Lease(ws)
        st = _stream(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        H = torch.empty((n, 80), **f32)
        _lib.check(lib.mil_wide_forward(C.byref(desc), pp, _ptr(bag), int(bag.dtype == torch.uint8), _ptr(idx), n, side,
                                        _ptr(ws.buf), nbytes, _ptr(H), st), "mil_wide_forward")
        # ---- head (gbm/model.py:200-246): the same kernels as the ResNet-26 path ----
        hws_bytes = int(lib.mil_head_workspace_bytes(n))
        hws = torch.empty(hws_bytes, dtype=torch.uint8, device=dev)
        small = torch.empty(160 + 16 + 160, dtype=torch.float64, device=dev)
        stats, sums = small[:160], small[160:176]
        _lib.check(lib.mil_head_stats(_ptr(H), n, _ptr(hws), hws_bytes, _ptr(stats), st), "mil_head_stats")
        group.all_reduce_sum(stats)                                                     # AR-1
        raw = torch.empty((n, 3), **f32)
        g = torch.empty((n, 3), **f32)
        b = torch.empty((n, 1), **f32)
        _lib.check(lib.mil_head_scores(pph, _ptr(H), _ptr(drop), n, n_global, _ptr(stats), _ptr(raw), _ptr(g), _ptr(b),
                                       _ptr(hws), hws_bytes, _ptr(sums), st), "mil_head_scores")
        group.all_reduce_sum(sums)                                                      # AR-2
        A = torch.empty((3, n), **f32)
        wroi = torch.empty((3, n), **f32)
        scal = torch.empty(32, **f32)
        cw = owner._class_weights(dev)
        _lib.check(lib.mil_head_finalize(_ptr(sums), _ptr(stats), n_global, _ptr(Y), _ptr(cw), n, _ptr(g), _ptr(b),
                                         _ptr(A), _ptr(wroi), _ptr(scal), st), "mil_head_finalize")
        if need_grad:
            ctx.lease = lease
            ctx.owner = owner
            ctx.meta = (n, side, n_global, nbytes)
            ctx.drop = drop
            ctx.params = params
            ctx.save_for_backward(H, raw, g, b, scal, small)
        else:
            lease.release()
        sc = scal.clone()
        loss = scal[12].clone()
        outs = (loss, A, wroi, b, sc[0:3].reshape(3, 1), H, sc[13], sc[14], sc[15], sc[3:6].reshape(1, 3),
                sc[16].long(), sc[17:18])
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, gloss, *_):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            return _WideFunction._backward(ctx, gloss)

    @staticmethod
    def _backward(ctx, gloss):
        lib = _lib.load()
        H, raw, g, b, scal, small = ctx.saved_tensors
        stats, bnsums = small[:160], small[176:]
        owner = ctx.owner
        n, side, n_global, nbytes = ctx.meta
        params = ctx.params
        dev = H.device
        group: BagGroup = owner.bag_group
        pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        pph = owner._head_pointer_array(params)
        st = _stream(dev)
        desc = owner._desc
        total = int(lib.mil_wide_param_total(C.byref(desc)))
        gflat = getattr(owner, "_gflat", None)
        direct = gflat is not None and gflat.device == dev and all(ctx.needs_input_grad[6:])
        grads = gflat if (direct and group.world == 1) else torch.zeros(total, dtype=torch.float32, device=dev)
        # the head's kernels write their parameter gradients at the ResNet-26 table's offsets: a scratch buffer of that
        # layout, copied into this model's flat buffer below
        hg = torch.zeros(int(lib.mil_param_total()), dtype=torch.float32, device=dev)
        gl = gloss.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dHz = torch.empty_like(H)
        dHi = torch.empty_like(H)
        dH = torch.empty_like(H)
        hws_bytes = int(lib.mil_head_workspace_bytes(n))
        hws = torch.empty(hws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.mil_head_backward_a(pph, _ptr(H), _ptr(ctx.drop), n, n_global, _ptr(stats), _ptr(raw), _ptr(g),
                                           _ptr(b), _ptr(scal), _ptr(gl), _ptr(dHz), _ptr(dHi), _ptr(hg),
                                           _ptr(hws), hws_bytes, _ptr(bnsums), st), "mil_head_backward_a")
        group.all_reduce_sum(bnsums)                                                    # AR-3
        _lib.check(lib.mil_head_backward_b(pph, _ptr(H), n, n_global, _ptr(stats), _ptr(bnsums), _ptr(dHz), _ptr(dHi),
                                           _ptr(dH), st), "mil_head_backward_b")
        for src_off, dst_off, numel in owner._head_grad_map:
            grads[dst_off:dst_off + numel].add_(hg[src_off:src_off + numel])
        ws = ctx.lease.ws
        _lib.check(lib.mil_wide_backward(C.byref(desc), pp, n, side, _ptr(ws.buf), nbytes, _ptr(dH), _ptr(grads), st),
                   "mil_wide_backward")
        ctx.lease.release()
        if group.world > 1:
            group.all_reduce_grads(grads)                                               # AR-4
        if direct:
            if grads is not gflat:
                gflat.add_(grads)
            return (None,) * (6 + len(params))
        out = []
        for (nm, shape, off), need in zip(owner._param_table, ctx.needs_input_grad[6:]):
            numel = 1
            for s in shape:
                numel *= s
            out.append(grads[off:off + numel].view(shape) if need else None)
        return (None, None, None, None, None, None, *out)


# ------------------------------------------------------------------------------------------------------
# the module
# ------------------------------------------------------------------------------------------------------
class WideAttention(Attention):
    """`Attention` (gbm/model.py:114-264) with alt_resnet.ResNet(BasicBlock, layers, num_classes=80) as `cnn`.

    layers  blocks per layer (alt_resnet.py:157-165: resnet18 = (2, 2, 2, 2); resnet34's (3, 4, 6, 3) works too)
    widths  channels of layer1..4 (alt_resnet.py:87-90: 64, 128, 256, 512); the stem has widths[0] channels (:81)
    slope   negative slope of the extractor's activation (0 = alt_resnet's ReLU)
    Everything else -- forward(full_input, Y) -> the reference's 13-key dict, train / eval behaviour, bag_group, uint8
    bags, FusedAdam / GraphedStep -- is inherited.  bf16 only."""

    def __init__(self, n_classes, class_weights=None, layers=(2, 2, 2, 2), widths=(64, 128, 256, 512), slope: float = 0.0):
        object.__setattr__(self, "_wide_cfg", (tuple(int(v) for v in layers), tuple(int(v) for v in widths), float(slope)))
        super().__init__(n_classes, class_weights)
        self.precision = "bf16"
        lay, wid, sl = self._wide_cfg
        self._desc = _WideDesc((C.c_int * 4)(*lay), (C.c_int * 4)(*wid), wid[0], self.L, sl)
        self._head_map_cache = None

    def _make_extractor(self):
        lay, wid, _ = self._wide_cfg
        return AltResNet(layers=lay, num_classes=self.L, widths=wid)

    # ---- plumbing: this model's own parameter table ----
    @property
    def _param_table(self):
        if self._param_table_cache is None:
            lib = _lib.load()
            cnt = int(lib.mil_wide_param_count(C.byref(self._desc)))
            if cnt == 0:
                _lib.check(1, "mil_wide_param_count")
            table = []
            name = C.create_string_buffer(160)
            nd = C.c_int(0)
            shp = (C.c_longlong * 4)()
            off = C.c_longlong(0)
            for i in range(cnt):
                _lib.check(lib.mil_wide_param_info(C.byref(self._desc), i, name, 160, C.byref(nd), shp, C.byref(off)),
                           "mil_wide_param_info")
                table.append((name.value.decode(), tuple(int(shp[k]) for k in range(nd.value)), int(off.value)))
            mine = dict(self.named_parameters())
            if list(mine.keys()) != [t[0] for t in table]:
                raise RuntimeError("module parameter order differs from the library's state-dict order")
            for nm, shape, _ in table:
                if tuple(mine[nm].shape) != shape:
                    raise RuntimeError(f"parameter {nm}: shape {tuple(mine[nm].shape)} != library {shape}")
            self._param_table_cache = table
        return self._param_table_cache

    def _head_tables(self):
        """(indices of the head's tensors in this model's table, their indices / offsets in the ResNet-26 table)."""
        if self._head_map_cache is None:
            thin = _lib.param_table()
            thin_idx = {nm: (i, off) for i, (nm, _, off) in enumerate(thin)}
            ptr_map, grad_map = [], []
            for i, (nm, shape, off) in enumerate(self._param_table):
                if nm.startswith("cnn."):
                    continue
                ti, toff = thin_idx[nm]
                numel = 1
                for s in shape:
                    numel *= s
                ptr_map.append((i, ti))
                grad_map.append((toff, off, numel))
            self._head_map_cache = (len(thin), ptr_map, grad_map)
        return self._head_map_cache

    @property
    def _head_grad_map(self):
        return self._head_tables()[2]

    def _head_pointer_array(self, params):
        n_thin, ptr_map, _ = self._head_tables()
        arr = (C.c_void_p * n_thin)()
        for i, ti in ptr_map:
            arr[ti] = params[i].data_ptr()
        return arr

    def _layer_bucket_bounds(self):  # pragma: no cover
        raise RuntimeError("WideAttention reduces its gradients in one all-reduce")

    def forward(self, full_input: torch.Tensor, Y: Optional[torch.Tensor] = None, bag_tiles: Optional[int] = None):
        if self.precision != "bf16":
            raise ValueError("WideAttention runs in bf16 (fp32 accumulation); there is no fp32 check mode for it")
        bag, Yl, idx, drop, n_global = self._prepare(full_input, Y, bag_tiles)
        params = self._params()
        (loss, A, wroi, b, M, H, amu, avar, kld, ypred, yhat, err) = _WideFunction.apply(
            self, bag, Yl, idx, drop, n_global, *params)
        l2 = (self.buffer.lin1.weight.norm() + self.buffer.classifier.weight.norm()) * 0.5       # gbm/model.py:246
        return {
            'Aterm': A, 'wROIs': wroi, 'Bterm': b, 'Mterm': M, 'Fterm': H, 'Aterm_mu': amu, 'Aterm_var': avar,
            'loss': loss, 'l2': l2, 'KLD': kld, 'y_pred': ypred, 'y_pred_hat': yhat, 'error': err,
        }

    def features(self, full_input):
        raise RuntimeError("WideAttention: use forward(...)['Fterm']")
