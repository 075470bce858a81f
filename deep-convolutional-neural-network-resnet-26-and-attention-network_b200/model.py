"""Drop-in `Attention` module: the reference's nn.Module surface (gbm/model.py:114-264) over the B200 kernels.

Same constructor (`Attention(n_classes, class_weights=None)`), same parameter tree and state-dict keys
(`cnn.module.layer1.0.conv1.weight`, ... -- the DataParallel prefix is kept so that the reference's checkpoints
and its `'cnn' in k and 'conv' in k` transfer filter, gbm/classify_combined.py:524-535, keep working), same
`forward(full_input, Y) -> dict` with the same 13 keys, shapes, dtypes and detach pattern
(gbm/model.py:249-264), same init (gbm/model.py:161-181), same train/eval behaviour (20 % CPU-randperm
subsample + Dropout(0.25) in train mode, gbm/model.py:192-196,107).

The nn.Conv2d / nn.Linear / nn.BatchNorm1d sub-modules are parameter HOLDERS only: their forward is never
called.  All arithmetic happens in libmil_b200.so through one torch.autograd.Function; PyTorch provides device
memory, the stream and (for a bag sharded over several GPUs) torch.distributed.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
from torch import nn
from torch.nn import init

from . import _lib
from .distributed import BagGroup

DTYPE_CODES = {"fp32": 0, "bf16": 1}
SUBSAMPLE = 0.2        # gbm/model.py:193
DROP_P = 0.25          # gbm/model.py:107


# ------------------------------------------------------------------------------------------------------
# parameter holders mirroring the reference module tree
# ------------------------------------------------------------------------------------------------------
class BasicResBlock(nn.Module):
    """Parameter holder with the reference's attribute names (nnBlocks.py:157-173)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64):
        super().__init__()
        if groups != 1 or base_width != 64:
            raise ValueError('BasicBlock only supports groups=1 and base_width=64')     # nnBlocks.py:165-166
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=True)
        self.relu = nn.LeakyReLU(0.1)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter holder: the B200 path runs the whole extractor in libmil_b200.so")


class ResNet(nn.Module):
    """Parameter holder for the bias-only LeakyReLU ResNet-26 (gbm/model.py:14-48)."""

    def __init__(self, block=BasicResBlock, layers=(3, 3, 3, 3), num_classes=80):
        super().__init__()
        if tuple(layers) != (3, 3, 3, 3) or num_classes != 80:
            raise ValueError("the B200 kernels implement the reference configuration layers=[3,3,3,3], fc 80->80")
        self.inplanes = 20
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3)
        self.relu = nn.LeakyReLU(0.1, inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 20, layers[0])
        self.layer2 = self._make_layer(block, 40, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 60, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 80, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(80, num_classes, bias=False)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("parameter holder: use Attention.forward / Attention.features")


class _ModuleHolder(nn.Module):
    """Stands where the reference has nn.DataParallel(ResNet) (gbm/model.py:132-135): keeps the `.module`
    attribute and therefore the `cnn.module.*` state-dict keys.  Tile parallelism is one process per GPU."""

    def __init__(self, module):
        super().__init__()
        self.module = module


class ContextLayer(nn.Module):
    """Parameter holder for gbm/model.py:89-111 (bag-wide BatchNorm1d without running stats, Dropout 0.25)."""

    def __init__(self, features):
        super().__init__()
        self.L = features
        self.bn = nn.BatchNorm1d(self.L, track_running_stats=False)
        self.relu = nn.LeakyReLU(0.1)
        self.do = nn.Dropout(DROP_P)


class CrossEntropyWithProbs(nn.Module):
    """Configuration holder for nnBlocks.py:47-69 (the loss itself is evaluated by mil_head_finalize)."""

    def __init__(self, classes: int, smoothing=0.0, weight: Optional[torch.Tensor] = None, reduction: str = "mean"):
        super().__init__()
        self.smoothing = smoothing
        self.num_classes = classes
        self.weight = weight
        self.reduction = reduction


# ------------------------------------------------------------------------------------------------------
# workspace pool
# ------------------------------------------------------------------------------------------------------
class _Workspace:
    def __init__(self, key, nbytes, device):
        self.key = key
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.busy = False


class _WorkspacePool:
    """Extractor workspaces keyed by (tiles, side, dtype, device).  A workspace handed to a forward pass that
    recorded a graph stays busy until its backward has run or the graph is dropped."""

    def __init__(self):
        self._items: Dict[tuple, List[_Workspace]] = {}

    def acquire(self, key, nbytes, device) -> _Workspace:
        lst = self._items.setdefault(key, [])
        for ws in lst:
            if not ws.busy and ws.buf.numel() >= nbytes:     # the size also depends on the library's runtime switches
                ws.busy = True
                return ws
        lst[:] = [w for w in lst if w.busy]
        # drop idle workspaces of other shapes before growing (they can be several GB)
        for k in list(self._items):
            if k != key:
                self._items[k] = [w for w in self._items[k] if w.busy]
        ws = _Workspace(key, nbytes, device)
        ws.busy = True
        lst.append(ws)
        return ws

    @staticmethod
    def release(ws: _Workspace):
        ws.busy = False


class _Lease:
    """Releases the workspace when the autograd context that owns it goes away."""

    def __init__(self, ws):
        self.ws = ws
        self._fin = weakref.finalize(self, _WorkspacePool.release, ws)

    def release(self):
        self._fin()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_param(name, p):
    if not p.is_cuda:
        raise RuntimeError(f"parameter {name} is on {p.device}: the B200 path has no CPU fallback; call .cuda()")
    if p.dtype != torch.float32 or not p.is_contiguous():
        raise RuntimeError(f"parameter {name} must be a contiguous fp32 tensor (got {p.dtype})")


# ------------------------------------------------------------------------------------------------------
# the autograd function: everything between the bag and the 13 outputs
# ------------------------------------------------------------------------------------------------------
class _MilFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, bag, Y, idx, drop, n_global, *params):
        with torch.cuda.device(bag.device):      # the library launches on the CURRENT device: make it the bag's
            return _MilFunction._forward(ctx, owner, bag, Y, idx, drop, n_global, *params)

    @staticmethod
    def _forward(ctx, owner, bag, Y, idx, drop, n_global, *params):
        lib = _lib.load()
        dev = bag.device
        names = owner._param_names
        for nm, p in zip(names, params):
            _check_param(nm, p)
            if p.device != dev:
                raise RuntimeError(f"parameter {nm} is on {p.device} but the bag is on {dev}")
        pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        n = int(idx.numel()) if idx is not None else int(bag.shape[0])
        side = int(bag.shape[2])
        dt = DTYPE_CODES[owner.precision]
        group: BagGroup = owner.bag_group
        if n_global < 2:
            # nn.BatchNorm1d refuses a single row, in train and eval mode alike (reference behaviour)
            raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                             f"torch.Size([{n_global}, 80])")
        need_grad = any(ctx.needs_input_grad[6:])
        # no gradient wanted (validation / attention-map extraction under torch.no_grad(),
        # gbm/classify_combined.py:221-357): the forward-only extractor keeps nothing for a backward pass
        wsb = lib.mil_extractor_workspace_bytes if need_grad else lib.mil_extractor_infer_workspace_bytes
        nbytes = int(wsb(n, side, dt))
        if nbytes == 0:
            _lib.check(1, "mil_extractor_workspace_bytes")
        ws = owner._pool.acquire((n, side, dt, dev.index, need_grad), nbytes, dev)
        lease = _Lease(ws)
        st = _stream(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        H = torch.empty((n, 80), **f32)
        if need_grad:
            fwd = lib.mil_extractor_forward_u8 if bag.dtype == torch.uint8 else lib.mil_extractor_forward
            _lib.check(fwd(pp, _ptr(bag), _ptr(idx), n, side, dt, _ptr(ws.buf), nbytes, _ptr(H), st),
                       "mil_extractor_forward")
        else:
            _lib.check(lib.mil_extractor_infer(pp, _ptr(bag), int(bag.dtype == torch.uint8), _ptr(idx), n, side, dt,
                                               _ptr(ws.buf), nbytes, _ptr(H), st), "mil_extractor_infer")
        # ---- head (gbm/model.py:200-246) with the three bag-wide sums reduced across the bag group ----
        hws_bytes = int(lib.mil_head_workspace_bytes(n))
        hws = torch.empty(hws_bytes, dtype=torch.uint8, device=dev)
        small = torch.empty(160 + 16 + 160, dtype=torch.float64, device=dev)
        stats, sums, bnsums = small[:160], small[160:176], small[176:]
        _lib.check(lib.mil_head_stats(_ptr(H), n, _ptr(hws), hws_bytes, _ptr(stats), st), "mil_head_stats")
        group.all_reduce_sum(stats)                                                     # AR-1
        raw = torch.empty((n, 3), **f32)
        g = torch.empty((n, 3), **f32)
        b = torch.empty((n, 1), **f32)
        _lib.check(lib.mil_head_scores(pp, _ptr(H), _ptr(drop), n, n_global, _ptr(stats), _ptr(raw), _ptr(g), _ptr(b),
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
            ctx.meta = (n, side, dt, n_global, nbytes)
            ctx.bag, ctx.idx, ctx.drop = bag, idx, drop
            ctx.params = params
            ctx.save_for_backward(H, raw, g, b, scal, small)
        else:
            lease.release()
        sc = scal.clone()     # the detached scalar outputs are views of ONE copy (scal itself is saved for backward)
        loss = scal[12].clone()
        outs = (loss, A, wroi, b, sc[0:3].reshape(3, 1), H, sc[13], sc[14], sc[15], sc[3:6].reshape(1, 3),
                sc[16].long(), sc[17:18])
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, gloss, *_):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            return _MilFunction._backward(ctx, gloss)

    @staticmethod
    def _backward(ctx, gloss):
        lib = _lib.load()
        H, raw, g, b, scal, small = ctx.saved_tensors
        stats, bnsums = small[:160], small[176:]
        owner = ctx.owner
        n, side, dt, n_global, nbytes = ctx.meta
        params = ctx.params
        dev = H.device
        group: BagGroup = owner.bag_group
        pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        st = _stream(dev)
        total = int(lib.mil_param_total())
        # flattened module (optim.flatten_parameters / FusedAdam): the library accumulates straight into the flat
        # gradient buffer the parameters' .grad are views of -- no per-tensor AccumulateGrad work (65 small kernels
        # and ~1 ms of host time per step otherwise).  Sharded bags reduce this step's gradients first.
        gflat = getattr(owner, "_gflat", None)
        direct = gflat is not None and gflat.device == dev and all(ctx.needs_input_grad[6:])
        if direct and group.world == 1:
            grads = gflat
        else:
            grads = torch.zeros(total, dtype=torch.float32, device=dev)
        gl = gloss.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dHz = torch.empty_like(H)
        dHi = torch.empty_like(H)
        dH = torch.empty_like(H)
        hws_bytes = int(lib.mil_head_workspace_bytes(n))
        hws = torch.empty(hws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.mil_head_backward_a(pp, _ptr(H), _ptr(ctx.drop), n, n_global, _ptr(stats), _ptr(raw), _ptr(g),
                                           _ptr(b), _ptr(scal), _ptr(gl), _ptr(dHz), _ptr(dHi), _ptr(grads),
                                           _ptr(hws), hws_bytes, _ptr(bnsums), st), "mil_head_backward_a")
        group.all_reduce_sum(bnsums)                                                    # AR-3
        _lib.check(lib.mil_head_backward_b(pp, _ptr(H), n, n_global, _ptr(stats), _ptr(bnsums), _ptr(dHz), _ptr(dHi),
                                           _ptr(dH), st), "mil_head_backward_b")
        ws = ctx.lease.ws
        if group.world > 1:
            # AR-4, bucketed and overlapped: the extractor backward records an event per finished layer, the
            # bucket of that layer is all-reduced on a side stream while the earlier layers are still running
            events = [torch.cuda.Event() for _ in range(4)]
            for ev in events:
                ev.record()                         # materialise the handles (they are re-recorded by the library)
            evp = (C.c_void_p * 4)(*[ev.cuda_event for ev in events])
            _lib.check(lib.mil_extractor_backward_staged(pp, _ptr(ctx.bag), _ptr(ctx.idx), n, side, dt, _ptr(ws.buf),
                                                         nbytes, _ptr(dH), _ptr(grads), evp, st),
                       "mil_extractor_backward_staged")
            events[0] = torch.cuda.Event()          # conv1 (stem) gradients come last: final when the call's work is
            events[0].record()
            group.all_reduce_grads_staged(grads, events, owner._layer_bucket_bounds())
        else:
            _lib.check(lib.mil_extractor_backward(pp, _ptr(ctx.bag), _ptr(ctx.idx), n, side, dt, _ptr(ws.buf), nbytes,
                                                  _ptr(dH), _ptr(grads), st), "mil_extractor_backward")
        ctx.lease.release()
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
class Attention(nn.Module):
    """B200-native drop-in for the reference `model.Attention` (gbm/model.py:114-264).

    Extra, non-reference attributes (all optional):
      precision   "bf16" (default; bf16 activations, fp32 accumulation) or "fp32" (check mode);
                  env MIL_B200_PRECISION overrides the default
      bag_group   BagGroup: the ranks that share ONE bag (its tiles sharded over them); default single rank
      drop_mask   inject a fixed {0,1} keep mask [n,80] for Dropout(0.25) (parity tests); None -> random
      subsample_indices  inject the train-mode tile subset (parity tests); None -> torch.randperm as in the
                  reference (gbm/model.py:193)
    """

    def __init__(self, n_classes, class_weights=None):
        super().__init__()
        self.L = 80    # Input features to attention mechanism        (gbm/model.py:120-124)
        self.D = 40    # Hidden dimension for attention mechanism
        self.O = 1     # Output nodes
        self.K = 3     # Attention maps
        self.C = n_classes

        if class_weights is not None:
            print("I'm weighting classes according to class_weights=", class_weights)
            self.loss = CrossEntropyWithProbs(classes=3, weight=class_weights, smoothing=0.25)
        else:
            self.loss = CrossEntropyWithProbs(classes=3, smoothing=0.25)

        self.cnn = _ModuleHolder(self._make_extractor())
        self.context = ContextLayer(self.L)
        self.attention = nn.Sequential(OrderedDict([
            ('lin1', nn.Linear(self.L, self.D)),
            ('tanh', nn.Tanh()),
            ('lin2', nn.Linear(self.D, self.K)),
        ]))
        self.buffer = nn.Sequential(OrderedDict([
            ('lin1', nn.Linear(self.L, self.D)),
            ('relu', nn.LeakyReLU(0.1)),
            ('classifier', nn.Linear(self.D, self.O)),
        ]))
        self.weight_mask = nn.Parameter(torch.tensor([0.25, 0.25, 0.25]))
        self.off_diag = 1 - torch.eye(3)

        self.precision = os.environ.get("MIL_B200_PRECISION", "bf16")
        self.bag_group = BagGroup()
        self.drop_mask = None
        self.subsample_indices = None
        self.verbose_init = False
        self._pool = _WorkspacePool()
        self._param_table_cache = None
        self._param_list_cache = None
        self._gflat = None            # set by optim.flatten_parameters: flat gradient buffer the .grad tensors view
        self.reset_params()

    def _make_extractor(self):
        """The tile feature extractor's parameter holder (gbm/model.py:132); wide.WideAttention substitutes alt_resnet's."""
        return ResNet(block=BasicResBlock, layers=[3, 3, 3, 3], num_classes=self.L)

    # ---- init: identical distributions to gbm/model.py:161-187 ----
    def weight_init(self, m, name=''):
        if isinstance(m, nn.Linear):
            if 'attention' in name:
                init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='tanh')
            elif 'classifier' in name:
                init.xavier_normal_(m.weight)
            else:
                init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='leaky_relu', a=0.1)
            if m.bias is not None:
                init.zeros_(m.bias)
        if isinstance(m, nn.Conv2d):
            init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='leaky_relu', a=0.1)
            if m.bias is not None:
                init.zeros_(m.bias)

    def reset_params(self):
        for name, m in self.named_modules():
            self.weight_init(m, name)

    def reset_linear(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='tanh')
                if m.bias is not None:
                    init.zeros_(m.bias)

    # ---- plumbing ----
    @property
    def _param_table(self):
        if self._param_table_cache is None:
            table = _lib.param_table()
            mine = OrderedDict(self.named_parameters())
            if list(mine.keys()) != [t[0] for t in table]:
                raise RuntimeError("module parameter order differs from the library's state-dict order")
            for nm, shape, _ in table:
                if tuple(mine[nm].shape) != shape:
                    raise RuntimeError(f"parameter {nm}: shape {tuple(mine[nm].shape)} != library {shape}")
            self._param_table_cache = table
        return self._param_table_cache

    @property
    def _param_names(self):
        return [t[0] for t in self._param_table]

    def _layer_bucket_bounds(self):
        """Float offsets [start of layer2, layer3, layer4] inside the flat gradient buffer (state-dict order)."""
        offs = {nm: off for nm, _, off in self._param_table}
        return [offs[f"cnn.module.layer{k}.0.conv1.weight"] for k in (2, 3, 4)]

    def _class_weights(self, device):
        w = self.loss.weight
        if w is None:
            return None
        return torch.as_tensor(w, dtype=torch.float32).to(device).contiguous()

    def _prepare(self, full_input, Y, bag_tiles=None):
        if not full_input.is_cuda:
            raise RuntimeError("Attention.forward needs a CUDA tensor: the B200 path has no CPU fallback")
        if full_input.dim() != 4 or full_input.shape[1] != 3 or full_input.shape[2] != full_input.shape[3]:
            raise ValueError(f"expected a bag [N,3,S,S], got {tuple(full_input.shape)}")
        if self.precision not in DTYPE_CODES:
            raise ValueError(f"precision must be one of {list(DTYPE_CODES)}, got {self.precision!r}")
        bag = full_input.detach()                                                     # gbm/model.py:194,196
        if bag.dtype == torch.uint8:
            # raw 8-bit tiles: ToTensor() + Normalize(.5, .5) of the reference's loader (RoiBuilder.py:199-202) is
            # fused into the stem's load in bf16 mode; the fp32 check mode normalises here with the same fp32 ops
            bag = bag.contiguous()
            if self.precision != "bf16":
                bag = ((bag.float() / 255.0 - 0.5) / 0.5).contiguous()
        elif bag.dtype != torch.float32 or not bag.is_contiguous():
            bag = bag.float().contiguous()
        dev = bag.device
        if Y is None:
            Y = torch.tensor([1])                                                    # gbm/model.py:189 default
        Yl = Y.to(dev).long().reshape(-1)[:1].contiguous()                            # gbm/model.py:239
        idx = None
        drop = None
        group = self.bag_group
        if self.training:
            n_bag = bag.shape[0]
            si = self.subsample_indices
            if isinstance(si, torch.Tensor) and si.is_cuda:
                # a device-resident int32 index list is used as it is (no host round trip): what graph.GraphedStep
                # refills before every replay of a captured step
                if si.dtype != torch.int32 or not si.is_contiguous() or si.device != dev:
                    raise ValueError("a CUDA subsample_indices tensor must be contiguous int32 on the bag's device")
                idx = si
                n = int(si.numel())
                n_global = group.total(n, device=dev)
            else:
                if si is not None:
                    indices = torch.as_tensor(si).long().cpu()
                    n_global = group.total(int(indices.numel()), device=dev)
                else:
                    indices, n_global = group.subsample(n_bag, SUBSAMPLE, device=dev)     # gbm/model.py:193
                idx = indices.to(torch.int32).to(dev, non_blocking=True)
                n = int(indices.numel())
            if self.drop_mask is not None:
                drop = torch.as_tensor(self.drop_mask, dtype=torch.float32).to(dev).contiguous()
                if tuple(drop.shape) != (n, 80):
                    raise ValueError(f"drop_mask must be [{n},80], got {tuple(drop.shape)}")
            else:
                drop = (torch.rand((n, 80), device=dev) >= DROP_P).float()           # Dropout(0.25), gbm/model.py:107
        else:
            n_global = group.total(int(bag.shape[0]), hint=bag_tiles, device=dev)
        return bag, Yl, idx, drop, n_global

    # ---- the reference surface ----
    def _params(self):
        """The 65 parameters in state-dict order.  Walking the module tree costs ~0.3 ms per call, a tenth of a
        256-tile step, so the list is kept and only re-validated (conversions replace .data, not the Parameter)."""
        c = self._param_list_cache
        if c is None or c[0] is not self.weight_mask or c[-1] is not self.buffer.classifier.bias:
            _ = self._param_table            # checks names / shapes against the library once
            c = self._param_list_cache = [p for _, p in self.named_parameters()]
        return c

    def forward(self, full_input: torch.Tensor, Y: Optional[torch.Tensor] = None, bag_tiles: Optional[int] = None):
        """The reference's `forward(full_input, Y) -> dict` (gbm/model.py:189-264).  `bag_tiles` (extra, optional):
        total tiles of the whole bag when it is sharded over a BagGroup and the caller knows the number -- saves the
        all-gather of the shard sizes in eval mode."""
        bag, Yl, idx, drop, n_global = self._prepare(full_input, Y, bag_tiles)
        params = self._params()
        (loss, A, wroi, b, M, H, amu, avar, kld, ypred, yhat, err) = _MilFunction.apply(
            self, bag, Yl, idx, drop, n_global, *params)
        # Classifier penalty (gbm/model.py:246): two tiny norms, kept in autograd like the reference
        l2 = (self.buffer.lin1.weight.norm() + self.buffer.classifier.weight.norm()) * 0.5
        return {
            'Aterm': A, 'wROIs': wroi, 'Bterm': b, 'Mterm': M, 'Fterm': H, 'Aterm_mu': amu, 'Aterm_var': avar,
            'loss': loss, 'l2': l2, 'KLD': kld, 'y_pred': ypred, 'y_pred_hat': yhat, 'error': err,
        }

    def logits_and_attention(self, full_input, Y=None):
        """North-star view `forward(bag) -> (logits [1,3], attention weights [3,N])`."""
        out = self.forward(full_input, Y)
        return out['Mterm'].view(1, self.O * self.K), out['Aterm']

    @torch.no_grad()
    def features(self, full_input: torch.Tensor) -> torch.Tensor:
        """Extractor only: bag [N,3,S,S] -> Fterm [N,80] (no subsample, no head)."""
        lib = _lib.load()
        if not full_input.is_cuda:
            raise RuntimeError("features() needs a CUDA tensor: the B200 path has no CPU fallback")
        bag = full_input.detach()
        if bag.dtype == torch.uint8 and self.precision == "bf16":
            bag = bag.contiguous()          # raw 8-bit tiles: normalised inside the stem's load
        elif bag.dtype == torch.uint8:
            bag = ((bag.float() / 255.0 - 0.5) / 0.5).contiguous()
        else:
            bag = bag.float().contiguous()
        params = self._params()
        for nm, p in zip(self._param_names, params):
            _check_param(nm, p)
        pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        n, side, dt = int(bag.shape[0]), int(bag.shape[2]), DTYPE_CODES[self.precision]
        nbytes = int(lib.mil_extractor_infer_workspace_bytes(n, side, dt))
        ws = self._pool.acquire((n, side, dt, bag.device.index, False), nbytes, bag.device)
        try:
            with torch.cuda.device(bag.device):
                H = torch.empty((n, 80), dtype=torch.float32, device=bag.device)
                _lib.check(lib.mil_extractor_infer(pp, _ptr(bag), int(bag.dtype == torch.uint8), None, n, side, dt,
                                                   _ptr(ws.buf), nbytes, _ptr(H), _stream(bag.device)),
                           "mil_extractor_infer")
        finally:
            _WorkspacePool.release(ws)
        return H
