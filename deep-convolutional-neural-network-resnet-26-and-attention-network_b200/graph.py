"""CUDA-graph replay of a whole training step for bags of one fixed size.

The reference's live pipeline caps a bag at 2 500 tiles and pushes 20 % of them through the CNN (RoiBuilder.py:230,
gbm/model.py:193): ~500 tiles, 2-3 ms of device work spread over ~165 kernel launches -- at that size the step is
bound by launch latency and host enqueue time, not by the kernels.  `GraphedStep` captures forward + backward
(+ the optimizer step) once into a CUDA graph and replays it: one host call per step, no Python between the kernels.
The library's entry points are graph-safe by construction (no allocation, no synchronisation, everything on the
caller's stream); the only host-side work that remains per step is the reference's own CPU `randperm` subsample
draw, whose result is copied into a static index buffer before the replay.
"""
from __future__ import annotations

from typing import Optional

import torch

from .model import SUBSAMPLE


class GraphedStep:
    """step = GraphedStep(net, n_bag, side, optimizer=opt);  out = step(bag, Y)

    `net`  an `Attention` module on a CUDA device, in train() or eval() mode (the mode is baked into the capture)
    `n_bag`, `side`, `bag_dtype`  the fixed bag shape [n_bag, 3, side, side] (float32 or uint8 tiles)
    `optimizer`  optional `FusedAdam` (flat buffers, one-launch step): its `zero_grad` / `step` become part of the graph
           (parameters and Adam state are untouched by the capture itself).  Without one the gradients of the step
           are left in `.grad`.
    The returned dict holds STATIC tensors: they are overwritten by the next call.  Sharded bags (BagGroup with more
    than one rank) are not captured (their collectives run on side streams): use the eager path there."""

    def __init__(self, net, n_bag: int, side: int, optimizer=None, bag_dtype=torch.float32, warmup: int = 3):
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs the module on a CUDA device (no CPU fallback)")
        if getattr(net.bag_group, "world", 1) != 1 and type(net.bag_group).__name__ == "BagGroup":
            raise RuntimeError("GraphedStep: a bag sharded over several ranks is not captured; use the eager path")
        if net.loss.weight is not None and not isinstance(net.loss.weight, torch.Tensor):
            # the class weights are uploaded once here: a host -> device copy cannot be part of a capture
            net.loss.weight = torch.as_tensor(net.loss.weight, dtype=torch.float32, device=dev)
        self.net, self.optimizer, self.device = net, optimizer, dev
        self.n_bag, self.side = int(n_bag), int(side)
        self.training = bool(net.training)
        self.bag = torch.zeros((n_bag, 3, side, side), dtype=bag_dtype, device=dev)
        self.Y = torch.ones(1, dtype=torch.long, device=dev)
        self.idx: Optional[torch.Tensor] = None
        self._idx_host: Optional[torch.Tensor] = None
        if self.training:
            k = int(n_bag * SUBSAMPLE)
            if k < 2:
                raise ValueError(f"a {n_bag}-tile bag leaves {k} tiles after the 20 % subsample: BatchNorm1d needs 2")
            self.idx = torch.arange(k, dtype=torch.int32, device=dev)
            self._idx_host = torch.empty(k, dtype=torch.int32).pin_memory()
        if optimizer is None and getattr(net, "_gflat", None) is not None:
            raise ValueError("this module's gradients accumulate into a FusedAdam flat buffer: pass that optimizer to "
                             "GraphedStep (its zero_grad / step become part of the graph)")
        self._fused = optimizer is not None and hasattr(optimizer, "hyper_values")
        if optimizer is not None and not self._fused:
            raise TypeError("GraphedStep captures FusedAdam (or no optimizer): the warm-up steps of another optimizer "
                            "would be real parameter updates")
        self._hyper_host = None
        if self._fused:
            # FusedAdam inside a graph: its six scalars (the bias corrections change every step) are read from device
            # memory, refreshed here before every replay
            self._hyper_host = torch.zeros(6, dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.zeros(6, dtype=torch.float32, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.out = None
        self._capture(warmup)

    def _one_step(self):
        net, opt = self.net, self.optimizer
        if opt is not None:
            opt.zero_grad()
        else:
            for p in net.parameters():
                p.grad = None
        out = net(self.bag, self.Y)
        out["loss"].backward()
        if opt is not None:
            opt.step()
        return out

    def _capture(self, warmup: int):
        net = self.net
        saved = net.subsample_indices
        if self.training:
            net.subsample_indices = self.idx
        state = None
        if self._fused:
            # the warm-up steps below must not move the parameters or the Adam state: a zero step size (and the
            # state restored afterwards)
            opt = self.optimizer
            state = (opt._flat.clone(), opt._m.clone(), opt._v.clone())
            opt._hyper = self._hyper_dev
            self._push_hyper(max(1, opt._t))
        try:
            side_stream = torch.cuda.Stream(device=self.device)
            side_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side_stream):          # warm-up off the default stream (allocator, smem attributes)
                for _ in range(max(1, warmup)):
                    self._one_step()
            torch.cuda.current_stream(self.device).wait_stream(side_stream)
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(self.graph):
                self.out = self._one_step()
            if state is not None:
                opt = self.optimizer
                opt._flat.copy_(state[0]); opt._m.copy_(state[1]); opt._v.copy_(state[2])
        finally:
            net.subsample_indices = saved
            if self._fused:
                self.optimizer._hyper = None

    def _push_hyper(self, t):
        self._hyper_host.copy_(torch.tensor(self.optimizer.hyper_values(t), dtype=torch.float32))
        self._hyper_dev.copy_(self._hyper_host, non_blocking=True)

    def __call__(self, bag: torch.Tensor, Y: Optional[torch.Tensor] = None):
        if tuple(bag.shape) != tuple(self.bag.shape) or bag.dtype != self.bag.dtype:
            raise ValueError(f"GraphedStep was captured for {tuple(self.bag.shape)} {self.bag.dtype}, "
                             f"got {tuple(bag.shape)} {bag.dtype}")
        if bool(self.net.training) != self.training:
            raise RuntimeError("the module's train/eval mode changed since the capture")
        if bag.data_ptr() != self.bag.data_ptr():
            self.bag.copy_(bag, non_blocking=True)
        if Y is not None:
            self.Y.copy_(Y.reshape(-1)[:1].to(torch.long), non_blocking=True)
        if self.training:
            # the reference's own draw (gbm/model.py:193), on the host like there; staged through pinned memory
            self._idx_host.copy_(torch.randperm(self.n_bag)[: self.idx.numel()].to(torch.int32))
            self.idx.copy_(self._idx_host, non_blocking=True)
        if self._fused:
            self.optimizer._t += 1
            self._push_hyper(self.optimizer._t)
        self.graph.replay()
        return self.out
