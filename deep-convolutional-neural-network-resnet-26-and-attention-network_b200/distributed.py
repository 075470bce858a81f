"""Sharding one bag (slide) over the ranks of a process group -- one process per GPU, NCCL over NVLink.

The reference only has single-process `nn.DataParallel` over 4 GPUs for the CNN and runs the head on GPU 0
(gbm/model.py:132-135).  Here every rank owns a contiguous shard of the bag's tiles, runs the extractor AND the
head on its shard, and the ranks exchange only the bag-wide sums the head needs:

  AR-1  double[160]  sum_n x, sum_n x^2        BatchNorm1d batch statistics        (gbm/model.py:105-109)
  AR-2  double[16]   sum g, sum g*b, ...       L1-normalised attention pooling     (gbm/model.py:213,227)
  AR-3  double[160]  sum dHz, sum dHz*xhat     BatchNorm1d backward
  AR-4  float[640967] parameter gradients      (the reference's DataParallel reduce-add onto GPU 0)

The sums are exact (not approximations): every rank ends up with the same Mterm / logits / loss as a single
GPU would, up to fp64 summation order.  `BagGroup()` without a process group is the single-GPU no-op.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class BagGroup:
    def __init__(self, group=None, seed: int = 0, grad_buckets: int = 1):
        """group: a torch.distributed process group whose ranks share ONE bag, or None (single rank).
        seed: seed of the shared CPU generator used for the train-mode subsample when world > 1 (all ranks
        must draw the same permutation; with one rank the default torch RNG is used exactly like the reference)."""
        self.group = group
        self.world = 1
        self.rank = 0
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        self.grad_buckets = max(1, int(grad_buckets))
        self._gen = torch.Generator().manual_seed(int(seed)) if self.world > 1 else None
        self._sizes: Dict[int, List[int]] = {}
        self._n_global_pending: Optional[int] = None
        self._comm_stream = None

    # ---- shard bookkeeping -------------------------------------------------------------------------
    def shard_sizes(self, n_local_bag: int, device=None) -> List[int]:
        """Bag sizes of every rank (one all-gather per distinct local size, then cached)."""
        if self.world == 1:
            return [n_local_bag]
        if n_local_bag not in self._sizes:
            import torch.distributed as dist
            backend = dist.get_backend(self.group)
            if backend == "nccl":
                dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            else:
                dev = torch.device("cpu")
            mine = torch.tensor([n_local_bag], dtype=torch.int64, device=dev)
            out = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine, group=self.group)
            self._sizes[n_local_bag] = [int(t.item()) for t in out]
        return self._sizes[n_local_bag]

    def subsample(self, n_local_bag: int, frac: float, device=None) -> torch.Tensor:
        """Train-mode tile subset (gbm/model.py:193): `randperm(N)[:int(N*frac)]` over the WHOLE bag; returns
        the local indices of the chosen tiles that live in this rank's shard (order of the permutation kept)."""
        if self.world == 1:
            idx = torch.randperm(n_local_bag)[: int(n_local_bag * frac)]
            self._n_global_pending = int(idx.numel())
            return idx
        sizes = self.shard_sizes(n_local_bag, device)
        total = sum(sizes)
        lo = sum(sizes[: self.rank])
        perm = torch.randperm(total, generator=self._gen)[: int(total * frac)]
        self._n_global_pending = int(perm.numel())
        # every rank sees the same permutation, so every rank reaches the same verdict here (nobody is left waiting
        # in a collective): a shard without a single chosen tile has nothing to launch the kernels on
        bounds = torch.tensor([0] + sizes).cumsum(0)
        counts = torch.bucketize(perm, bounds[1:], right=True).bincount(minlength=self.world)
        if int(counts.min()) == 0 and perm.numel() > 0:
            raise ValueError(f"train-mode subsample of {int(perm.numel())} tiles leaves rank(s) "
                             f"{[r for r in range(self.world) if int(counts[r]) == 0]} without a tile; "
                             f"shard bags this small over fewer ranks")
        mine = perm[(perm >= lo) & (perm < lo + sizes[self.rank])] - lo
        return mine

    def total(self, n_local: int, n_local_bag: Optional[int] = None) -> int:
        """Number of tiles the head sees over the whole bag."""
        if self._n_global_pending is not None:      # set by subsample() for this forward
            n = self._n_global_pending
            self._n_global_pending = None
            return n
        if self.world == 1:
            return n_local
        return sum(self.shard_sizes(n_local))

    # ---- collectives ------------------------------------------------------------------------------
    def all_reduce_sum(self, t: torch.Tensor) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_grads(self, flat: torch.Tensor) -> None:
        """Sum the flat gradient buffer over the ranks of the bag (bag-level loss: a sum, not a mean)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        n = flat.numel()
        step = (n + self.grad_buckets - 1) // self.grad_buckets
        works = []
        for s in range(0, n, step):
            works.append(dist.all_reduce(flat[s:s + step], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()

    def all_reduce_grads_staged(self, flat: torch.Tensor, events, bounds) -> None:
        """Bucketed AR-4 overlapped with backward.  events[l] (l = 3..0) fires when layer l+1's gradients are final
        (mil_extractor_backward_staged); bounds = float offsets where layer2 / layer3 / layer4 start.  Buckets in
        completion order: [layer4 .. end] (incl. fc + head), [layer3], [layer2], [start .. layer1]."""
        if self.world == 1:
            return
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=flat.device)
        b2, b3, b4 = bounds
        n = flat.numel()
        buckets = [(events[3], b4, n), (events[2], b3, b4), (events[1], b2, b3), (events[0], 0, b2)]
        works = []
        with torch.cuda.stream(self._comm_stream):
            for ev, lo, hi in buckets:
                self._comm_stream.wait_event(ev)
                works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        cur = torch.cuda.current_stream(flat.device)
        for w in works:
            w.wait()                      # makes the current stream wait for the collective
        cur.wait_stream(self._comm_stream)
