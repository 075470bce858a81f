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

from typing import List, Optional

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
        self._comm_stream = None

    # True: the ranks hold shards of ONE bag (AR-1..3 couple them).  SlideGroup overrides this.
    shares_bag = True

    # ---- shard bookkeeping -------------------------------------------------------------------------
    def shard_sizes(self, n_local_bag: int, device=None) -> List[int]:
        """Bag sizes of every rank: ONE all-gather of 8 bytes per call, entered by every rank unconditionally.
        (Nothing is cached: real slides differ in size from bag to bag, and a rank that skipped the collective
        because ITS shard happened to keep its size would leave the others waiting in it.)"""
        if self.world == 1:
            return [int(n_local_bag)]
        import torch.distributed as dist
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        else:
            dev = torch.device("cpu")
        mine = torch.tensor([n_local_bag], dtype=torch.int64, device=dev)
        out = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        return [int(v) for v in out.tolist()]

    def subsample(self, n_local_bag: int, frac: float, device=None, sizes: Optional[List[int]] = None):
        """Train-mode tile subset (gbm/model.py:193): `randperm(N)[:int(N*frac)]` over the WHOLE bag.  Returns
        (local indices of the chosen tiles that live in this rank's shard, in permutation order; number of tiles
        chosen over the whole bag).  `sizes` injects the per-rank shard sizes (tests)."""
        if self.world == 1:
            idx = torch.randperm(n_local_bag)[: int(n_local_bag * frac)]
            return idx, int(idx.numel())
        if sizes is None:
            sizes = self.shard_sizes(n_local_bag, device)
        total = sum(sizes)
        lo = sum(sizes[: self.rank])
        perm = torch.randperm(total, generator=self._gen)[: int(total * frac)]
        # every rank sees the same permutation, so every rank reaches the same verdict here (nobody is left waiting
        # in a collective): a shard without a single chosen tile has nothing to launch the kernels on
        bounds = torch.tensor([0] + sizes).cumsum(0)
        counts = torch.bucketize(perm, bounds[1:], right=True).bincount(minlength=self.world)
        if int(counts.min()) == 0 and perm.numel() > 0:
            raise ValueError(f"train-mode subsample of {int(perm.numel())} tiles leaves rank(s) "
                             f"{[r for r in range(self.world) if int(counts[r]) == 0]} without a tile; "
                             f"shard bags this small over fewer ranks")
        mine = perm[(perm >= lo) & (perm < lo + sizes[self.rank])] - lo
        return mine, int(perm.numel())

    def total(self, n_local: int, hint: Optional[int] = None, device=None) -> int:
        """Number of tiles the head sees over the whole bag.  `hint`: the caller already knows it (a loader that
        shards a slide does) -- saves the all-gather and its host synchronisation."""
        if self.world == 1:
            return int(n_local)
        if hint is not None:
            return int(hint)
        return sum(self.shard_sizes(n_local, device))

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


class SlideGroup(BagGroup):
    """Multi-slide data parallelism (BASELINE configs[4]; the reference accumulates five slides' gradients before an
    optimizer step, gbm/classify_combined.py:446-454): every rank holds its OWN bag.  The bag-wide sums AR-1..3 stay
    local -- mixing BatchNorm1d / pooling statistics of different slides would be wrong -- and only AR-4, the bucketed
    all-reduce (sum over the slides) of the parameter gradients overlapped with backward, remains."""

    shares_bag = False

    def shard_sizes(self, n_local_bag: int, device=None) -> List[int]:
        return [int(n_local_bag)]

    def subsample(self, n_local_bag: int, frac: float, device=None, sizes=None):
        idx = torch.randperm(n_local_bag)[: int(n_local_bag * frac)]      # the reference's own draw, per slide
        return idx, int(idx.numel())

    def total(self, n_local: int, hint: Optional[int] = None, device=None) -> int:
        return int(n_local)

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        return None
