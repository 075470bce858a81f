"""Host -> device staging of bags: double-buffered asynchronous copies on a side stream.

The reference moves one slide per step with a blocking `batch[0].squeeze(0).cuda()` (gbm/classify_combined.py:423)
-- 2.5 GB of fp32 for a 4,096-tile bag, ~45 ms over PCIe, during which the GPU idles.  `BagStager` lets the caller
(a training loop, or a DataLoader wrapper) submit the NEXT bag while the current one is being processed:

    stager = BagStager(device)
    ticket = stager.submit(next_host_bag)          # pinned host tensor; returns immediately
    ...
    bag = stager.get(ticket)                       # device tensor; the current stream waits for the copy only
    out = classifier(bag, label); out['loss'].backward()
    stager.release(ticket)                         # the buffer may be overwritten once this step's kernels ran

PyTorch plumbing only (streams, events, pinned memory); no arithmetic happens here.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class _Slot:
    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.copied = torch.cuda.Event()
        self.consumed: Optional[torch.cuda.Event] = None
        self.busy = False
        self.view: Optional[torch.Tensor] = None


class BagStager:
    def __init__(self, device=None, n_buffers: int = 2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots: List[_Slot] = [_Slot() for _ in range(max(2, n_buffers))]

    def submit(self, host_bag: torch.Tensor) -> int:
        """Start copying `host_bag` (ideally pinned) into a free device buffer; returns a ticket."""
        if host_bag.is_cuda:
            raise ValueError("BagStager.submit expects a host tensor")
        for i, s in enumerate(self.slots):
            if not s.busy:
                break
        else:
            raise RuntimeError("all staging buffers are in flight: release() a ticket first")
        s.busy = True
        with torch.cuda.stream(self.stream):
            if s.buf is None or s.buf.numel() < host_bag.numel() or s.buf.dtype != host_bag.dtype:
                # allocated ON the side stream: the caching allocator then never hands out a block whose last use is
                # still queued on another stream (a block just freed on the main stream could otherwise be
                # overwritten by this copy underneath kernels that have not run yet)
                s.buf = torch.empty(host_bag.numel(), dtype=host_bag.dtype, device=self.device)
            s.view = s.buf[: host_bag.numel()].view(host_bag.shape)
            if s.consumed is not None:
                self.stream.wait_event(s.consumed)      # previous user of this buffer has finished reading it
            s.view.copy_(host_bag, non_blocking=True)
            s.copied.record(self.stream)
        return i

    def get(self, ticket: int) -> torch.Tensor:
        """The staged bag; work queued on the current stream after this call sees the finished copy."""
        s = self.slots[ticket]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(s.copied)
        s.view.record_stream(cur)       # the buffer lives on the side stream's pool: tell the allocator who reads it
        return s.view

    def release(self, ticket: int) -> None:
        """Call after the last kernel that reads the staged bag has been enqueued on the current stream."""
        s = self.slots[ticket]
        s.consumed = torch.cuda.Event()
        s.consumed.record(torch.cuda.current_stream(self.device))
        s.busy = False
