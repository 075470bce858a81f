"""Host -> device copy rate per GPU when 1, 2, 4, ... GPUs of the box pull from pinned host memory at the same time
(one process, one pinned buffer + one stream per GPU, CUDA events).  Explains the end-to-end scaling of bench.py:
the `e2e` figure copies every step's bag from the host (2.47 GB of fp32 tiles, or 0.62 GB of 8-bit tiles, per 4096-tile
bag and GPU).   usage: python tools/h2d_rates.py [MB per copy] [repeats]"""
import sys

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_dev = torch.cuda.device_count()
host = [torch.empty(mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
devb = [torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n_dev)]
streams = [torch.cuda.Stream(device=i) for i in range(n_dev)]
print(f"# {n_dev} GPUs, {mb} MiB per copy, {reps} copies per GPU and measurement, pinned host memory")
k = 1
while k <= n_dev:
    for first in sorted({0, n_dev - k}):
        ids = list(range(first, first + k))
        ev = {}
        for i in ids:
            torch.cuda.synchronize(i)
        for i in ids:
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    devb[i].copy_(host[i], non_blocking=True)
                e1.record()
                ev[i] = (e0, e1)
        rates = []
        for i in ids:
            torch.cuda.synchronize(i)
            rates.append(reps * (mb << 20) / (ev[i][0].elapsed_time(ev[i][1]) * 1e-3) / 1e9)
        print(f"GPUs {ids[0]}..{ids[-1]} ({k} at once): per GPU " + " ".join(f"{r:5.1f}" for r in rates) +
              f" GB/s   sum {sum(rates):6.1f} GB/s")
    k *= 2
