"""Host-side cost of one training step (how long `forward + backward` takes to ENQUEUE, GPU work excluded) next to
the device time of the step:   python tools/cpu_overhead.py [tiles]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

mil = importlib.import_module("deep-convolutional-neural-network-resnet-26-and-attention-network_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = mil.Attention(n_classes=3).cuda().eval()
bag = torch.rand((n, 3, 224, 224), device="cuda") * 2 - 1
Y = torch.tensor([1]).cuda()


def step():
    net.zero_grad(set_to_none=True)
    out = net(bag, Y)
    out["loss"].backward()
    return out


for _ in range(3):
    step()
torch.cuda.synchronize()
host, dev = [], []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    dev.append(e0.elapsed_time(e1))
print(f"tiles {n}: host enqueue {1e3 * sorted(host)[len(host) // 2]:.2f} ms per step, device {sorted(dev)[len(dev) // 2]:.2f} ms per step")

import cProfile  # noqa: E402
import pstats  # noqa: E402

pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
