"""Where does the end-to-end loop on 8-bit tiles lose time against the device-resident step?  Times, on one GPU:
(a) resident fp32 bag, (b) resident uint8 bag, (c) uint8 bag copied from pinned host memory every step (BagStager, no result
read-back), (d) the same with the per-step result read-back of bench.py's e2e loop.   usage: python tools/e2e_gap.py [tiles]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
net = mil.Attention(n_classes=3).to(dev).eval()
bag = bench.make_device_bag(mil, n, 224, dev, seed=1)
u8 = ((bag + 1.0) * 127.5).round_().clamp_(0, 255).to(torch.uint8)
host_u8 = u8.cpu().pin_memory()
Y = torch.tensor([1], device=dev)
steps = 10


def step(b):
    net.zero_grad(set_to_none=True)
    out = net(b, Y)
    out["loss"].backward()
    return out


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn_n = fn
    for _ in range(steps):
        fn_n()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print(f"(a) resident fp32 bag: {timed(lambda: step(bag)):.3f} ms/step")
print(f"(b) resident uint8 bag: {timed(lambda: step(u8)):.3f} ms/step")
stager = mil.BagStager(dev)
state = {"ticket": stager.submit(host_u8)}


def e2e(readback):
    nxt = stager.submit(host_u8)
    o = step(stager.get(state["ticket"]))
    if readback:
        r = o["loss"].detach().reshape(-1).to("cpu", non_blocking=True)
    stager.release(state["ticket"])
    state["ticket"] = nxt


print(f"(c) uint8 bag from pinned host memory every step, no read-back: {timed(lambda: e2e(False)):.3f} ms/step")
print(f"(d) ... with the loss copied back every step: {timed(lambda: e2e(True)):.3f} ms/step")
# (e) the copy alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dst = torch.empty_like(u8)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    dst.copy_(host_u8, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"(e) the host -> device copy alone: {ms:.3f} ms = {host_u8.numel() / ms / 1e6:.1f} GB/s")
