"""One whole-slide bag on ONE GPU (default 32 768 tiles x 224^2, 8-bit tiles): fwd + bwd, tiles/s, and the size-independent
property of tests/test_gpu_parity.py at this size -- the features of the first 64 tiles are bit-identical to those of a
64-tile bag.   usage: python tools/big_bag_probe.py [tiles] [side]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = mil.Attention(n_classes=3).to(dev).eval()
base = bench.make_device_bag(mil, 256, side, dev, seed=1)
base8 = ((base * 0.5 + 0.5) * 255).round().clamp_(0, 255).to(torch.uint8)
bag = base8.repeat((n + 255) // 256, 1, 1, 1)[:n].contiguous()
Y = torch.tensor([1], device=dev)
print(f"bag {tuple(bag.shape)} uint8 = {bag.numel() / 1e9:.2f} GB", flush=True)


def step():
    net.zero_grad(set_to_none=True)
    out = net(bag, Y)
    out["loss"].backward()
    return out


out = step()
torch.cuda.synchronize()
print(f"peak device memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB, loss {float(out['loss']):.6f}", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{n} tiles x {side}^2 on one GPU: {ms:.2f} ms per fwd+bwd step = {n / ms * 1e3:,.0f} tiles/s")
F_big = out["Fterm"][:64].clone()
A_big = out["Aterm"].clone()
g_big = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
assert torch.isfinite(g_big).all() and torch.isfinite(A_big).all()
with torch.no_grad():
    F_small = net(bag[:64].contiguous(), Y)["Fterm"]
same = bool((F_big == F_small).all())
print(f"features of the first 64 tiles bit-identical to a 64-tile bag: {same};  sum(Aterm) per class = "
      f"{[round(float(v), 6) for v in A_big.sum(1)]};  |grad| = {float(g_big.norm()):.4e}")
# the bag is 128 copies of the same 256 tiles: attention weights must repeat with period 256
per = A_big[:, :256]
rep = bool(torch.allclose(A_big[:, 256 * 5: 256 * 6], per, rtol=0, atol=0))
print(f"attention weights periodic with the bag's 256-tile period (bitwise): {rep}")
assert same
