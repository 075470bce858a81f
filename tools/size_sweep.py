"""Runs one fwd+bwd step at the given bag size and synchronises after each half (finding the smallest failing size of a
kernel change):   python tools/size_sweep.py <tiles> [side]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n = int(sys.argv[1])
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = mil.Attention(n_classes=3).to(dev).eval()
bag = bench.make_device_bag(mil, n, side, dev, seed=1)
Y = torch.tensor([1], device=dev)
for rep in range(int(os.environ.get("SWEEP_REPS", 3))):
    net.zero_grad(set_to_none=True)
    out = net(bag, Y)
    torch.cuda.synchronize()
    print(n, rep, "forward ok", float(out["loss"]), flush=True)
    out["loss"].backward()
    torch.cuda.synchronize()
    print(n, rep, "backward ok", float(sum(p.grad.abs().sum() for p in net.parameters())), flush=True)
