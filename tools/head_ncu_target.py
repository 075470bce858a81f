"""The MIL head alone at a whole-slide bag size (default 32 768 tiles, BASELINE configs[2] on one GPU): forward + backward of
`Attention` on tiny 32 x 32 tiles, so that the head kernels (mil_head.cu) see their real N while the extractor stays
small.  Target of the `ncu --set full -k regex:head_` capture summarised in profiles/r2_ncu_head.txt.
usage: python tools/head_ncu_target.py [tiles]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = mil.Attention(n_classes=3).to(dev).eval()
bag = torch.rand((n, 3, 32, 32), device=dev) * 2 - 1
Y = torch.tensor([1], device=dev)
for _ in range(2):
    net.zero_grad(set_to_none=True)
    out = net(bag, Y)
    out["loss"].backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print("ok", n, float(out["loss"]))
