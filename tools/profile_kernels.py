"""Runs the two dominant tensor-core kernels alone, at the layer-1 shape of a 1024-tile bag (56x56x20), for ncu:

    python tools/profile_kernels.py && ncu --set full --clock-control none --import-source on \
        -k regex:'conv_tc_kernel|wgrad_tc_kernel' -c 4 -o gpurun_out/prof_tc python tools/profile_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import gpu_ops as G  # noqa: E402

n, c, h = int(os.environ.get("PROF_TILES", 1024)), 20, 56
x = torch.randn(n, c, h, h, device="cuda")
w = torch.randn(c, c, 3, 3, device="cuda") * 0.1
b = torch.zeros(c, device="cuda")
X = G.PF8.from_nchw(x, "bf16")
DZ = G.PF8.from_nchw(torch.randn(n, c, h, h, device="cuda"), "bf16")
R = G.PF8.from_nchw(torch.randn(n, c, h, h, device="cuda"), "bf16")          # residual: a separate map
for _ in range(2):
    out = G.conv(X, w, bias=b, res=R, stride=1, epi=0, impl=2)          # forward conv, full epilogue
    dw, db = G.wgrad(X, DZ, 3, 1, impl=2)                                # weight gradient
torch.cuda.synchronize()
print("ok", float(out.to_nchw().abs().mean()), float(dw.abs().mean()))
