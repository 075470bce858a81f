"""Layer-3 shaped convolution (60 -> 60 @ 14x14, forward with residual) alone, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -s 3 python tools/profile_l3.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import gpu_ops as G  # noqa: E402
from tests.gpu_ops import DT, _p, _s, check, lib  # noqa: E402

n, c, h = int(os.environ.get("PROF_TILES", 4096)), int(os.environ.get("PROF_C", 60)), int(os.environ.get("PROF_H", 14))
x = torch.randn(n, c, h, h, device="cuda")
w = (torch.randn(c, c, 3, 3, device="cuda") * 0.05).contiguous()
b = torch.zeros(c, device="cuda")
X = G.PF8.from_nchw(x, "bf16")
R = G.PF8.from_nchw(torch.randn_like(x), "bf16")
O = G.PF8(n, c, h, h, "bf16")
nb = int(lib().mil_conv_workspace_bytes(n, c, h, h, c, h, h, 3))
ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
for _ in range(5):
    check(lib().mil_conv_pf8(DT["bf16"], 2, 0, _p(X.buf), n, c, h, h, _p(w), c, c, 3, 1, _p(b), _p(R.buf), None, _p(O.buf),
                             h, h, 0, _p(ws), nb, _s()), "conv")
torch.cuda.synchronize()
print("ok")
