import os, sys, torch
sys.path.insert(0, '/root/repo')
from tests import gpu_ops as G
import ctypes as C
n, c, h = 1024, 20, 56
x = torch.randn(n, c, h, h, device="cuda"); w = torch.randn(c, c, 3, 3, device="cuda") * 0.1; b = torch.zeros(c, device="cuda")
X = G.PF8.from_nchw(x, "bf16")
for _ in range(3): out = G.conv(X, w, bias=b, res=X, stride=1, epi=0, impl=2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib = G.lib(); P = lambda t: C.c_void_p(t.data_ptr())
O = G.PF8(n, c, h, h, "bf16"); nb = int(lib.mil_conv_workspace_bytes(n, c, h, h, c, h, h, 3)); ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
def run(res=True):
    G.check(lib.mil_conv_pf8(1, 2, 0, P(X.buf), n, c, h, h, P(w), c, c, 3, 1, P(b), P(X.buf) if res else None, None, P(O.buf), h, h, 0, P(ws), nb, None), "conv")
for res in (True, False):
    run(res); torch.cuda.synchronize(); e0.record()
    for _ in range(10): run(res)
    e1.record(); torch.cuda.synchronize()
    print("MIL_TC_DBG=%s res=%s: %.1f us per launch" % (os.environ.get("MIL_TC_DBG", "0"), res, e0.elapsed_time(e1) * 100))
