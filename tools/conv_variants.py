"""Times the tcgen05 convolution / weight-gradient kernels alone at every layer shape of the network, per epilogue
kind (CUDA events, 20 launches after warm-up):   python tools/conv_variants.py [tiles]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import gpu_ops as G  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
HBM = 6530.3


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


from tests.gpu_ops import DT, _p, _s, check, lib  # noqa: E402

for c, h in ((20, 56), (40, 28), (60, 14), (80, 7)):
    x = torch.randn(n, c, h, h, device="cuda")
    w = (torch.randn(c, c, 3, 3, device="cuda") * 0.1).float().contiguous()
    b = torch.zeros(c, device="cuda")
    X = G.PF8.from_nchw(x, "bf16")
    R = G.PF8.from_nchw(torch.randn_like(x), "bf16")
    A = G.PF8.from_nchw(torch.randn_like(x), "bf16")
    O = G.PF8(n, c, h, h, "bf16")
    nbytes = int(lib().mil_conv_workspace_bytes(n, c, h, h, c, h, h, 3))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    dw = torch.zeros((c, c, 3, 3), device="cuda")
    db = torch.zeros(c, device="cuda")
    mapb = n * c * h * h * 2 / 1e6     # MB per map, un-padded

    def conv(tr, bias, res, act, epi):
        check(lib().mil_conv_pf8(DT["bf16"], 2, tr, _p(X.buf), n, c, h, h, _p(w), c, c, 3, 1, _p(bias),
                                 _p(res.buf) if res else None, _p(act.buf) if act else None, _p(O.buf), h, h, epi,
                                 _p(ws), nbytes, _s()), "mil_conv_pf8")

    def wgrad():
        check(lib().mil_conv_wgrad_pf8(DT["bf16"], 2, _p(X.buf), n, c, h, h, _p(R.buf), c, h, h, 3, 1, _p(dw), _p(db),
                                       _p(ws), nbytes, _s()), "mil_conv_wgrad_pf8")

    rows = [
        ("fwd+res ", lambda: conv(0, b, R, None, 0), 3),
        ("fwd     ", lambda: conv(0, b, None, None, 0), 2),
        ("plain   ", lambda: conv(0, None, None, None, 2), 2),
        ("dgrad   ", lambda: conv(1, None, None, A, 1), 3),
        ("dgrad+r ", lambda: conv(1, None, R, A, 1), 4),
        ("wgrad   ", wgrad, 2),
    ]
    for name, fn, maps in rows:
        us = timed(fn)
        print(f"c={c:2d} h={h:2d} tiles={n} {name}: {us:8.1f} us   {maps * mapb / us:7.2f} TB/s algorithmic "
              f"({maps * mapb / us * 1e3 / HBM * 100:5.1f} % of HBM)   "
              f"{2 * 9 * c * c * h * h * n / us / 1e6:7.1f} TFLOP/s")
