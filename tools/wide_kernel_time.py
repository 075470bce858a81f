"""Times the wide-channel kernels (csrc/mil_wide_conv.cu, mil_wide_wgrad.cu) alone at the layer shapes of the
alt_resnet parameterisation (224x224 tiles: 64 ch @ 56^2, 128 @ 28^2, 256 @ 14^2, 512 @ 7^2) and prints TFLOP/s against
the measured dense bf16 peak.   usage: python tools/wide_kernel_time.py [tiles] [reps]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_ops as G  # noqa: E402
from tests.test_gpu_wide import wide_conv, wide_wgrad  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peak = 1651.7
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak = float(pk.get("bf16_tflops", peak))
except Exception:
    pass


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us (includes the weight-pack launch of the layer-level entry point)


print(f"# {n} tiles, peak {peak:.0f} TFLOP/s; times include the (small) weight pack / reduction launches")
for C_, H in ((64, 56), (128, 28), (256, 14), (512, 7)):
    X = G.PF8(n, C_, H, H, "bf16")
    X.buf.view(torch.bfloat16)[:] = 0
    x = torch.randn(min(n, 64), C_, H, H, device="cuda")
    X = G.PF8.from_nchw(torch.randn(n, C_, H, H, device="cuda") if n * C_ * H * H < 2 ** 31 else x, "bf16")
    w = torch.randn(C_, C_, 3, 3, device="cuda") / (C_ * 9) ** 0.5
    flop = 2.0 * X.n * H * H * C_ * C_ * 9
    for tm in (0, 2, 1):
        t = timed(lambda: wide_conv(X, w, res=X, epi=0, tm=tm))
        print(f"conv3x3 {C_:4d}ch {H:3d}^2 fwd+res tm={tm}: {t:8.1f} us  {flop / t / 1e6:7.1f} TFLOP/s  {100 * flop / t / 1e6 / peak:5.1f} %")
    t = timed(lambda: wide_conv(X, w, transposed=True, res=X, act=X, epi=1))
    print(f"conv3x3 {C_:4d}ch {H:3d}^2 dgrad+res    : {t:8.1f} us  {flop / t / 1e6:7.1f} TFLOP/s  {100 * flop / t / 1e6 / peak:5.1f} %")
    if C_ >= 128:
        t = timed(lambda: wide_wgrad(X, X, 3, (C_, C_, 3, 3)))
        print(f"wgrad3x3 {C_:4d}ch {H:3d}^2            : {t:8.1f} us  {flop / t / 1e6:7.1f} TFLOP/s  {100 * flop / t / 1e6 / peak:5.1f} %")
    del X
    torch.cuda.empty_cache()
