"""One launch of every heavy wide-channel kernel at the layer shapes of the alt_resnet parameterisation (the target of the
`ncu --set full` capture summarised in profiles/r2_ncu_wide_kernels.txt).   usage: python tools/wide_ncu_target.py [tiles]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_ops as G  # noqa: E402
from tests.test_gpu_wide import wide_conv, wide_wgrad  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for C_, H in ((64, 56), (128, 28), (256, 14), (512, 7)):
    X = G.PF8.from_nchw(torch.randn(n, C_, H, H, device="cuda"), "bf16")
    w = torch.randn(C_, C_, 3, 3, device="cuda") / (C_ * 9) ** 0.5
    wide_conv(X, w, res=X, epi=0)                                  # forward + identity + ReLU
    wide_conv(X, w, transposed=True, res=X, act=X, epi=1)          # data gradient + identity gradient, ReLU' mask
    wide_wgrad(X, X, 3, (C_, C_, 3, 3))                            # weight gradient
    torch.cuda.synchronize()
    del X
print("ok")
