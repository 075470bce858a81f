"""Can a thin layer's data-gradient convolution and weight gradient share an SM?  Times, at a layer-1 / layer-2 shape,
(a) the two kernels back to back on one stream in their normal forms and (b) their small-footprint forms
(mil_set_option("compact", 10 a + b)) launched concurrently on two streams.   usage: python tools/overlap_probe.py [tiles]"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_ops as G  # noqa: E402

lib = G.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = 6
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def probe(c, h):
    X = G.PF8.from_nchw(torch.randn(n, c, h, h, device="cuda"), "bf16")
    DZ = G.PF8.from_nchw(torch.randn(n, c, h, h, device="cuda"), "bf16")
    R = G.PF8.from_nchw(torch.randn(n, c, h, h, device="cuda"), "bf16")
    OUT = G.PF8(n, c, h, h, "bf16")
    w = (torch.randn(c, c, 3, 3, device="cuda") * 0.1).contiguous()
    dw = torch.zeros_like(w)
    db = torch.zeros(c, device="cuda")
    nbytes = int(lib.mil_conv_workspace_bytes(n, c, h, h, c, h, h, 3))
    ws1 = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    ws2 = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")

    def dgrad(st):
        G.check(lib.mil_conv_pf8(1, 2, 1, P(DZ.buf), n, c, h, h, P(w), c, c, 3, 1, None, P(R.buf), P(X.buf), P(OUT.buf), h, h,
                                 1, P(ws1), nbytes, C.c_void_p(st.cuda_stream)), "conv")

    def wgrad(st):
        G.check(lib.mil_conv_wgrad_pf8(1, 2, P(X.buf), n, c, h, h, P(DZ.buf), c, h, h, 3, 1, P(dw), P(db), P(ws2), nbytes,
                                       C.c_void_p(st.cuda_stream)), "wgrad")

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(concurrent):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(s1)
        s2.wait_event(e0)
        for _ in range(reps):
            dgrad(s1)
            wgrad(s2 if concurrent else s1)
        ej = torch.cuda.Event()
        ej.record(s2)
        s1.wait_event(ej)
        e1.record(s1)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    def alone(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(s1)
        for _ in range(reps):
            fn(s1)
        e1.record(s1)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    for opt in (0, 32, 34, 22, 24, 42):
        G.check(lib.mil_set_option(b"compact", opt), "set_option")
        for _ in range(2):
            run(False)
            run(True)
        print(f"{c:3d} ch @ {h}^2, {n} tiles, compact={opt:2d}: dgrad alone {alone(dgrad):7.1f} us, wgrad alone {alone(wgrad):7.1f} us, "
              f"serial pair {run(False):7.1f} us, two streams {run(True):7.1f} us", flush=True)
    G.check(lib.mil_set_option(b"compact", 0), "set_option")


probe(20, 56)
probe(40, 28)
