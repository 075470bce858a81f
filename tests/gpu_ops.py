"""Thin test-side wrappers over the layer-level C ABI (mil_conv_pf8, mil_conv_wgrad_pf8, mil_to_pf8 ...).
Used by the -m gpu parity tests and by tools/gpu_check.py.  Everything goes through ctypes -> libmil_b200.so."""
import ctypes as C
import importlib

import torch

PKG = "deep-convolutional-neural-network-resnet-26-and-attention-network_b200"
DT = {"fp32": 0, "bf16": 1}


def pkg():
    return importlib.import_module(PKG)


def lib():
    return pkg()._lib.load()


def check(rc, what):
    pkg()._lib.check(rc, what)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PF8:
    """A PF8 activation buffer on the GPU."""

    def __init__(self, n, c, h, w, dtype):
        self.n, self.c, self.h, self.w, self.dtype = n, c, h, w, dtype
        self.buf = torch.zeros(int(lib().mil_pf8_bytes(n, c, h, w, DT[dtype])), dtype=torch.uint8, device="cuda")

    @staticmethod
    def from_nchw(x, dtype):
        x = x.float().contiguous().cuda()
        t = PF8(x.shape[0], x.shape[1], x.shape[2], x.shape[3], dtype)
        check(lib().mil_to_pf8(DT[dtype], _p(x), _p(t.buf), t.n, t.c, t.h, t.w, _s()), "mil_to_pf8")
        return t

    def to_nchw(self):
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device="cuda")
        check(lib().mil_from_pf8(DT[self.dtype], _p(self.buf), _p(out), self.n, self.c, self.h, self.w, _s()),
              "mil_from_pf8")
        return out

    def raw(self):
        """The whole buffer as a flat tensor of its element type (to check guards / pad pixels)."""
        return self.buf.view(torch.bfloat16 if self.dtype == "bf16" else torch.float32)


def conv(x: PF8, w, bias=None, res: PF8 = None, act: PF8 = None, stride=1, epi=0, transposed=False, out_hw=None,
         impl=0):
    """out = epilogue(conv(x, w)).  transposed=True: data gradient (x has the conv's output geometry)."""
    cout, cin, ks, _ = w.shape
    w = w.float().contiguous().cuda()
    if transposed:
        ho, wo = out_hw
        oc = cin
    else:
        ho = (x.h + 2 * (ks // 2) - ks) // stride + 1
        wo = (x.w + 2 * (ks // 2) - ks) // stride + 1
        oc = cout
    out = PF8(x.n, oc, ho, wo, x.dtype)
    if transposed:
        nbytes = int(lib().mil_conv_workspace_bytes(x.n, cin, ho, wo, cout, x.h, x.w, ks))
    else:
        nbytes = int(lib().mil_conv_workspace_bytes(x.n, cin, x.h, x.w, cout, ho, wo, ks))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    b = None if bias is None else bias.float().contiguous().cuda()
    check(lib().mil_conv_pf8(DT[x.dtype], impl, int(transposed), _p(x.buf), x.n, x.c, x.h, x.w, _p(w), cout, cin, ks,
                             stride, _p(b), _p(res.buf) if res else None, _p(act.buf) if act else None, _p(out.buf),
                             ho, wo, epi, _p(ws), nbytes, _s()), "mil_conv_pf8")
    return out


def upsample2(x: PF8, ho, wo):
    """Zero-stuffed copy at the stride-2 convolution's input resolution (bf16 only)."""
    out = PF8(x.n, x.c, ho, wo, x.dtype)
    check(lib().mil_upsample2_pf8(_p(x.buf), x.n, x.c, x.h, x.w, _p(out.buf), ho, wo, _s()), "mil_upsample2_pf8")
    return out


def wgrad(x: PF8, dz: PF8, ks, stride, with_bias=True, impl=0):
    cin, cout = x.c, dz.c
    dw = torch.zeros((cout, cin, ks, ks), dtype=torch.float32, device="cuda")
    db = torch.zeros(cout, dtype=torch.float32, device="cuda") if with_bias else None
    nbytes = int(lib().mil_conv_workspace_bytes(x.n, cin, x.h, x.w, cout, dz.h, dz.w, ks))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    check(lib().mil_conv_wgrad_pf8(DT[x.dtype], impl, _p(x.buf), x.n, cin, x.h, x.w, _p(dz.buf), cout, dz.h, dz.w,
                                   ks, stride, _p(dw), _p(db), _p(ws), nbytes, _s()), "mil_conv_wgrad_pf8")
    return dw, db


def read_activation(net, n, side, which):
    """Saved activation `which` of the most recent forward of `net` (see mil_extractor_read_activation)."""
    mil = pkg()
    dt = mil.model.DTYPE_CODES[net.precision]
    key = (n, side, dt, torch.cuda.current_device(), True)      # the workspace of a forward that kept its activations
    ws = net._pool._items[key][0]
    # geometry
    hc = (side - 1) // 2 + 1
    h = [(hc - 1) // 2 + 1]
    for _ in range(3):
        h.append((h[-1] - 1) // 2 + 1)
    if which < 0:
        c, hh = 20, h[0]
    else:
        layer = (which // 2) // 3
        c, hh = (20, 40, 60, 80)[layer], h[layer]
    out = torch.empty((n, c, hh, hh), dtype=torch.float32, device="cuda")
    check(lib().mil_extractor_read_activation(n, side, dt, _p(ws.buf), which, _p(out), _s()),
          "mil_extractor_read_activation")
    return out


def relerr(a, b, atol=1e-7):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    diff = float((a - b).abs().max()) if a.numel() else 0.0
    if diff <= atol:
        return 0.0
    den = float(b.abs().max())
    return diff / (den if den > 0 else 1.0)


def extractor_forward_backward(params, bag, dH, precision, idx=None):
    """mil_extractor_forward + mil_extractor_backward through the C ABI with a CALLER-CHOSEN upstream gradient dH
    [n,80]: the extractor's kernels without the head in front of them (whose bag-wide BatchNorm1d backward makes the
    whole-path gradients of small bags an ill-conditioned difference of large numbers).
    params: dict name -> tensor (state-dict order is taken from the library).  Returns (H, {name: gradient})."""
    import ctypes
    mil = pkg()
    L = lib()
    table = mil._lib.param_table()
    dev = torch.device("cuda")
    ps = [params[nm].detach().float().contiguous().to(dev) for nm, _, _ in table]
    pp = (ctypes.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
    dt = DT[precision]
    bag = bag.contiguous().to(dev)
    n = int(bag.shape[0]) if idx is None else int(idx.numel())
    side = int(bag.shape[2])
    nbytes = int(L.mil_extractor_workspace_bytes(n, side, dt))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    H = torch.empty((n, 80), dtype=torch.float32, device=dev)
    ix = None if idx is None else idx.to(torch.int32).to(dev)
    fwd = L.mil_extractor_forward_u8 if bag.dtype == torch.uint8 else L.mil_extractor_forward
    check(fwd(pp, _p(bag), _p(ix), n, side, dt, _p(ws), nbytes, _p(H), _s()), "mil_extractor_forward")
    grads = torch.zeros(int(L.mil_param_total()), dtype=torch.float32, device=dev)
    dH = dH.float().contiguous().to(dev)
    check(L.mil_extractor_backward(pp, _p(bag), _p(ix), n, side, dt, _p(ws), nbytes, _p(dH), _p(grads), _s()),
          "mil_extractor_backward")
    torch.cuda.synchronize()
    out = {}
    for nm, shape, off in table:
        if nm.startswith("cnn."):
            numel = 1
            for d in shape:
                numel *= d
            out[nm] = grads[off:off + numel].view(shape).clone()
    return H, out


class Stem:
    """Layer-level stem (mil_stem_forward / mil_stem_backward): keeps the workspace between the two calls."""

    def __init__(self, bag, w, b, dtype, impl=0):
        self.bag = bag.float().contiguous().cuda()
        self.w, self.b = w.float().contiguous().cuda(), b.float().contiguous().cuda()
        self.dtype, self.impl = dtype, impl
        self.n, self.side = int(bag.shape[0]), int(bag.shape[2])
        nbytes = int(lib().mil_stem_workspace_bytes(self.n, self.side, DT[dtype]))
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        hc = (self.side - 1) // 2 + 1
        self.h0 = (hc - 1) // 2 + 1

    def forward(self) -> PF8:
        out = PF8(self.n, 20, self.h0, self.h0, self.dtype)
        check(lib().mil_stem_forward(DT[self.dtype], self.impl, _p(self.bag), self.n, self.side, _p(self.w), _p(self.b),
                                     _p(out.buf), _p(self.ws), self.ws.numel(), _s()), "mil_stem_forward")
        return out

    def backward(self, g: PF8):
        dw = torch.zeros((20, 3, 7, 7), dtype=torch.float32, device="cuda")
        db = torch.zeros(20, dtype=torch.float32, device="cuda")
        check(lib().mil_stem_backward(DT[self.dtype], self.impl, _p(self.bag), self.n, self.side, _p(g.buf), _p(dw), _p(db),
                                      _p(self.ws), self.ws.numel(), _s()), "mil_stem_backward")
        return dw, db
