"""-m gpu tests of the wide-channel kernel family (the alt_resnet.py parameterisation, 64 .. 512
channels, ReLU, no convolution bias -- alt_resnet.py:24-32,35-67,70-145) against torch's fp32 GPU convolutions on the
same bf16-rounded inputs.  Tolerance: 1.2e-2 normwise for the bf16 outputs (the layer-level tolerance of the thin
kernels, tests/test_gpu_parity.py), 3e-3 for the fp32 weight gradients."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import gpu_ops as G

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def q(x):
    return x.bfloat16().float()


def wide_conv(x: G.PF8, w, mode=0, transposed=False, bias=None, res=None, act=None, epi=0, slope=0.0, tm=0, nout=None):
    lib = G.lib()
    wcout, wcin, ks, _ = w.shape
    w = w.float().contiguous().cuda()
    if nout is None:
        nout = wcin if transposed else wcout
    out = G.PF8(x.n, nout, x.h, x.w, "bf16")
    nbytes = int(lib.mil_wide_conv_workspace_bytes(mode, int(transposed), wcout, wcin, ks))
    assert nbytes > 0
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    b = None if bias is None else bias.float().contiguous().cuda()
    G.check(lib.mil_wide_conv_pf8(mode, int(transposed), G._p(x.buf), x.n, x.c, x.h, x.w, G._p(w), wcout, wcin, ks, G._p(b),
                                  G._p(res.buf) if res else None, G._p(act.buf) if act else None, G._p(out.buf), epi,
                                  C.c_float(slope), tm, G._p(ws), nbytes, G._s()), "mil_wide_conv_pf8")
    return out


def wide_wgrad(x: G.PF8, dz: G.PF8, ks, shape):
    lib = G.lib()
    dw = torch.zeros(shape, dtype=torch.float32, device="cuda")
    nbytes = int(lib.mil_wide_wgrad_workspace_bytes(x.n, x.c, dz.c, x.h, x.w, ks))
    assert nbytes > 256
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    G.check(lib.mil_wide_wgrad_pf8(G._p(x.buf), x.n, x.c, x.h, x.w, G._p(dz.buf), dz.c, ks, G._p(dw), G._p(ws), nbytes,
                                   G._s()), "mil_wide_wgrad_pf8")
    return dw


def zero_halo_ok(t: G.PF8):
    raw = t.raw().float()
    return abs(float(raw.abs().sum()) - float(t.to_nchw().abs().sum())) <= 1e-5 * float(raw.abs().sum())


# (cin, cout, ks, H, n, tm)
S1_CASES = [(64, 64, 3, 14, 3, 0), (128, 128, 3, 14, 3, 0), (128, 128, 3, 28, 5, 2), (256, 256, 3, 7, 9, 0),
            (512, 512, 3, 4, 11, 0), (128, 256, 1, 7, 6, 0), (64, 128, 1, 9, 4, 1), (256, 512, 1, 4, 7, 0),
            (128, 128, 3, 13, 1, 4), (64, 64, 3, 56, 2, 0)]


@pytest.mark.parametrize("cin,cout,ks,H,n,tm", S1_CASES)
def test_wide_conv_forward_and_dgrad_vs_torch(cin, cout, ks, H, n, tm):
    gen = torch.Generator(device="cuda").manual_seed(7)
    pad = ks // 2
    x = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    w = torch.randn(cout, cin, ks, ks, device="cuda", generator=gen) / (cin * ks * ks) ** 0.5
    res = q(torch.randn(n, cout, H, H, device="cuda", generator=gen))
    X, R = G.PF8.from_nchw(x, "bf16"), G.PF8.from_nchw(res, "bf16")
    # block tail: relu(conv(x) + identity)   (alt_resnet.py:57-65)
    out = wide_conv(X, w, res=R, epi=0, slope=0.0, tm=tm)
    ref = F.relu(F.conv2d(x, q(w), padding=pad) + res)
    assert G.relerr(out.to_nchw(), ref) < 1.2e-2
    assert zero_halo_ok(out)
    # plain (projection shortcut: no activation), with a LeakyReLU variant and a bias for the epilogue's other branches
    b = torch.randn(cout, device="cuda", generator=gen) * 0.1
    out = wide_conv(X, w, bias=b, epi=0, slope=0.1, tm=tm)
    assert G.relerr(out.to_nchw(), F.leaky_relu(F.conv2d(x, q(w), b, padding=pad), 0.1)) < 1.2e-2
    out = wide_conv(X, w, epi=2, tm=tm)
    assert G.relerr(out.to_nchw(), F.conv2d(x, q(w), padding=pad)) < 1.2e-2
    # data gradient: (conv_transpose(dz) + res) * relu'(act)
    dz = q(torch.randn(n, cout, H, H, device="cuda", generator=gen))
    act = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    rs = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    gi = torch.nn.grad.conv2d_input(x.shape, q(w), dz, padding=pad)
    DZ, ACT, RS = (G.PF8.from_nchw(t, "bf16") for t in (dz, act, rs))
    out = wide_conv(DZ, w, transposed=True, res=RS, act=ACT, epi=1, slope=0.0, tm=tm)
    assert G.relerr(out.to_nchw(), (gi + rs) * (act > 0).float()) < 1.2e-2
    assert zero_halo_ok(out)


@pytest.mark.parametrize("cin,cout,H,n", [(64, 128, 14, 3), (128, 256, 28, 2), (256, 512, 7, 5), (64, 128, 13, 2)])
def test_wide_conv_stride2_forward_on_the_phase_split_input(cin, cout, H, n):
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=gen) / (cin * 9) ** 0.5
    Ho = (H - 1) // 2 + 1
    X = G.PF8.from_nchw(x, "bf16")
    XS = G.PF8(n, 4 * cin, Ho, Ho, "bf16")
    G.check(G.lib().mil_split2_pf8(G._p(X.buf), n, cin, H, H, G._p(XS.buf), G._s()), "mil_split2_pf8")
    out = wide_conv(XS, w, mode=1, epi=0, slope=0.0)
    ref = F.relu(F.conv2d(x, q(w), stride=2, padding=1))
    assert G.relerr(out.to_nchw(), ref) < 1.2e-2
    assert zero_halo_ok(out)


def s2d4(x, h0):
    """xs[(c, ry, rx)][Y][X] = x[c][4Y + ry][4X + rx] (zero beyond the tile)."""
    n, c, S, _ = x.shape
    xp = torch.zeros(n, c, 4 * h0, 4 * h0, device=x.device)
    xp[:, :, :S, :S] = x
    return xp.view(n, c, h0, 4, h0, 4).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 16, h0, h0).contiguous()


@pytest.mark.parametrize("C_,side,n", [(64, 64, 3), (64, 224, 2), (32, 48, 2)])
def test_wide_stem_conv_and_wgrad_in_space_to_depth_form(C_, side, n):
    gen = torch.Generator(device="cuda").manual_seed(13)
    x = q(torch.randn(n, 3, side, side, device="cuda", generator=gen))
    w = torch.randn(C_, 3, 7, 7, device="cuda", generator=gen) / 147 ** 0.5
    hc = (side - 1) // 2 + 1
    h0 = (hc - 1) // 2 + 1
    XS = G.PF8.from_nchw(s2d4(x, h0), "bf16")
    out = wide_conv(XS, w, mode=2, epi=0, slope=0.0, nout=4 * C_).to_nchw()      # [n, (co, a, b), h0, h0]
    ref = F.relu(F.conv2d(x, q(w), stride=2, padding=3))                        # [n, C, hc, hc]
    refp = torch.zeros(n, C_, 2 * h0, 2 * h0, device="cuda")
    refp[:, :, :hc, :hc] = ref
    ref4 = refp.view(n, C_, h0, 2, h0, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, 4 * C_, h0, h0)
    if hc % 2 == 0:
        assert G.relerr(out, ref4) < 1.2e-2
    else:   # phases beyond the conv map hold values nobody reads: compare inside the map only
        mask = torch.zeros_like(refp)
        mask[:, :, :hc, :hc] = 1
        m4 = mask.view(n, C_, h0, 2, h0, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, 4 * C_, h0, h0)
        assert G.relerr(out * m4, ref4) < 1.2e-2
    if C_ % 32 != 0:
        return
    # weight gradient from the four-phase gradient map
    dz = q(torch.randn(n, C_, hc, hc, device="cuda", generator=gen))
    dzp = torch.zeros(n, C_, 2 * h0, 2 * h0, device="cuda")
    dzp[:, :, :hc, :hc] = dz
    dz4 = dzp.view(n, C_, h0, 2, h0, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, 4 * C_, h0, h0).contiguous()
    dw = wide_wgrad(XS, G.PF8.from_nchw(dz4, "bf16"), 7, (C_, 3, 7, 7))
    gw = torch.nn.grad.conv2d_weight(x, w.shape, dz, stride=2, padding=3)
    assert G.relerr(dw, gw) < 3e-3


@pytest.mark.parametrize("cin,cout,ks,H,n", [(128, 128, 3, 14, 3), (128, 128, 3, 28, 4), (256, 256, 3, 7, 9),
                                             (512, 512, 3, 4, 11), (128, 256, 1, 7, 6), (64, 128, 1, 9, 4),
                                             (64, 128, 3, 14, 2), (256, 512, 3, 8, 3), (64, 64, 3, 14, 3), (64, 64, 3, 56, 2)])
def test_wide_wgrad_vs_torch(cin, cout, ks, H, n):
    gen = torch.Generator(device="cuda").manual_seed(17)
    x = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    dz = q(torch.randn(n, cout, H, H, device="cuda", generator=gen))
    dw = wide_wgrad(G.PF8.from_nchw(x, "bf16"), G.PF8.from_nchw(dz, "bf16"), ks, (cout, cin, ks, ks))
    gw = torch.nn.grad.conv2d_weight(x, (cout, cin, ks, ks), dz, padding=ks // 2)
    assert G.relerr(dw, gw) < 3e-3
    # accumulation (+=) and determinism
    dw2 = wide_wgrad(G.PF8.from_nchw(x, "bf16"), G.PF8.from_nchw(dz, "bf16"), ks, (cout, cin, ks, ks))
    assert torch.equal(dw, dw2)


# ---------------------------------------------------------------------------------------------------------------
# the whole wide extractor / WideAttention against the oracle and the golden vectors of the unmodified reference classes
# ---------------------------------------------------------------------------------------------------------------
from oracle import wide_oracle  # noqa: E402
from tests.helpers import grad_sample, wide_case_inputs, wide_golden_cases  # noqa: E402

WCASES = wide_golden_cases()


def wide_extractor_forward_backward(net, bag, dH):
    """mil_wide_forward + mil_wide_backward through the C ABI with a CALLER-CHOSEN upstream gradient dH [n, 80] (the
    extractor's kernels without the head, whose bag-wide BatchNorm1d backward is ill-conditioned on small bags)."""
    lib = G.lib()
    params = [p.detach() for p in net._params()]
    pp = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
    n, side = int(bag.shape[0]), int(bag.shape[2])
    desc = net._desc
    nbytes = int(lib.mil_wide_workspace_bytes(C.byref(desc), n, side))
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    H = torch.empty((n, 80), dtype=torch.float32, device="cuda")
    bag = bag.contiguous().cuda()
    G.check(lib.mil_wide_forward(C.byref(desc), pp, G._p(bag), int(bag.dtype == torch.uint8), None, n, side, G._p(ws), nbytes,
                                 G._p(H), G._s()), "mil_wide_forward")
    grads = torch.zeros(int(lib.mil_wide_param_total(C.byref(desc))), dtype=torch.float32, device="cuda")
    dH = dH.float().contiguous().cuda()
    G.check(lib.mil_wide_backward(C.byref(desc), pp, n, side, G._p(ws), nbytes, G._p(dH), G._p(grads), G._s()),
            "mil_wide_backward")
    torch.cuda.synchronize()
    out = {}
    for nm, shape, off in net._param_table:
        if nm.startswith("cnn."):
            numel = 1
            for d in shape:
                numel *= d
            out[nm] = grads[off:off + numel].view(shape).clone()
    return H, out


def build_wide(meta, params):
    net = G.pkg().WideAttention(n_classes=3, class_weights=meta["cw"], layers=tuple(meta["layers"])).cuda()
    net.load_state_dict(params)
    return net


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("layers,n,side", [((2, 2, 2, 2), 6, 64), ((1, 1, 1, 1), 5, 96), ((2, 2, 2, 2), 3, 100),
                                           ((1, 2, 1, 1), 40, 32)])
def test_wide_extractor_vs_bf16_emulating_oracle(layers, n, side):
    """Features and all extractor gradients against autograd through the oracle with the CUDA path's rounding points
    (bf16 stored maps, bf16 weights / tiles, fp32 accumulation) -- the tight gate -- and against the fp32 oracle."""
    torch.manual_seed(3)
    params = wide_oracle.init_params(7, layers)
    meta = dict(cw=None, layers=list(layers))
    net = build_wide(meta, params).eval()
    from oracle import synth
    bag = torch.from_numpy(synth.make_bag(n, side, seed=3))
    dH = torch.randn(n, 80)
    H, grads = wide_extractor_forward_backward(net, bag, dH)

    def oracle(emulate):
        q = {k: v.clone().requires_grad_(k.startswith("cnn.")) for k, v in params.items()}
        Ho = wide_oracle.alt_resnet_forward(q, bag, layers, emulate_bf16=emulate)
        (Ho * dH).sum().backward()
        return Ho.detach(), {k: v.grad for k, v in q.items() if k.startswith("cnn.")}

    He, ge = oracle("act+w")
    Hf, gf = oracle("")
    e_emul, e_fp32 = G.relerr(H, He), G.relerr(He, Hf)
    assert e_emul < 6e-3, (e_emul, e_fp32)
    assert G.relerr(H, Hf) < 2e-2
    worst = 0.0
    for k in grads:
        if gf[k].abs().max() < 1e-6:
            continue
        d_emul = G.relerr(grads[k], ge[k])
        d_round = G.relerr(ge[k], gf[k])
        worst = max(worst, d_emul)
        # no further from the emulation than twice what bf16 rounding itself moves the gradient (+ a floor for the tiny tensors)
        assert d_emul < max(2e-2, 2.0 * d_round), (k, d_emul, d_round)
        # direction: tight against the emulation; against the fp32 reference the stem's gradient has passed through
        # every bf16-stored map of the network (0.98 on a 6-tile bag, with the emulation at the same distance)
        assert cosine(grads[k], ge[k]) > 0.99, (k, cosine(grads[k], ge[k]))
        assert cosine(grads[k], gf[k]) > min(0.99, cosine(ge[k], gf[k]) - 0.01), (k, cosine(grads[k], gf[k]), cosine(ge[k], gf[k]))
    assert worst > 0.0


@pytest.mark.parametrize("meta,rec", WCASES, ids=[c[0]["name"] for c in WCASES])
def test_wide_attention_vs_reference_golden(meta, rec):
    """WideAttention.forward / backward against the golden vectors of the unmodified reference classes.  The golden
    bags are small (CPU cost), where the head's bag-wide BatchNorm1d amplifies bf16 feature noise: 5e-2 on the head's
    outputs (the thin path's documented tolerance for bags under 32 tiles), 2e-2 on the features themselves."""
    params, bag, Y, cw, kw = wide_case_inputs(meta, rec)
    net = build_wide(meta, params)
    if meta["training"]:
        net.train()
        net.subsample_indices = kw["indices"]
        net.drop_mask = kw["drop_mask"]
    else:
        net.eval()
    out = net(bag.cuda(), Y.cuda())
    out["loss"].backward()
    for k in ("Aterm", "wROIs", "Bterm", "Mterm", "Fterm", "Aterm_mu", "Aterm_var", "loss", "l2", "KLD", "y_pred", "error"):
        assert tuple(out[k].shape) == tuple(rec[f"out.{k}"].shape), k
    ref = {k[4:]: torch.from_numpy(np.asarray(v)) for k, v in rec.items() if k.startswith("out.")}
    assert G.relerr(out["Fterm"], ref["Fterm"]) < 2e-2
    for k in ("Aterm", "Mterm", "y_pred", "loss"):
        assert G.relerr(out[k], ref[k]) < 5e-2, (k, G.relerr(out[k], ref[k]))
    assert G.relerr(out["l2"], ref["l2"]) < 1e-5
    # gradients: finite, and the well-conditioned ones (everything that does not pass through the tiny bag's BatchNorm
    # backward twice) point the same way as the reference's
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    for name in ("cnn.module.fc.weight", "cnn.module.layer4.0.conv2.weight", "cnn.module.conv1.weight"):
        key = f"grad.{name}" if f"grad.{name}" in rec else f"gsample.{name}"
        ref = torch.from_numpy(rec[key])
        got = dict(net.named_parameters())[name].grad
        got = got if key.startswith("grad.") else grad_sample(got)
        if ref.abs().max() > 1e-6:
            assert cosine(got.cpu().reshape(ref.shape), ref) > 0.9, (name, cosine(got.cpu().reshape(ref.shape), ref))


def test_wide_attention_with_fused_adam_and_uint8_tiles():
    """Flat-buffer training step (FusedAdam's direct gradient accumulation) equals the per-tensor path; 8-bit tiles give
    the bits of the normalised fp32 bag."""
    mil = G.pkg()
    from oracle import synth
    layers = (1, 1, 1, 1)
    params = wide_oracle.init_params(9, layers)
    bag = torch.from_numpy(synth.make_bag(12, 64, seed=5)).cuda()
    Y = torch.tensor([1]).cuda()
    a = build_wide(dict(cw=None, layers=list(layers)), params).eval()
    oa = a(bag, Y)
    oa["loss"].backward()
    b = build_wide(dict(cw=None, layers=list(layers)), params).eval()
    opt = mil.FusedAdam(b, lr=1e-3)
    opt.zero_grad()
    ob = b(bag, Y)
    ob["loss"].backward()
    assert torch.equal(oa["Aterm"], ob["Aterm"]) and torch.equal(oa["loss"], ob["loss"])
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa.grad, pb.grad), k
    opt.step()
    u8 = ((bag + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
    f32 = ((u8.cpu().float() / 255.0 - 0.5) / 0.5).cuda()      # ToTensor() + Normalize(.5, .5) as the loader does it, on the CPU
    with torch.no_grad():
        assert torch.equal(a(u8, Y)["Fterm"], a(f32, Y)["Fterm"])


@pytest.mark.parametrize("cin,cout,H,n", [(64, 128, 14, 3), (128, 256, 28, 2), (256, 512, 7, 5), (64, 128, 13, 2), (64, 128, 56, 2)])
def test_wide_stride2_gradients_at_the_output_resolution(cin, cout, H, n):
    """Weight gradient from the phase-split input, data gradient per input parity phase + merge (no zero-stuffing)."""
    lib = G.lib()
    gen = torch.Generator(device="cuda").manual_seed(19)
    x = q(torch.randn(n, cin, H, H, device="cuda", generator=gen))
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=gen) / (cin * 9) ** 0.5
    Ho = (H - 1) // 2 + 1
    dz = q(torch.randn(n, cout, Ho, Ho, device="cuda", generator=gen))
    X, DZ = G.PF8.from_nchw(x, "bf16"), G.PF8.from_nchw(dz, "bf16")
    XS = G.PF8(n, 4 * cin, Ho, Ho, "bf16")
    G.check(lib.mil_split2_pf8(G._p(X.buf), n, cin, H, H, G._p(XS.buf), G._s()), "mil_split2_pf8")
    # weight gradient
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    nbytes = int(lib.mil_wide_wgrad_s2_workspace_bytes(n, cin, cout, Ho, Ho))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    G.check(lib.mil_wide_wgrad_s2_pf8(G._p(XS.buf), n, cin, Ho, Ho, G._p(DZ.buf), cout, G._p(dw), G._p(ws), nbytes, G._s()),
            "mil_wide_wgrad_s2_pf8")
    gw = torch.nn.grad.conv2d_weight(x, w.shape, dz, stride=2, padding=1)
    assert G.relerr(dw, gw) < 3e-3
    # data gradient: (conv_transpose(dz) + projection gradient at the even positions) * relu'(x)
    tsub = q(torch.randn(n, cin, Ho, Ho, device="cuda", generator=gen))
    TS = G.PF8.from_nchw(tsub, "bf16")
    DS = G.PF8(n, 4 * cin, Ho, Ho, "bf16")
    plane_bytes = DS.buf.numel() // (4 * cin // 8)
    phase_bytes = plane_bytes * (cin // 8)

    class View:
        def __init__(self, t, ph):
            self.buf = t.buf[ph * phase_bytes:(ph + 1) * phase_bytes]

    for ph in range(4):
        nb = int(lib.mil_wide_conv_workspace_bytes(3 + ph, 0, cout, cin, 3))
        wsk = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        G.check(lib.mil_wide_conv_pf8(3 + ph, 0, G._p(DZ.buf), n, cout, Ho, Ho, G._p(w.contiguous()), cout, cin, 3, None,
                                      G._p(TS.buf) if ph == 0 else None, G._p(View(XS, ph).buf), G._p(View(DS, ph).buf), 1,
                                      C.c_float(0.0), 0, G._p(wsk), nb, G._s()), "mil_wide_conv_pf8")
    # the same four phases in ONE launch (mode 7: output channels = (phase, ci); what the extractor runs)
    DS7 = G.PF8(n, 4 * cin, Ho, Ho, "bf16")
    nb = int(lib.mil_wide_conv_workspace_bytes(7, 0, cout, cin, 3))
    wsk = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    TS4 = G.PF8(n, 4 * cin, Ho, Ho, "bf16")          # residual: phase (0, 0) = the projection's gradient, zero elsewhere
    TS4.buf[:phase_bytes] = TS.buf[:phase_bytes]
    G.check(lib.mil_wide_conv_pf8(7, 0, G._p(DZ.buf), n, cout, Ho, Ho, G._p(w.contiguous()), cout, cin, 3, None, G._p(TS4.buf),
                                  G._p(XS.buf), G._p(DS7.buf), 1, C.c_float(0.0), 0, G._p(wsk), nb, G._s()), "mil_wide_conv_pf8")
    assert torch.equal(DS7.buf, DS.buf)
    OUT = G.PF8(n, cin, H, H, "bf16")
    OUT.buf.fill_(0x7F)      # the merge must overwrite every pixel, pads with zeros
    G.check(lib.mil_merge2_pf8(G._p(DS.buf), n, cin, H, H, G._p(OUT.buf), G._s()), "mil_merge2_pf8")
    gi = torch.nn.grad.conv2d_input(x.shape, q(w), dz, stride=2, padding=1)
    gi[:, :, ::2, ::2] += tsub
    assert G.relerr(OUT.to_nchw(), gi * (x > 0).float()) < 1.2e-2
    # every pixel of the map (incl. the pad row / column) was written: only the guards still hold the fill pattern
    ref = G.PF8.from_nchw(OUT.to_nchw(), "bf16")
    g0 = int(lib.mil_pf8_bytes(n, cin, H, H, 1))
    assert g0 == OUT.buf.numel()
    a = OUT.raw().float().view(cin // 8, -1, 8)
    b = ref.raw().float().view(cin // 8, -1, 8)
    Hp = H + 1
    lead = ((Hp + 1) + 7) // 8 * 8
    body = slice(lead, lead + n * Hp * Hp)
    assert torch.equal(a[:, body], b[:, body])


def test_wide_full_size_bag_properties():
    """The bag bench.py times for the wide extractor (4096 tiles x 224^2, alt_resnet18 shape, bf16, fwd+bwd) is out of the
    CPU oracle's reach; it is tied to the oracle-checked small cases through size-independent properties:
      * the extractor is tile-independent: the features of the first 64 tiles are BIT-identical to those of a 64-tile bag;
      * attention weights are non-negative and L1-normalised over the bag, logits = attention-weighted instance codes
        (gbm/model.py:213,227-229);
      * shuffling the tiles permutes Fterm / Aterm / Bterm and leaves the bag-level outputs unchanged;
      * all gradients are finite and a second run reproduces them bit for bit (deterministic split-K reductions)."""
    from tests.test_gpu_parity import _device_bag
    n, side = 4096, 224
    layers = (2, 2, 2, 2)
    net = build_wide(dict(cw=None, layers=list(layers)), wide_oracle.init_params(11, layers)).eval()
    bag = _device_bag(n, side, seed=7)
    Y = torch.tensor([0]).cuda()
    out = net(bag, Y)
    out["loss"].backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    assert all(torch.isfinite(g).all() for g in grads.values())
    A, B, Fm = out["Aterm"].double(), out["Bterm"].double(), out["Fterm"]
    assert A.shape == (3, n) and B.shape == (n, 1) and Fm.shape == (n, 80)
    assert (A >= 0).all() and torch.allclose(A.sum(1), torch.ones(3, dtype=torch.float64, device="cuda"), atol=1e-5)
    assert G.relerr(out["Mterm"].double(), A @ B) < 1e-5
    assert G.relerr(out["y_pred"].double(), torch.softmax((A @ B).view(1, 3), 1)) < 1e-5
    with torch.no_grad():
        small = net(bag[:64].contiguous(), Y)["Fterm"]
        assert torch.equal(small, Fm[:64])
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(0)).cuda()
        outp = net(bag[perm].contiguous(), Y)
    assert torch.equal(outp["Fterm"], Fm[perm])
    assert G.relerr(outp["Aterm"], out["Aterm"][:, perm]) < 1e-5 and G.relerr(outp["Bterm"], out["Bterm"][perm]) < 1e-5
    for k in ("Mterm", "y_pred", "loss", "KLD", "Aterm_mu"):
        assert G.relerr(outp[k], out[k]) < 1e-5, k
    assert int(outp["y_pred_hat"]) == int(out["y_pred_hat"])
    net.zero_grad(set_to_none=True)
    net(bag, Y)["loss"].backward()
    for k, p in net.named_parameters():
        assert torch.equal(p.grad, grads[k]), k


def test_wide_programmatic_dependent_launches_equal_plain_launches():
    """`no_pdl` (ordinary launches) against the default programmatic dependent launches of wide_conv_kernel /
    wide_wgrad_kernel: bit-identical outputs and gradients (see tests/test_gpu_crosscheck.py)."""
    lib = G.pkg()._lib
    from oracle import synth
    layers = (1, 1, 1, 1)
    net = build_wide(dict(cw=None, layers=list(layers)), wide_oracle.init_params(5, layers)).eval()
    bag = torch.from_numpy(synth.make_bag(24, 96, seed=6)).cuda()
    Y = torch.tensor([2]).cuda()

    def run():
        net.zero_grad(set_to_none=True)
        out = net(bag, Y)
        out["loss"].backward()
        torch.cuda.synchronize()
        return out, {k: p.grad.clone() for k, p in net.named_parameters()}

    old = lib.get_option("no_pdl")
    try:
        lib.set_option("no_pdl", 1)
        out_p, g_p = run()
        lib.set_option("no_pdl", 0)
        for _ in range(3):
            out_d, g_d = run()
            for k in ("Fterm", "Aterm", "Mterm", "loss"):
                assert torch.equal(out_p[k], out_d[k]), k
            for k in g_p:
                assert torch.equal(g_p[k], g_d[k]), k
    finally:
        lib.set_option("no_pdl", old)


def test_wide_workspace_reuse_leaves_no_state():
    """A second bag through the workspaces a first bag used gives the bits of a fresh module (tests/test_gpu_parity.py)."""
    from oracle import synth
    layers = (2, 2, 2, 2)
    params = wide_oracle.init_params(13, layers)
    bag1 = torch.from_numpy(synth.make_bag(9, 96, seed=31)).cuda()
    bag2 = torch.from_numpy(synth.make_bag(9, 96, seed=32)).cuda() * 0.6
    Y = torch.tensor([1]).cuda()
    used = build_wide(dict(cw=None, layers=list(layers)), params).eval()
    fresh = build_wide(dict(cw=None, layers=list(layers)), params).eval()
    for _ in range(2):
        used(bag1, Y)["loss"].backward()
        with torch.no_grad():
            used(bag1, Y)
    used.zero_grad(set_to_none=True)
    a = used(bag2, Y)
    a["loss"].backward()
    b = fresh(bag2, Y)
    b["loss"].backward()
    for k in ("Fterm", "Aterm", "Mterm", "loss"):
        assert torch.equal(a[k], b[k]), k
    for (k, pa), (_, pb) in zip(used.named_parameters(), fresh.named_parameters()):
        assert torch.equal(pa.grad, pb.grad), k
