"""Diagnostic sweep for the GPU box: every kernel against an independent checker, printing error numbers and
continuing after failures (pytest -x stops at the first one; this shows them all in one gpurun call).

    python tools/gpu_check.py [--quick]      # writes gpurun_out/check.log as well as stdout

Checkers: torch's own fp32 GPU convolutions (TF32 off) for the layer-level ops, the CPU oracle
(oracle/mil_oracle.py, pinned to the reference's golden vectors) for the network-level runs."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import mil_oracle, synth  # noqa: E402
from tests import gpu_ops as G  # noqa: E402
from tests.helpers import golden_cases, golden_weights  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
LOG = []
FAIL = []


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.append(s)


def report(name, err, tol):
    ok = err < tol
    say(f"  [{'ok' if ok else 'FAIL'}] {name:58s} err={err:.3e} tol={tol:.0e}")
    if not ok:
        FAIL.append(name)


def section(fn):
    def run(*a, **k):
        say(f"== {fn.__name__} {a if a else ''}")
        try:
            fn(*a, **k)
        except Exception:
            say("  EXCEPTION\n" + traceback.format_exc())
            FAIL.append(fn.__name__ + str(a))
        torch.cuda.synchronize()
    return run


def lrelu(x):
    return F.leaky_relu(x, 0.1)


def lgrad(a):
    return torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.1))


def q(x, dtype):
    return x.bfloat16().float() if dtype == "bf16" else x


@section
def layout(dtype):
    x = torch.randn(3, 20, 9, 7, device="cuda")
    t = G.PF8.from_nchw(x, dtype)
    y = t.to_nchw()
    report(f"{dtype} to_pf8/from_pf8 round trip", G.relerr(y, q(x, dtype)), 1e-7)
    raw = t.raw().float()
    report(f"{dtype} sum|pf8| == sum|x| (guards/pads/pad channels are zero)",
           abs(float(raw.abs().sum()) - float(q(x, dtype).abs().sum())) / float(x.abs().sum()), 1e-5)


CONV_CASES = [  # cin, cout, ks, stride, H, n
    (20, 20, 3, 1, 14, 3), (20, 40, 3, 2, 14, 2), (20, 40, 1, 2, 14, 2), (40, 40, 3, 1, 7, 5),
    (40, 60, 3, 2, 13, 3), (40, 60, 1, 2, 13, 3), (60, 60, 3, 1, 9, 2), (60, 80, 3, 2, 10, 2),
    (80, 80, 3, 1, 5, 37), (20, 20, 3, 1, 56, 2),
]


@section
def conv_ops(dtype, impl=0):
    tol = 1e-5 if dtype == "fp32" else 1.2e-2
    gen = torch.Generator(device="cuda").manual_seed(5)
    for (cin, cout, ks, stride, H, n) in CONV_CASES + ([(40, 40, 3, 1, 28, 9), (60, 60, 3, 1, 14, 40), (80, 80, 3, 1, 7, 64),
                                                       (20, 20, 3, 1, 75, 2)] if impl == 2 else []):
        pad = ks // 2
        x = q(torch.randn(n, cin, H, H, device="cuda", generator=gen), dtype)
        w = torch.randn(cout, cin, ks, ks, device="cuda", generator=gen) * (1.0 / (cin * ks * ks) ** 0.5)
        b = torch.randn(cout, device="cuda", generator=gen) * 0.1
        Ho = (H + 2 * pad - ks) // stride + 1
        res = q(torch.randn(n, cout, Ho, Ho, device="cuda", generator=gen), dtype)
        tag = f"{dtype} impl{impl} conv {cin}->{cout} k{ks} s{stride} {H}x{H} n{n}"
        wq = q(w, dtype) if (dtype == "bf16" and impl == 2) else w
        # forward, full epilogue
        X = G.PF8.from_nchw(x, dtype)
        R = G.PF8.from_nchw(res, dtype)
        out = G.conv(X, w, bias=b, res=R, stride=stride, epi=0, impl=impl)
        ref = lrelu(F.conv2d(x, wq, b, stride=stride, padding=pad) + res)
        report(tag + " fwd(+bias+res,lrelu)", G.relerr(out.to_nchw(), ref), tol)
        raw = out.raw().float()
        report(tag + " fwd pads/guards zero",
               abs(float(raw.abs().sum()) - float(out.to_nchw().abs().sum())) / max(1e-9, float(raw.abs().sum())), 1e-5)
        # plain (no bias, no res)
        out = G.conv(X, w, stride=stride, epi=2, impl=impl)
        report(tag + " plain", G.relerr(out.to_nchw(), F.conv2d(x, wq, None, stride=stride, padding=pad)), tol)
        # data gradient with DGRAD epilogue
        dz = q(torch.randn(n, cout, Ho, Ho, device="cuda", generator=gen), dtype)
        act = q(torch.randn(n, cin, H, H, device="cuda", generator=gen), dtype)
        rs = q(torch.randn(n, cin, H, H, device="cuda", generator=gen), dtype)
        DZ, ACT, RS = (G.PF8.from_nchw(t, dtype) for t in (dz, act, rs))
        gstride = stride
        if impl == 2 and stride == 2:
            # tensor-core route for the stride-2 gradients: zero-stuff dz to the input resolution, then stride 1
            DZ = G.upsample2(DZ, H, H)
            gstride = 1
        out = G.conv(DZ, w, res=RS, act=ACT, stride=gstride, epi=1, transposed=True, out_hw=(H, H), impl=impl)
        gi = torch.nn.grad.conv2d_input(x.shape, wq, dz, stride=stride, padding=pad)
        report(tag + " dgrad((acc+res)*lrelu')", G.relerr(out.to_nchw(), (gi + rs) * lgrad(act)), tol)
        # weight gradient
        dw, db = G.wgrad(X, DZ, ks, gstride, impl=impl)
        gw = torch.nn.grad.conv2d_weight(x, w.shape, dz, stride=stride, padding=pad)
        report(tag + " wgrad", G.relerr(dw, gw), 2e-5 if dtype == "fp32" else 2e-3)
        report(tag + " bgrad", G.relerr(db, dz.sum(dim=(0, 2, 3))), 2e-5 if dtype == "fp32" else 2e-3)


def build_net(precision, wm=None, cw=None):
    mil = G.pkg()
    net = mil.Attention(n_classes=3, class_weights=cw).cuda().eval()
    sd = golden_weights()
    if wm is not None:
        sd["weight_mask"] = torch.tensor(wm)
    net.load_state_dict(sd)
    net.precision = precision
    return net


@section
def activations(precision, n=3, side=64):
    """Every saved activation of the extractor against the oracle's taps."""
    tol = 2e-5 if precision == "fp32" else 3e-2
    net = build_net(precision)
    bag = torch.from_numpy(synth.make_bag(n, side, seed=3))
    taps = {}
    p = golden_weights()
    Href = mil_oracle.resnet26_forward(p, bag, taps=taps)
    with torch.no_grad():
        H = net.features(bag.cuda())
    report(f"{precision} stem (conv7x7+lrelu+maxpool)", G.relerr(G.read_activation(net, n, side, -1), taps["stem"]), tol)
    for l in range(4):
        for b in range(3):
            lb = l * 3 + b
            report(f"{precision} layer{l+1}.{b} h", G.relerr(G.read_activation(net, n, side, 2 * lb),
                                                          taps[f"layer{l+1}.{b}.y1"]), tol)
            report(f"{precision} layer{l+1}.{b} y", G.relerr(G.read_activation(net, n, side, 2 * lb + 1),
                                                          taps[f"layer{l+1}.{b}"]), tol)
    report(f"{precision} Fterm", G.relerr(H, Href), tol)


def l2rel(a, b):
    a = a.detach().double().cpu().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a = a.detach().double().cpu().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))


@section
def golden(precision):
    tol_o = 1e-4 if precision == "fp32" else 1e-2
    for meta, rec in golden_cases():
        net = build_net(precision, wm=meta["wm"], cw=None if meta["cw"] is None else torch.tensor(meta["cw"]))
        bag_cpu = torch.from_numpy(synth.make_bag(meta["n"], meta["side"], seed=meta.get("seed", 1)))
        bag = bag_cpu.cuda()
        idx = None
        if meta["training"]:
            net.train()
            idx = torch.from_numpy(rec["extra.indices"])
            net.subsample_indices = idx
            net.drop_mask = torch.from_numpy(synth.make_drop_mask(len(idx), seed=2))
        out = net(bag, torch.tensor([meta["Y"]]).cuda())
        out["loss"].backward()
        torch.cuda.synchronize()
        tag = f"{precision} {meta['name']}"
        tol_g = max(1e-3, 5 * meta.get("gnoise", 0.0)) if precision == "fp32" else 1.0
        for k in ("Fterm", "Aterm", "wROIs", "Bterm", "Mterm", "y_pred", "loss", "Aterm_mu", "Aterm_var", "KLD", "l2"):
            assert tuple(out[k].shape) == tuple(rec[f"out.{k}"].shape), (k, out[k].shape, rec[f"out.{k}"].shape)
            r = torch.from_numpy(rec[f"out.{k}"])
            report(f"{tag} {k} (l2rel {l2rel(out[k], r):.1e})", G.relerr(out[k], r), tol_o)
        report(f"{tag} y_pred_hat", float(int(out["y_pred_hat"]) != int(rec["out.y_pred_hat"])), 0.5)
        report(f"{tag} error", float(float(out["error"]) != float(rec["out.error"])), 0.5)
        if precision == "bf16":
            # kernel check proper: against the oracle that rounds stored activations to bf16 at the same points
            p = golden_weights()
            p["weight_mask"] = torch.tensor(meta["wm"])
            x = bag_cpu if idx is None else bag_cpu[idx]
            He = mil_oracle.resnet26_forward(p, x, emulate_bf16="act+w")
            report(f"{tag} Fterm vs bf16-emulating oracle (l2rel {l2rel(out['Fterm'], He):.1e})",
                   G.relerr(out["Fterm"], He), 4e-3)
        worst, worst_k, wcos, wcos_k, wnr, wnr_k = 0.0, "", 1.0, "", 0.0, ""
        for k, prm in net.named_parameters():
            dig = rec[f"gdigest.{k}"]
            if dig[2] > 1e-5:
                nr = abs(float(prm.grad.double().norm()) - dig[2]) / dig[2]
                if nr > wnr:
                    wnr, wnr_k = nr, k
            if f"grad.{k}" in rec and np.abs(rec[f"grad.{k}"]).max() > 1e-5:
                r = torch.from_numpy(rec[f"grad.{k}"])
                e = G.relerr(prm.grad, r)
                c = cosine(prm.grad, r)
                if e > worst:
                    worst, worst_k = e, k
                if c < wcos:
                    wcos, wcos_k = c, k
        report(f"{tag} worst grad max-norm err ({worst_k}) gnoise={meta.get('gnoise', 0):.1e}", worst, tol_g)
        report(f"{tag} worst grad 1-cosine ({wcos_k})", 1 - wcos, 1e-5 if precision == "fp32" else 2e-2)
        report(f"{tag} worst grad norm ratio err ({wnr_k})", wnr, tol_g if precision == "fp32" else 1e-1)


@section
def timing(precision, n=256, side=224, iters=3):
    net = build_net(precision)
    bag = torch.from_numpy(synth.make_bag(16, side, seed=1)).cuda().repeat(n // 16, 1, 1, 1)
    Y = torch.tensor([1]).cuda()
    for i in range(iters + 1):
        if i == 1:
            torch.cuda.synchronize()
            t0 = time.time()
        net.zero_grad(set_to_none=True)
        out = net(bag, Y)
        out["loss"].backward()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / iters
    say(f"  {precision}: {n} tiles x {side}^2 fwd+bwd {dt*1e3:.1f} ms -> {n/dt:.0f} tiles/s")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        net.features(bag)
        ev0.record()
        for _ in range(iters):
            net.features(bag)
        ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    say(f"  {precision}: {n} tiles x {side}^2 extractor fwd {ms:.1f} ms -> {n/ms*1e3:.0f} tiles/s")


def main():
    quick = "--quick" in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    want = lambda name: only is None or name in only
    say(torch.cuda.get_device_name(0), torch.__version__)
    for dt in ("fp32", "bf16"):
        if want("layout"):
            layout(dt)
    for dt in ("fp32", "bf16"):
        if want("conv"):
            conv_ops(dt, 1)
    if want("tc"):
        conv_ops("bf16", 2)
    for dt in ("fp32", "bf16"):
        if want("act"):
            activations(dt)
    for dt in ("fp32", "bf16"):
        if want("golden"):
            golden(dt)
    if not quick and want("timing"):
        for dt in ("fp32", "bf16"):
            timing(dt)
    say(f"\n{len(FAIL)} failures" + ("" if not FAIL else ":\n  " + "\n  ".join(FAIL)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "check.log"), "w") as f:
        f.write("\n".join(LOG) + "\n")
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
