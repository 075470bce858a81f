"""-m gpu cross-checks of the bf16 production kernels (the path bench.py times), all through the C ABI:

  * tcgen05 path vs the CUDA-core path of the SAME library under `mil_set_option("disable_tc", 1)`: the CUDA-core
    kernels round where the tensor-core ones round (bf16 weights, bf16 input tiles, bf16 conv map before the pool), so
    the two differ by fp32 summation order only -- every one of the 65 gradient tensors must agree to 2e-3 normwise.
    This covers the stride-2 phase-split forms, the fused stem forward / backward and the TMEM lane-half accumulators,
    none of which the layer-level operators reach;
  * fused stem (pool in the conv epilogue, un-pool inside the weight-gradient kernel) vs `stem_unfused`;
  * all-zero tiles (the reference's loader really feeds them: RoiBuilder.py:234-236): every activation is exactly 0 at
    zero bias, LeakyReLU'(0) must be the slope like ATen's;
  * bf16 gradients vs the bf16-EMULATING oracle (autograd through the restatement with the product's rounding points);
  * BASELINE configs[1] itself (4096 tiles x 224^2): bf16 mode vs the library's fp32 check mode, which is pinned to the
    reference's golden vectors at 1e-4.
"""
import pytest
import torch

from oracle import mil_oracle, synth
from tests import gpu_ops as G
from tests.helpers import golden_cases, golden_weights
from tests.test_gpu_parity import _device_bag, build_net, cosine, l2rel

pytestmark = pytest.mark.gpu
CASES = golden_cases()


@pytest.fixture
def option():
    """Set library switches for one test; everything is switched back afterwards."""
    lib = G.pkg()._lib
    touched = {}

    def set_(name, value):
        touched.setdefault(name, lib.get_option(name))
        lib.set_option(name, value)

    yield set_
    for name, old in touched.items():
        lib.set_option(name, old)


def run(net, bag, Y, **kw):
    net.zero_grad(set_to_none=True)
    out = net(bag, Y, **kw)
    out["loss"].backward()
    torch.cuda.synchronize()
    return out, {k: p.grad.detach().clone() for k, p in net.named_parameters()}


def test_options_round_trip():
    lib = G.pkg()._lib
    for name in ("disable_tc", "stem_unfused"):
        old = lib.get_option(name)
        lib.set_option(name, 1)
        assert lib.get_option(name) == 1
        lib.set_option(name, old)
    with pytest.raises(RuntimeError):
        lib.set_option("no_such_switch", 1)


@pytest.mark.parametrize("n,side", [(40, 224), (33, 96), (6, 129), (5, 256)])
def test_tcgen05_path_vs_cuda_core_path(option, n, side):
    """224: BASELINE tile size (all phase-split forms, fused stem).  96: small even maps.  129: odd maps at every
    level (un-fused stem fallback, odd stride-2 inputs).  256: configs[4] tile size."""
    net = build_net("bf16")
    bag = torch.from_numpy(synth.make_bag(n, side, seed=7)).cuda()
    Y = torch.tensor([2]).cuda()
    out_tc, g_tc = run(net, bag, Y)
    option("disable_tc", 1)
    out_cc, g_cc = run(net, bag, Y)
    assert G.relerr(out_tc["Fterm"], out_cc["Fterm"]) < 2e-3
    for k in ("Aterm", "Mterm", "y_pred", "loss"):
        assert G.relerr(out_tc[k], out_cc[k]) < 2e-3, k
    worst = max((l2rel(g_tc[k], g_cc[k].cpu()), k) for k in g_tc if float(g_cc[k].norm()) > 1e-12)
    assert worst[0] < 2e-3, worst


@pytest.mark.parametrize("n,side", [(24, 224), (9, 64)])
def test_fused_stem_vs_unfused_stem(option, n, side):
    net = build_net("bf16")
    bag = torch.from_numpy(synth.make_bag(n, side, seed=8)).cuda()
    Y = torch.tensor([0]).cuda()
    out_f, g_f = run(net, bag, Y)
    option("stem_unfused", 1)
    out_u, g_u = run(net, bag, Y)
    for k in ("Fterm", "Aterm", "Mterm", "loss"):
        assert torch.equal(out_f[k], out_u[k]), k          # same arithmetic, another kernel split
    for k in g_f:
        if k.startswith("cnn.module.conv1."):
            # same MMAs over the same bf16 operands; the split-K partial sums are reduced in the same order
            assert l2rel(g_f[k], g_u[k].cpu()) < 1e-5, (k, l2rel(g_f[k], g_u[k].cpu()))
        else:
            assert torch.equal(g_f[k], g_u[k]), k


def test_all_zero_tiles_take_the_slope_branch(option):
    """Tiles of exact zeros (RoiBuilder.py:234-236 returns torch.zeros(20,3,128,128) for an empty slide region): with
    the reference's zero-initialised biases EVERY activation of such a tile is +0, and ATen's LeakyReLU backward gives
    the slope there (x > 0 ? g : slope * g).  The bias gradients see the difference (weight gradients of a zero tile
    vanish either way).  fp32 mode vs the oracle; bf16 tensor-core path (sign masks) vs the CUDA-core path (reads the
    activations) and vs the bf16-emulating oracle."""
    n, side = 40, 64
    bag_cpu = torch.from_numpy(synth.make_bag(n, side, seed=9))
    bag_cpu[::4] = 0.0                                   # ten all-zero tiles
    Y = torch.tensor([1])
    ref, ref_g = mil_oracle.forward_backward(golden_weights(), bag_cpu, Y)
    net32 = build_net("fp32")
    out, g32 = run(net32, bag_cpu.cuda(), Y.cuda())
    assert float(out["Fterm"][0].abs().max()) == 0.0
    for k in g32:
        if float(ref_g[k].norm()) > 1e-9:
            assert l2rel(g32[k], ref_g[k]) < 1e-3, (k, l2rel(g32[k], ref_g[k]))
    net = build_net("bf16")
    out_tc, g_tc = run(net, bag_cpu.cuda(), Y.cuda())
    assert float(out_tc["Fterm"][0].abs().max()) == 0.0
    option("disable_tc", 1)
    _, g_cc = run(net, bag_cpu.cuda(), Y.cuda())
    for k in g_tc:
        if float(g_cc[k].norm()) > 1e-12:
            assert l2rel(g_tc[k], g_cc[k].cpu()) < 2e-3, (k, l2rel(g_tc[k], g_cc[k].cpu()))
    _, emu_g = mil_oracle.forward_backward(golden_weights(), bag_cpu, Y, emulate_bf16="act+w")
    for k in g_tc:
        if k.endswith(".bias") and float(emu_g[k].norm()) > 1e-9:
            assert l2rel(g_tc[k], emu_g[k]) < 3e-2, (k, l2rel(g_tc[k], emu_g[k]))


BIG = [c for c in CASES if (c[0]["n"] if not c[0]["training"] else len(c[1]["extra.indices"])) >= 32]


@pytest.mark.parametrize("meta,rec", BIG, ids=[c[0]["name"] for c in BIG])
def test_bf16_gradients_vs_bf16_emulating_oracle(meta, rec):
    """The gate for the backward pass in the precision the bench times: autograd through the oracle with the product's
    rounding points (bf16 stored activations and gradient maps, bf16 tensor-core operands, fp32 accumulation).  What is
    left is summation order plus the handful of elements that sit within an fp32 ulp of a bf16 rounding boundary or of a
    LeakyReLU kink: every gradient tensor within 3e-2 normwise."""
    cw = None if meta["cw"] is None else torch.tensor(meta["cw"])
    net = build_net("bf16", wm=meta["wm"], cw=cw)
    bag_cpu = torch.from_numpy(synth.make_bag(meta["n"], meta["side"], seed=meta.get("seed", 1)))
    Y = torch.tensor([meta["Y"]])
    idx = drop = None
    if meta["training"]:
        net.train()
        idx = torch.from_numpy(rec["extra.indices"])
        drop = torch.from_numpy(synth.make_drop_mask(len(idx), seed=2))
        net.subsample_indices, net.drop_mask = idx, drop
    out, g = run(net, bag_cpu.cuda(), Y.cuda())
    p = golden_weights()
    if meta["wm"] is not None:
        p["weight_mask"] = torch.tensor(meta["wm"])
    emu, emu_g = mil_oracle.forward_backward(p, bag_cpu, Y, class_weights=cw, training=meta["training"], indices=idx,
                                             drop_mask=drop, emulate_bf16="act+w")
    assert G.relerr(out["Fterm"], emu["Fterm"]) < 6e-3
    for k in ("Aterm", "Mterm", "y_pred", "loss"):
        assert G.relerr(out[k], emu[k]) < 1e-2, (k, G.relerr(out[k], emu[k]))
    worst = max((l2rel(g[k], emu_g[k]), k) for k in g if float(emu_g[k].norm()) > 1e-9)
    assert worst[0] < 3e-2, worst


def test_headline_config_bf16_vs_fp32_check_mode():
    """BASELINE configs[1] (4096 tiles x 224^2, the bag bench.py times) is out of the CPU oracle's reach; the library's
    own fp32 mode is pinned to the reference's golden vectors at 1e-4 and runs at this size on the GPU, so it serves
    as the oracle here: north_star's bf16 tolerances on the named outputs, identical predicted class, identical
    strongest tiles, per-tensor gradient cosine >= 0.99."""
    n, side = 4096, 224
    bag = _device_bag(n, side, seed=5)
    Y = torch.tensor([2]).cuda()
    net16 = build_net("bf16", wm=[-1.0, -1.0, -1.0])     # peaked attention: a ranking worth comparing
    out16, g16 = run(net16, bag, Y)
    out16 = {k: v.detach().clone() for k, v in out16.items()}
    del net16
    torch.cuda.empty_cache()
    net32 = build_net("fp32", wm=[-1.0, -1.0, -1.0])
    out32, g32 = run(net32, bag, Y)
    for k in ("Aterm", "Mterm", "y_pred", "loss"):
        assert G.relerr(out16[k], out32[k]) < 1e-2, (k, G.relerr(out16[k], out32[k]))
    assert G.relerr(out16["Fterm"], out32["Fterm"]) < 1e-2
    assert int(out16["y_pred_hat"]) == int(out32["y_pred_hat"])
    for m in range(3):
        a32 = out32["Aterm"][m]
        top = torch.topk(a32, 9)
        if float(top.values[7] - top.values[8]) > 2e-2 * float(a32.max()):
            assert set(torch.topk(out16["Aterm"][m], 8).indices.tolist()) == set(top.indices[:8].tolist()), m
        if float(top.values[0] - top.values[1]) > 2e-2 * float(top.values[0]):
            assert int(out16["Aterm"][m].argmax()) == int(top.indices[0])
    worst = min((cosine(g16[k], g32[k].cpu()), k) for k in g16 if float(g32[k].norm()) > 1e-12)
    assert worst[0] >= 0.99, worst
    for k in g16:
        if float(g32[k].norm()) > 1e-12:
            r = float(g16[k].norm() / g32[k].norm())
            assert 0.9 < r < 1.1, (k, r)
