"""Prints the numbers behind tests/test_gpu_crosscheck.py without asserting (used to set its tolerances)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mil_oracle, synth  # noqa: E402
from tests import gpu_ops as G  # noqa: E402
from tests.helpers import golden_cases, golden_weights, perturbed_weights  # noqa: E402
from tests.test_gpu_parity import build_net, cosine, l2rel  # noqa: E402
from tests.test_gpu_crosscheck import run  # noqa: E402

lib = G.pkg()._lib


def worst(ga, gb):
    items = [(l2rel(ga[k], gb[k].cpu() if gb[k].is_cuda else gb[k]), k) for k in ga if float(gb[k].norm()) > 1e-12]
    items.sort(reverse=True)
    return items[:3]


SKIP = os.environ.get("CC_ONLY_EXTRACTOR") == "1"
for n, side in ([] if SKIP else [(40, 224), (33, 96), (6, 129), (5, 256)]):
    net = build_net("bf16")
    net.load_state_dict(perturbed_weights(1))
    bag = torch.from_numpy(synth.make_bag(n, side, seed=7)).cuda()
    Y = torch.tensor([2]).cuda()
    o1, g1 = run(net, bag, Y)
    lib.set_option("disable_tc", 1)
    o2, g2 = run(net, bag, Y)
    lib.set_option("disable_tc", 0)
    print(f"tc-vs-cc n={n} side={side}:", {k: f"{G.relerr(o1[k], o2[k]):.2e}" for k in ("Fterm", "Aterm", "Mterm", "y_pred", "loss")},
          "worst grads", [(f"{v:.2e}", k) for v, k in worst(g1, g2)], flush=True)

for n, side in [(24, 224), (9, 64), (7, 129), (6, 256)]:
    net = build_net("bf16")
    bag = torch.from_numpy(synth.make_bag(n, side, seed=8)).cuda()
    Y = torch.tensor([0]).cuda()
    o1, g1 = run(net, bag, Y)
    lib.set_option("stem_unfused", 1)
    o2, g2 = run(net, bag, Y)
    lib.set_option("stem_unfused", 0)
    print(f"fused-vs-unfused n={n} side={side}: equal outputs", all(torch.equal(o1[k], o2[k]) for k in ("Fterm", "Aterm", "loss")),
          "worst grads", [(f"{v:.2e}", k) for v, k in worst(g1, g2)], flush=True)

for meta, rec in ([] if SKIP else golden_cases()):
    nh = meta["n"] if not meta["training"] else len(rec["extra.indices"])
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
    p["weight_mask"] = torch.tensor(meta["wm"])
    emu, eg = mil_oracle.forward_backward(p, bag_cpu, Y, class_weights=cw, training=meta["training"], indices=idx,
                                          drop_mask=drop, emulate_bf16="act+w")
    ref_g = {k: torch.from_numpy(rec[f"grad.{k}"]) for k in g if f"grad.{k}" in rec}
    cosw = min((cosine(g[k], ref_g[k]), k) for k in ref_g if float(ref_g[k].norm()) > 1e-9)
    print(f"emu {meta['name']} (head tiles {nh}):", {k: f"{G.relerr(out[k], emu[k]):.2e}" for k in ("Fterm", "Aterm", "Mterm", "y_pred", "loss")},
          "worst grads vs emu", [(f"{v:.2e}", k) for v, k in worst(g, eg)], "worst cosine vs fp32 ref", cosw, flush=True)

for n, side in ([] if SKIP else [(64, 224), (256, 64), (40, 96)]):
    p = perturbed_weights(2)
    net = build_net("bf16")
    net.load_state_dict(p)
    bag_cpu = torch.from_numpy(synth.make_bag(n, side, seed=11))
    Y = torch.tensor([1])
    out, g = run(net, bag_cpu.cuda(), Y.cuda())
    emu, eg = mil_oracle.forward_backward(p, bag_cpu, Y, emulate_bf16="act+w")
    ref, rg = mil_oracle.forward_backward(p, bag_cpu, Y)
    print(f"emu-perturbed n={n} side={side}:", {k: f"{G.relerr(out[k], emu[k]):.2e}" for k in ("Fterm", "Aterm", "Mterm", "y_pred", "loss")},
          "worst grads vs emu", [(f"{v:.2e}", k) for v, k in worst(g, eg)], "vs fp32 oracle", [(f"{v:.2e}", k) for v, k in worst(g, rg)],
          "emu vs fp32 oracle", [(f"{v:.2e}", k) for v, k in worst(eg, rg)], flush=True)

def whole_path_sections():
    # zero tiles
    p = perturbed_weights(3, conv_bias=False)
    bag_cpu = torch.from_numpy(synth.make_bag(40, 64, seed=9))
    bag_cpu[::4] = 0.0
    Y = torch.tensor([1])
    ref, rg = mil_oracle.forward_backward(p, bag_cpu, Y)
    emu, eg = mil_oracle.forward_backward(p, bag_cpu, Y, emulate_bf16="act+w")
    net32 = build_net("fp32"); net32.load_state_dict(p)
    o32, g32 = run(net32, bag_cpu.cuda(), Y.cuda())
    net = build_net("bf16"); net.load_state_dict(p)
    o1, g1 = run(net, bag_cpu.cuda(), Y.cuda())
    lib.set_option("disable_tc", 1)
    o2, g2 = run(net, bag_cpu.cuda(), Y.cuda())
    lib.set_option("disable_tc", 0)
    print("zero tiles: fp32 vs oracle", [(f"{v:.2e}", k) for v, k in worst(g32, rg)], "tc vs cc", [(f"{v:.2e}", k) for v, k in worst(g1, g2)],
          "tc vs emu", [(f"{v:.2e}", k) for v, k in worst(g1, eg)], flush=True)

    # headline
    from tests.test_gpu_parity import _device_bag  # noqa: E402
    bag = _device_bag(4096, 224, seed=5)
    Y = torch.tensor([2]).cuda()
    for name, p in (("golden", golden_weights()), ("perturbed", perturbed_weights(4))):
        net16 = build_net("bf16"); net16.load_state_dict(p)
        o16, g16 = run(net16, bag, Y)
        o16 = {k: v.detach().clone() for k, v in o16.items()}
        del net16
        torch.cuda.empty_cache()
        net32 = build_net("fp32"); net32.load_state_dict(p)
        o32, g32 = run(net32, bag, Y)
        cosw = sorted((cosine(g16[k], g32[k].cpu()), k) for k in g16 if float(g32[k].norm()) > 1e-12)[:4]
        print(f"headline {name}:", {k: f"{G.relerr(o16[k], o32[k]):.2e}" for k in ("Fterm", "Aterm", "Mterm", "y_pred", "loss")},
              "worst grads l2", [(f"{v:.2e}", k) for v, k in worst(g16, g32)], "worst cos", cosw, flush=True)
        del net32
        torch.cuda.empty_cache()




if not SKIP:
    whole_path_sections()

# ---- extractor alone, caller-chosen dH ----
def oracle_extractor(p, bag, dH, emu):
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items() if k.startswith("cnn.")}
    H = mil_oracle.resnet26_forward(q, bag, emulate_bf16=emu)
    H.backward(dH)
    return H.detach(), {k: v.grad for k, v in q.items()}

for n, side, zero in [(40, 224, False), (33, 96, False), (6, 129, False), (5, 256, False), (40, 64, True)]:
    p = perturbed_weights(5, conv_bias=not zero)
    bag = torch.from_numpy(synth.make_bag(n, side, seed=12))
    if zero:
        bag[::4] = 0.0
    dH = torch.randn(n, 80, generator=torch.Generator().manual_seed(3))
    H1, g1 = G.extractor_forward_backward(p, bag, dH, "bf16")
    lib.set_option("disable_tc", 1)
    H2, g2 = G.extractor_forward_backward(p, bag, dH, "bf16")
    lib.set_option("disable_tc", 0)
    H3, g3 = G.extractor_forward_backward(p, bag, dH, "fp32")
    He, ge = oracle_extractor(p, bag, dH, "act+w")
    Hr, gr = oracle_extractor(p, bag, dH, "")
    print(f"extractor n={n} side={side} zero={zero}: H tc/cc {G.relerr(H1,H2):.2e} tc/emu {G.relerr(H1,He):.2e} fp32/ref {G.relerr(H3,Hr):.2e} emu/ref {G.relerr(He,Hr):.2e}",
          "\n   tc vs cc", [(f"{v:.2e}", k[11:]) for v, k in worst(g1, g2)],
          "\n   tc vs emu", [(f"{v:.2e}", k[11:]) for v, k in worst(g1, ge)],
          "\n   cc vs emu", [(f"{v:.2e}", k[11:]) for v, k in worst(g2, ge)],
          "\n   fp32 vs ref", [(f"{v:.2e}", k[11:]) for v, k in worst(g3, gr)],
          "\n   emu vs ref", [(f"{v:.2e}", k[11:]) for v, k in worst(ge, gr)], flush=True)
