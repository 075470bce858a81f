"""Multi-GPU parity check (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tools/multi_gpu_check.py

A bag is sharded over the ranks (uneven shards on purpose); every rank must obtain the SAME logits / loss /
gradients as a single GPU processing the whole bag (the bag-wide sums are exact), and its shard of Aterm / Fterm.
Then the multi-slide mode (SlideGroup, BASELINE configs[4]): every rank its own slide; its outputs must be the
single-GPU outputs of that slide and the reduced gradients the sum of the slides' single-GPU gradients.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

PKG = "deep-convolutional-neural-network-resnet-26-and-attention-network_b200"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    mil = importlib.import_module(PKG)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    def build(kind, precision):
        if kind == "wide":      # alt_resnet.py's network as extractor, one block per layer
            return mil.WideAttention(n_classes=3, layers=(1, 1, 1, 1)).to(dev).eval()
        m = mil.Attention(n_classes=3).to(dev).eval()
        m.precision = precision
        return m

    # same kernels, same inputs: only summation order differs (the wide weight gradient accumulates a whole shard in TMEM:
    # its split over the CTAs changes with the shard size -> 1e-4 on the gradients there)
    for kind, precision, tol in (("resnet26", "fp32", 2e-5), ("resnet26", "bf16", 2e-5), ("wide", "bf16", 3e-4)):
        n, side = 96, 64
        bag = torch.from_numpy(mil.synth.make_bag(n, side, seed=5)).to(dev)
        Y = torch.tensor([2], device=dev)
        torch.manual_seed(0)
        ref = build(kind, precision)
        with torch.no_grad():
            ref.weight_mask.copy_(torch.tensor([-1.0, 0.25, -0.5]))
        out1 = ref(bag, Y)
        out1["loss"].backward()
        g1 = torch.cat([p.grad.flatten() for p in ref.parameters()])
        # sharded run: uneven contiguous shards
        cuts = [0] + [int(n * (r + 1) / world + (3 if r == 0 and world > 1 else 0)) for r in range(world - 1)] + [n]
        lo, hi = cuts[rank], cuts[rank + 1]
        torch.manual_seed(0)
        net = build(kind, precision)
        with torch.no_grad():
            net.weight_mask.copy_(torch.tensor([-1.0, 0.25, -0.5]))
        net.bag_group = mil.BagGroup(dist.group.WORLD, seed=3)
        out = net(bag[lo:hi].contiguous(), Y)
        out["loss"].backward()
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        errs = {
            "loss": rel(out["loss"].detach(), out1["loss"].detach()), "Mterm": rel(out["Mterm"], out1["Mterm"]),
            "y_pred": rel(out["y_pred"], out1["y_pred"]), "Aterm(shard)": rel(out["Aterm"], out1["Aterm"][:, lo:hi]),
            "Fterm(shard)": rel(out["Fterm"], out1["Fterm"][lo:hi]), "KLD": rel(out["KLD"], out1["KLD"]),
            "Aterm_var": rel(out["Aterm_var"], out1["Aterm_var"]), "grads": rel(g, g1),
        }
        bad = {k: v for k, v in errs.items() if not v < tol}
        print(f"rank {rank} {kind} {precision} shard [{lo},{hi}) " + " ".join(f"{k}={v:.1e}" for k, v in errs.items())
              + ("  FAIL " + str(bad) if bad else "  ok"), flush=True)
        ok = ok and not bad
        # train mode: shared-seed subsample + dropout; ranks must agree on the bag-level results
        net.train()
        net.zero_grad(set_to_none=True)
        outt = net(bag[lo:hi].contiguous(), Y)
        outt["loss"].backward()
        v = torch.stack([outt["loss"].detach(), outt["Mterm"].flatten()[0], outt["KLD"]]).double()
        vs = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(vs, v)
        agree = all(torch.equal(vs[0], t) for t in vs)
        cnt = torch.tensor([outt["Aterm"].shape[1]], device=dev)
        dist.all_reduce(cnt)
        print(f"rank {rank} {kind} {precision} train: local tiles {outt['Aterm'].shape[1]} total {int(cnt)} (expect {int(n * 0.2)}) "
              f"ranks agree: {agree} finite: {bool(torch.isfinite(v).all())}", flush=True)
        ok = ok and agree and int(cnt) == int(n * 0.2) and bool(torch.isfinite(v).all())
    # ---- multi-slide data parallelism (BASELINE configs[4]): every rank its own bag; only the gradients cross ranks ----
    for kind, precision, tol in (("resnet26", "fp32", 2e-5), ("resnet26", "bf16", 2e-5), ("wide", "bf16", 3e-4)):
        side = 64
        sizes = [40 + 8 * r for r in range(world)]                      # slides differ in size
        bags = [torch.from_numpy(mil.synth.make_bag(sizes[r], side, seed=20 + r)).to(dev) for r in range(world)]
        labels = [torch.tensor([r % 3], device=dev) for r in range(world)]
        torch.manual_seed(0)
        ref = build(kind, precision)
        outs = []
        for r in range(world):                                          # the reference's loop: one slide after the other,
            o = ref(bags[r], labels[r])                                 # gradients accumulate (gbm/classify_combined.py:446-454)
            o["loss"].backward()
            outs.append(o)
        g1 = torch.cat([p.grad.flatten() for p in ref.parameters()])
        torch.manual_seed(0)
        net = build(kind, precision)
        net.bag_group = mil.SlideGroup(dist.group.WORLD)
        out = net(bags[rank], labels[rank])
        out["loss"].backward()
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        errs = {"loss(own slide)": rel(out["loss"].detach(), outs[rank]["loss"].detach()),
                "Aterm(own slide)": rel(out["Aterm"], outs[rank]["Aterm"]), "grads(sum over slides)": rel(g, g1)}
        same = torch.equal(out["Aterm"], outs[rank]["Aterm"]) and torch.equal(out["Fterm"], outs[rank]["Fterm"])
        bad = {k: v for k, v in errs.items() if not v < tol}
        print(f"rank {rank} {kind} {precision} multi-slide: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items())
              + f" own-slide outputs bit-identical: {same}" + ("  FAIL " + str(bad) if bad else "  ok"), flush=True)
        ok = ok and not bad and same
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
