"""Programmatic dependent launch on / off (mil_set_option("no_pdl", v)): step time of the bench bag (4096 tiles, eval) and of the
reference's live bag shape (2560-tile bag in train mode = 512 tiles through the CNN), eager launches (graph replay: MIL_B200_NO_PDL=1 python
tools/graph_check.py), and a bitwise
comparison of outputs and gradients between the two launch forms.   usage: python tools/pdl_ab.py [model]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
lib = mil._lib.load()
model = sys.argv[1] if len(sys.argv) > 1 else "resnet26"
dev = torch.device("cuda", 0)
Y = torch.tensor([1], device=dev)


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(n_bag, train, steps):
    torch.manual_seed(0)
    net = bench.make_net(mil, model).to(dev)
    net.train(train)
    opt = mil.FusedAdam(net, lr=2e-4)
    bag = bench.make_device_bag(mil, n_bag, 224, dev, seed=1)
    res = {}
    snap = {}
    for no_pdl in (1, 0, 1, 0):
        assert lib.mil_set_option(b"no_pdl", no_pdl) == 0
        if train:
            net.subsample_indices = torch.randperm(n_bag, generator=torch.Generator().manual_seed(5))[: int(n_bag * 0.2)]
            net.drop_mask = (torch.rand((int(n_bag * 0.2), 80), generator=torch.Generator().manual_seed(6)) > 0.25).float().to(dev)

        def eager():
            opt.zero_grad()
            out = net(bag, Y)
            out["loss"].backward()
            return out

        out = eager()
        torch.cuda.synchronize()
        cur = [out["loss"].detach().clone(), out["Aterm"].detach().clone(), opt._gflat.clone() if hasattr(opt, "_gflat") else None]
        if no_pdl in snap:
            pass
        snap[no_pdl] = cur
        ms = timed(eager, steps)
        res.setdefault(no_pdl, []).append(ms)
    same = all((a is None and b is None) or torch.equal(a, b) for a, b in zip(snap[0], snap[1]))
    print(f"{model} bag {n_bag} ({'train: 20 % through the CNN' if train else 'eval: all tiles'}): eager ms/step plain {min(res[1]):.3f} "
          f"PDL {min(res[0]):.3f};  outputs + gradients bit-identical: {same}", flush=True)


run(2560, True, 30)
run(4096, False, 10)
