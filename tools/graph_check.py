"""Small-bag step: eager launches vs the CUDA-graph replay (graph.GraphedStep) vs the summed kernel time.
The reference's live pipeline caps a bag at 2 500 tiles and pushes 20 % of them through the CNN (RoiBuilder.py:230,
gbm/model.py:193).   usage: python tools/graph_check.py [bag_tiles] [side] [steps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n_bag = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda", 0)
torch.manual_seed(0)
bag = bench.make_device_bag(mil, n_bag, side, dev, seed=1)
Y = torch.tensor([1], device=dev)


def timed(fn):
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


for mode in ("train", "eval"):
    net = mil.Attention(n_classes=3).to(dev)
    net.train(mode == "train")
    opt = mil.FusedAdam(net, lr=2e-4)

    def eager():
        opt.zero_grad()
        out = net(bag, Y)
        out["loss"].backward()
        opt.step()

    ms_eager = timed(eager)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            eager()
        torch.cuda.synchronize()
    kern = sum(e.time_range.elapsed_us() for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA) / 3 / 1e3
    step = mil.GraphedStep(net, n_bag, side, optimizer=opt)
    step.bag.copy_(bag)          # a loader writes its H2D copy straight into the static bag: no extra device copy
    ms_graph = timed(lambda: step(step.bag, Y))
    cnn = int(n_bag * 0.2) if mode == "train" else n_bag
    print(f"{mode}: bag {n_bag} x {side}^2 ({cnn} tiles through the CNN), fwd+bwd+Adam: eager {ms_eager:.3f} ms/step, "
          f"graph replay {ms_graph:.3f} ms/step, summed kernel time {kern:.3f} ms  "
          f"(graph / kernels = {ms_graph / kern:.2f}, eager / kernels = {ms_eager / kern:.2f})")
