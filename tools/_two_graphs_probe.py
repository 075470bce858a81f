import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
mil = importlib.import_module(bench.PKG)
dev = torch.device("cuda", 0)
Y = torch.tensor([1], device=dev)
def attempt(label, fn):
    try:
        fn(); torch.cuda.synchronize(); print(label, "ok", flush=True)
    except Exception as e:
        print(label, "FAILED:", str(e).splitlines()[0], flush=True)
        raise SystemExit(1)
for variant in ("two_steps_same_net", "inject_then_graph", "eager_nostep_then_graph"):
    torch.manual_seed(0)
    net = mil.Attention(n_classes=3).to(dev).train()
    opt = mil.FusedAdam(net, lr=2e-4)
    bag = bench.make_device_bag(mil, 200, 64, dev, seed=1)
    def eager(step=True):
        opt.zero_grad(); out = net(bag, Y); out["loss"].backward()
        if step: opt.step()
    if variant == "two_steps_same_net":
        attempt(variant + " eager", eager)
        s1 = mil.GraphedStep(net, 200, 64, optimizer=opt); attempt(variant + " g1", lambda: s1(bag, Y))
        del s1
        s2 = mil.GraphedStep(net, 200, 64, optimizer=opt); attempt(variant + " g2", lambda: s2(bag, Y))
    elif variant == "inject_then_graph":
        net.subsample_indices = torch.randperm(200)[:40]
        net.drop_mask = (torch.rand((40, 80)) > 0.25).float().to(dev)
        attempt(variant + " eager", eager)
        net.subsample_indices = None; net.drop_mask = None
        s1 = mil.GraphedStep(net, 200, 64, optimizer=opt); attempt(variant + " g1", lambda: s1(bag, Y))
    else:
        attempt(variant + " eager", lambda: eager(False))
        s1 = mil.GraphedStep(net, 200, 64, optimizer=opt); attempt(variant + " g1", lambda: s1(bag, Y))
print("all ok")
