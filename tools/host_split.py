import importlib, os, sys, time
sys.path.insert(0, "/root/repo")
import torch
mil = importlib.import_module("deep-convolutional-neural-network-resnet-26-and-attention-network_b200")
n = 256
net = mil.Attention(n_classes=3).cuda().eval()
bag = torch.rand((n, 3, 224, 224), device="cuda") * 2 - 1
Y = torch.tensor([1]).cuda()
def T(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        torch.cuda.synchronize(); t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    torch.cuda.synchronize()
    return 1e3*sorted(ts)[len(ts)//2]
print("features (C extractor forward only):", T(lambda: net.features(bag)))
with torch.no_grad():
    print("forward no_grad:", T(lambda: net(bag, Y)))
print("forward (grad):", T(lambda: net(bag, Y)))
def fb():
    out = net(bag, Y); out["loss"].backward()
print("forward+backward (grads accumulate):", T(fb))
def fbz():
    net.zero_grad(set_to_none=True); out = net(bag, Y); out["loss"].backward()
print("zero_grad+forward+backward:", T(fbz))
lib = mil._lib.load()
l0 = lib.mil_kernel_launch_count(); fb(); print("launches per fwd+bwd:", lib.mil_kernel_launch_count()-l0)
opt = mil.FusedAdam(net, lr=2e-4)
def fb_flat():
    out = net(bag, Y); out["loss"].backward()
print("flattened: forward+backward:", T(fb_flat))
def full():
    opt.zero_grad(); out = net(bag, Y); out["loss"].backward(); opt.step()
print("flattened: zero_grad+forward+backward+FusedAdam.step:", T(full))
net2 = mil.Attention(n_classes=3).cuda().eval()
opt2 = torch.optim.Adam(net2.parameters(), lr=2e-4)
def full2():
    opt2.zero_grad(); out = net2(bag, Y); out["loss"].backward(); opt2.step()
print("plain: zero_grad+forward+backward+torch Adam.step:", T(full2))
