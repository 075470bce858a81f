"""Debug aid: intermediate gradients of the extractor backward vs the fp64 oracle's autograd (one small bag)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import torch.nn.functional as F
from oracle import mil_oracle, synth
from tests import gpu_ops as G
from tests.helpers import golden_weights

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n, side = int(os.environ.get("DBG_N", 8)), int(os.environ.get("DBG_SIDE", 64))
p = {k: v.double().requires_grad_(True) for k, v in golden_weights().items()}
bag = torch.from_numpy(synth.make_bag(n, side, seed=1))
taps = {}
H = mil_oracle.resnet26_forward(p, bag.double(), taps=taps)
for t in taps.values():
    t.retain_grad()
out = mil_oracle.head_forward(p, H, torch.tensor([1]))
out["loss"].backward()
lg = lambda a: torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.1))

mil = G.pkg()
lib = G.lib()
net = mil.Attention(3).cuda().eval()
net.load_state_dict(golden_weights())
net.precision = precision
hc = (side - 1) // 2 + 1
hs = [(hc - 1) // 2 + 1]
for _ in range(3):
    hs.append((hs[-1] - 1) // 2 + 1)
W = (20, 40, 60, 80)
for l in range(3, -1, -1):
    for b in range(2, -1, -1):
        for which in (1, 0):
            if which == 0:
                li = l if b > 0 else max(l - 1, 0)
                name = f"layer{l+1}.{b-1}" if b > 0 else (f"layer{l}.2" if l > 0 else "stem")
            else:
                li, name = l, f"layer{l+1}.{b}.y1"
            buf = torch.zeros((n, W[li], hs[li], hs[li]), device="cuda")
            lib.mil_debug_dump_gradient(l, b, which, C.c_void_p(buf.data_ptr()))
            net.zero_grad(set_to_none=True)
            o = net(bag.cuda(), torch.tensor([1]).cuda())
            o["loss"].backward()
            torch.cuda.synchronize()
            lib.mil_debug_dump_gradient(-1, -1, 0, None)
            ref = taps[name].grad * lg(taps[name].detach())
            d = (buf.double().cpu() - ref)
            e = float(d.abs().max() / ref.abs().max())
            # where is the worst pixel?
            idx = int(d.abs().flatten().argmax())
            shp = ref.shape
            w_ = idx % shp[3]; h_ = (idx // shp[3]) % shp[2]; c_ = (idx // (shp[3] * shp[2])) % shp[1]; n_ = idx // (shp[3] * shp[2] * shp[1])
            per_img = (d.abs().flatten(1).max(dim=1).values / ref.abs().max()).tolist()
            print("   per-image:", " ".join(f"{v:.1e}" for v in per_img))
            print(f"block ({l},{b}) which={which} {name:16s} relerr={e:.3e} worst at n={n_} c={c_} y={h_} x={w_} "
                  f"ours={float(buf[n_, c_, h_, w_]):+.4e} ref={float(ref[n_, c_, h_, w_]):+.4e}", flush=True)
print("param grads vs f64:")
for k, prm in net.named_parameters():
    r = p[k].grad
    if r is not None and r.abs().max() > 1e-7:
        e = float((prm.grad.double().cpu() - r).abs().max() / r.abs().max())
        if e > 2e-4:
            print(f"  {k:45s} {e:.3e}")
