"""In-situ kernel times of the bench step (torch.profiler / CUPTI, kernels back to back at the step's real clocks and
cache state -- unlike an ncu launch list, which serialises and cools the GPU between launches).
usage: [MIL_MODEL=wide18] python tools/step_profile.py [tiles] [side] [steps]"""
import collections
import importlib
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = os.environ.get("MIL_MODEL", "resnet26")      # resnet26 | wide18 | wide34 (bench.py --model)
net = bench.make_net(mil, model).to(dev).eval()
bag = bench.make_device_bag(mil, n, side, dev, seed=1)
Y = torch.tensor([1], device=dev)


def step():
    net.zero_grad(set_to_none=True)
    out = net(bag, Y)
    out["loss"].backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    step()
ev1.record()
torch.cuda.synchronize()
print(f"# {n} tiles x {side}^2: {ev0.elapsed_time(ev1) / steps:.3f} ms per step without the profiler")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot, cnt = collections.defaultdict(float), collections.Counter()
for e in evs:
    key = re.sub(r"\(.*", "", e.name).replace("void ", "").strip()
    key = re.sub(r"^(at::native::|at::cuda::)", "", key)[:60]
    tot[key] += e.time_range.elapsed_us()
    cnt[key] += 1
busy = sum(tot.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"# profiled: {span / steps / 1e3:.3f} ms per step wall on the device, {busy / steps / 1e3:.3f} ms of kernels "
      f"({100 * busy / span:.1f} % busy), {len(evs) // steps} kernels per step")
print(f"{'us/step':>10s} {'share':>7s} {'n/step':>6s}  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v / steps:10.1f} {100 * v / busy:6.1f}% {cnt[k] / steps:6.1f}  {k}")
# idle gaps between consecutive kernels (where the device waited for the host or for a dependency)
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 5:
        gaps.append((g, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
print(f"# idle gaps > 5 us: {len(gaps) / steps:.1f} per step, {sum(g for g, _, _ in gaps) / steps:.1f} us per step")
for g, a, b in gaps[:25]:
    print(f"{g:9.1f} us  after {a}  before {b}")
# launches of the weight-gradient kernels in time order (first step): which layer costs what
wg = [e for e in evs if "wgrad" in e.name]
first = wg[: len(wg) // steps]
print("# wgrad launches in order (us):", " ".join(f"{re.sub(r'_kernel.*', '', e.name.replace('void ', ''))[:9]}:{e.time_range.elapsed_us():.0f}" for e in first))
cv = [e for e in evs if "conv_tc_kernel" in e.name]
firstc = cv[: len(cv) // steps]
print("# conv_tc launches in order (us):", " ".join(f"{re.sub(r'.*conv_tc_kernel<', '<', e.name)[:10].replace(' ', '')}:{e.time_range.elapsed_us():.0f}" for e in firstc))
wc = [e for e in evs if "wide_conv_kernel" in e.name]
if wc:
    print("# wide_conv launches in order (us):", " ".join(f"{e.time_range.elapsed_us():.0f}" for e in wc[: len(wc) // steps]))
