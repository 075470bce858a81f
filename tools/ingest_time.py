"""Times the GPU tile ingest (TileIngest) on a slide-sized tile cache and, on a sample, the reference's CPU pipeline
(torchvision / PIL per tile, RoiBuilder.py:193-203).   usage: python tools/ingest_time.py [tiles] [roi] [side]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mil = importlib.import_module(bench.PKG)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
roi = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
side = int(sys.argv[3]) if len(sys.argv) > 3 else 224
rois = torch.randint(0, 256, (T, roi, roi, 3), dtype=torch.uint8, device="cuda")
ing = mil.TileIngest(roi, side)
crops, flips = mil.ingest.draw_augmentations(T)
for train in (True, False):
    for _ in range(2):
        out = ing(rois, train=train, crops=crops, flips=flips)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ing(rois, train=train, crops=crops, flips=flips)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = T * roi * roi * 3 / 1e9
    print(f"GPU ingest {'train' if train else 'flat '}: {T} tiles {roi}^2 -> {side}^2: {ms:.2f} ms "
          f"({T / ms * 1e3:.0f} tiles/s, {gb / ms * 1e3:.0f} GB/s of cached tiles read)")
try:
    import torchvision.transforms as TV
    tf = TV.Compose([TV.ToPILImage(), TV.Pad(100), TV.RandomCrop(roi), TV.Resize(side), TV.RandomHorizontalFlip(0.5),
                     TV.RandomVerticalFlip(0.5), TV.ToTensor(), TV.Normalize((0.5,) * 3, (0.5,) * 3)])
    host = rois[:16].cpu().numpy()
    t0 = time.perf_counter()
    for r in host:
        tf(r)
    dt = (time.perf_counter() - t0) / len(host)
    print(f"reference CPU pipeline (torchvision / PIL, one thread): {dt * 1e3:.2f} ms per tile = {1 / dt:.0f} tiles/s")
except ImportError:
    print("torchvision not available: CPU pipeline not timed")
