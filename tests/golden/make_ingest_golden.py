"""Golden vectors for the tile ingest: the reference's `img_finalize` / `img_finalize_flat`
(RoiBuilder.py:193-208) spelled with the very torchvision transforms it composes, run on small synthetic 8-bit tiles
under a fixed seed.  Writes tests/golden/ingest_golden.npz (inputs, the draws, outputs as uint8 = round-trip exact:
ToTensor / Normalize are affine in the 8-bit value).   usage: python tests/golden/make_ingest_golden.py"""
import os

import numpy as np
import torch
import torchvision.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))


def finalize(roi_size, resolution):      # RoiBuilder.py:193-203
    return T.Compose([T.ToPILImage(), T.Pad(100), T.RandomCrop(roi_size), T.Resize(resolution),
                      T.RandomHorizontalFlip(p=0.5), T.RandomVerticalFlip(p=0.5), T.ToTensor(),
                      T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])


def finalize_flat(resolution):           # RoiBuilder.py:204-208
    return T.Compose([T.ToPILImage(), T.Resize(resolution), T.ToTensor(), T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])


def to_u8(x):                            # inverse of ToTensor + Normalize(.5, .5): exact for 8-bit sources
    return torch.round((x * 0.5 + 0.5) * 255.0).to(torch.uint8).numpy()


out = {}
rng = np.random.default_rng(7)
for name, (roi, res, n) in {"a": (120, 48, 5), "b": (97, 33, 4), "c": (64, 80, 3)}.items():
    yy, xx = np.mgrid[0:roi, 0:roi]
    tiles = []
    for t in range(n):
        base = (128 + 100 * np.sin(xx / (3.0 + t)) * np.cos(yy / (5.0 + 2 * t)))[..., None] + rng.integers(-20, 20, (roi, roi, 3))
        tiles.append(np.clip(base, 0, 255).astype(np.uint8))
    rois = np.stack(tiles)
    torch.manual_seed(11)
    train = torch.stack([finalize(roi, res)(r) for r in rois])
    flat = torch.stack([finalize_flat(res)(r) for r in rois])
    assert torch.equal((torch.from_numpy(to_u8(train)).float() / 255 - 0.5) / 0.5, train)
    out[f"{name}_rois"], out[f"{name}_train"], out[f"{name}_flat"] = rois, to_u8(train), to_u8(flat)
    out[f"{name}_meta"] = np.array([roi, res, 11])
np.savez_compressed(os.path.join(HERE, "ingest_golden.npz"), **out)
print("wrote", os.path.join(HERE, "ingest_golden.npz"), {k: v.shape for k, v in out.items()})
