"""-m gpu: the ingest kernel (mil_ingest_tiles_u8 through TileIngest) bit for bit against the CPU restatement of
Pillow's resampling and against the golden vectors of the reference's transform pipeline (RoiBuilder.py:193-210)."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO
from tests import gpu_ops as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz")


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_gpu_ingest_reproduces_reference_golden_vectors(name):
    z = np.load(GOLD)
    rois, (roi, res, seed) = z[f"{name}_rois"], z[f"{name}_meta"]
    ing = G.pkg().TileIngest(int(roi), int(res))
    torch.manual_seed(int(seed))
    got = ing(torch.from_numpy(rois).cuda(), train=True)          # draws crops / flips like torchvision does
    assert np.array_equal(got.cpu().numpy(), z[f"{name}_train"])
    flat = ing(torch.from_numpy(rois).cuda(), train=False)
    assert np.array_equal(flat.cpu().numpy(), z[f"{name}_flat"])


@pytest.mark.parametrize("roi,res,n", [(1200, 224, 3), (300, 64, 5), (97, 33, 6), (64, 96, 4), (128, 128, 2), (50, 7, 3),
                                       (401, 256, 2)])
def test_gpu_ingest_equals_oracle(roi, res, n):
    rng = np.random.default_rng(roi * 7 + res)
    rois = rng.integers(0, 256, (n, roi, roi, 3), dtype=np.uint8)
    crops = np.stack([rng.integers(0, 201, n), rng.integers(0, 201, n)], axis=1).astype(np.int32)
    crops[0] = (0, 0)
    crops[-1] = (200, 200)                                         # extreme offsets: 100 zero rows / columns
    flips = (np.arange(n) % 4).astype(np.uint8)
    ing = G.pkg().TileIngest(roi, res)
    got = ing(torch.from_numpy(rois).cuda(), train=True, crops=crops, flips=flips).cpu().numpy()
    ref = np.stack([IO.finalize_tile(r, res, crop=tuple(c), hflip=bool(f & 1), vflip=bool(f & 2))
                    for r, c, f in zip(rois, crops, flips)])
    assert np.array_equal(got, ref)
    flat = ing(torch.from_numpy(rois).cuda(), train=False).cpu().numpy()
    assert np.array_equal(flat, np.stack([IO.finalize_tile(r, res) for r in rois]))


def test_ingested_bag_feeds_the_extractor():
    """The 8-bit bag goes straight into the stem's 8-bit load: same features as the reference's float pipeline
    (ToTensor + Normalize on the resized tiles) fed as fp32."""
    mil = G.pkg()
    rng = np.random.default_rng(3)
    rois = rng.integers(0, 256, (12, 150, 150, 3), dtype=np.uint8)
    ing = mil.TileIngest(150, 64)
    torch.manual_seed(2)
    bag_u8 = ing(torch.from_numpy(rois).cuda(), train=True)
    net = mil.Attention(n_classes=3).cuda().eval()
    f_u8 = net.features(bag_u8)
    f_32 = net.features((bag_u8.float() / 255.0 - 0.5) / 0.5)
    assert G.relerr(f_u8, f_32) < 1e-2          # bf16 mode: the two normalisations round differently in the last bit
    net.precision = "fp32"
    assert torch.equal(net.features(bag_u8), net.features((bag_u8.float() / 255.0 - 0.5) / 0.5))
