"""Tile ingest (reference RoiBuilder.py:193-210): the CPU restatement of Pillow's 8-bit bilinear
resampling against Pillow itself and against golden vectors of the reference's transform pipeline; the host-side
coefficient tables of the product against the restatement."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO

PKG = "deep-convolutional-neural-network-resnet-26-and-attention-network_b200"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz")


def golden():
    return np.load(GOLD)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_reproduces_reference_pipeline_golden_vectors(name):
    z = golden()
    rois, (roi, res, seed) = z[f"{name}_rois"], z[f"{name}_meta"]
    torch.manual_seed(int(seed))
    crops, flips = IO.draw_params(len(rois), int(roi))          # torchvision's own draws, in its order
    got = np.stack([IO.finalize_tile(r, int(res), crop=tuple(c), hflip=bool(f & 1), vflip=bool(f & 2))
                    for r, c, f in zip(rois, crops, flips)])
    assert np.array_equal(got, z[f"{name}_train"])
    flat = np.stack([IO.finalize_tile(r, int(res)) for r in rois])
    assert np.array_equal(flat, z[f"{name}_flat"])


@pytest.mark.parametrize("roi,res", [(96, 40), (128, 128), (150, 64), (64, 96), (257, 100), (600, 224)])
def test_oracle_resize_equals_pillow(roi, res):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(roi + res).integers(0, 256, (roi, roi, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((res, res), Image.BILINEAR))
    assert np.array_equal(IO.resize_u8(img, res), ref)


@pytest.mark.parametrize("roi,res", [(1200, 224), (1200, 256), (96, 40), (64, 96), (257, 100), (128, 128)])
def test_product_coefficient_tables_equal_oracle(roi, res):
    mil = importlib.import_module(PKG)
    b0, c0 = IO.pil_bilinear_coeffs(roi, res)
    b1, c1 = mil.ingest.pil_bilinear_coeffs(roi, res)
    assert np.array_equal(b0, b1) and np.array_equal(c0, c1)
    assert int(c1.sum(axis=1).min()) > (1 << 22) - 16 and int(c1.sum(axis=1).max()) < (1 << 22) + 16


def test_ingest_refuses_cpu_tensors():
    mil = importlib.import_module(PKG)
    ing = mil.TileIngest(64, 32)
    with pytest.raises(RuntimeError):
        ing(torch.zeros(2, 64, 64, 3, dtype=torch.uint8))
