"""CPU-side tests: the C-ABI library loads and exports every symbol include/mil_b200.h declares, the host
mirror of the reference surface (constructor, init, state-dict keys) matches the reference, the product path
fails loudly without a GPU, and the multi-rank host logic works over gloo (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import gpu_ops as G
from tests.helpers import golden_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mil():
    m = G.pkg()
    m.build()
    return m


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mil_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mil_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(mil):
    lib = ctypes.CDLL(mil.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mil_b200.h but not exported"
    assert set(names) == set(mil._lib.PROTOTYPES), "ctypes prototypes out of sync with the header"


def test_param_table_matches_reference_state_dict(mil):
    table = mil._lib.param_table()
    gw = golden_weights()
    assert [t[0] for t in table] == list(gw.keys())
    for name, shape, off in table:
        assert tuple(gw[name].shape) == shape
    assert mil._lib.load().mil_param_total() == sum(v.numel() for v in gw.values()) == 640967


def test_workspace_sizes_are_sane(mil):
    lib = mil._lib.load()
    b64 = lib.mil_extractor_workspace_bytes(64, 224, 1)
    b128 = lib.mil_extractor_workspace_bytes(128, 224, 1)
    # ~4 MB per tile in bf16 + a fixed part (packed weights, the split-K partial records of a layer's weight gradients)
    assert 1.5e6 * 64 < b64 < 6e6 * 64 and 2.5e6 * 64 < b128 - b64 < 6e6 * 64
    assert lib.mil_extractor_workspace_bytes(64, 224, 0) > 1.2 * b64 - 5e8      # fp32: no tensor-core staging buffers
    assert lib.mil_extractor_workspace_bytes(0, 224, 1) == 0       # rejected, message set
    assert b"at least one tile" in lib.mil_last_error()


def test_constructor_reproduces_reference_init_bit_exactly(mil):
    """Same module tree built in the same order => same RNG stream as the reference constructor
    (gbm/model.py:118-181); tests/golden/weights.npz is the reference's own init under manual_seed(0)."""
    torch.manual_seed(0)
    net = mil.Attention(n_classes=3)
    sd = net.state_dict()
    gw = golden_weights()
    assert list(sd.keys()) == list(gw.keys())
    for k in gw:
        assert torch.equal(sd[k], gw[k]), k
    assert net.L == 80 and net.D == 40 and net.K == 3 and net.O == 1 and net.C == 3
    assert torch.equal(net.off_diag, 1 - torch.eye(3))
    assert net.loss.smoothing == 0.25 and net.loss.num_classes == 3 and net.loss.weight is None


def test_transfer_filter_and_strict_false_loading(mil):
    net = mil.Attention(n_classes=3, class_weights=torch.tensor([1.0, 2.0, 3.0]))
    gw = golden_weights()
    conv_only = {k: v for k, v in gw.items() if 'cnn' in k and 'conv' in k}   # gbm/classify_combined.py:531
    assert len(conv_only) == 50
    res = net.load_state_dict(conv_only, strict=False)
    assert not res.unexpected_keys
    assert torch.equal(net.cnn.module.layer4[2].conv2.weight, gw["cnn.module.layer4.2.conv2.weight"])
    net.reset_linear()
    assert float(net.attention.lin1.bias.detach().abs().max()) == 0.0


def test_product_path_fails_loudly_without_gpu(mil):
    net = mil.Attention(n_classes=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(4, 3, 64, 64), torch.tensor([1]))
    if not torch.cuda.is_available():
        lib = mil._lib.load()
        rc = lib.mil_head_stats(None, 4, None, 0, None, None)
        assert rc != 0 and b"CUDA" in lib.mil_last_error()


def test_product_code_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, G.PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


WORKER = r"""
import os, sys, importlib
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
mil = importlib.import_module({pkg!r})
from oracle import mil_oracle
from tests.helpers import golden_weights
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
grp = mil.BagGroup(dist.group.WORLD, seed=7)
# shard bookkeeping: uneven shards 5 and 3
n_local = 5 if rank == 0 else 3
assert grp.shard_sizes(n_local) == [5, 3]
assert grp.total(n_local) == 8
# two consecutive bags where only ONE rank's shard size changes (50/50 then 50/49): nothing is cached, every rank
# enters the all-gather every time, both see the new total (a stale cache on rank 0 used to desynchronise the ranks)
assert grp.total(50) == 100
assert grp.total(50 if rank == 0 else 49) == 99
assert grp.total(50 if rank == 0 else 49, hint=99) == 99
# train-mode subsample: the union over ranks is one global randperm prefix, disjoint, in shard range
grp2 = mil.BagGroup(dist.group.WORLD, seed=7)
n_bag = 40 if rank == 0 else 24
idx, k = grp2.subsample(n_bag, 0.2)
assert k == int(64 * 0.2)
cnt = torch.tensor([idx.numel()]); dist.all_reduce(cnt); assert int(cnt) == k
assert idx.numel() == 0 or (int(idx.min()) >= 0 and int(idx.max()) < n_bag)
ref = torch.randperm(64, generator=torch.Generator().manual_seed(7))[:k]
lo = 0 if rank == 0 else 40
assert torch.equal(idx, ref[(ref >= lo) & (ref < lo + n_bag)] - lo)
# the exchanged sums give the single-rank head result: emulate the kernels' partial sums with the oracle math
torch.manual_seed(3)
H = torch.randn(8, 80, dtype=torch.float64) * 3 + 1
p = {{k: v.double() for k, v in golden_weights().items()}}
full = mil_oracle.head_forward(p, H, torch.tensor([2]))
Hl = H[:5] if rank == 0 else H[5:]
stats = torch.cat([Hl.sum(0), (Hl * Hl).sum(0)])
grp.all_reduce_sum(stats)
mu = stats[:80] / 8; var = stats[80:] / 8 - mu * mu
Hz = (Hl - mu) / torch.sqrt(var + 1e-5) * p["context.bn.weight"] + p["context.bn.bias"]
raw = torch.tanh(Hz @ p["attention.lin1.weight"].t() + p["attention.lin1.bias"]) @ p["attention.lin2.weight"].t() + p["attention.lin2.bias"]
w = p["weight_mask"]
g = torch.sigmoid(-10 * w) * torch.nn.functional.softplus(raw) + torch.sigmoid(10 * w)
b = (torch.nn.functional.leaky_relu(torch.nn.functional.leaky_relu(Hl, 0.1) @ p["buffer.lin1.weight"].t() + p["buffer.lin1.bias"], 0.1)
     @ p["buffer.classifier.weight"].t() + p["buffer.classifier.bias"])
sums = torch.cat([g.sum(0), (g * b).sum(0)])
grp.all_reduce_sum(sums)
M = sums[3:] / sums[:3]
assert torch.allclose(M, full["Mterm"].view(3), rtol=1e-10, atol=1e-12), (M, full["Mterm"])
A_local = (g / sums[:3]).t()
A_full = full["Aterm"][:, :5] if rank == 0 else full["Aterm"][:, 5:]
assert torch.allclose(A_local, A_full, rtol=1e-10, atol=1e-14)
# gradient all-reduce: bucketed sum
flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
mil.BagGroup(dist.group.WORLD, grad_buckets=3).all_reduce_grads(flat)
assert torch.equal(flat, torch.arange(1000, dtype=torch.float32) * 3)
# multi-slide data parallelism (BASELINE configs[4]): every rank its own bag -- bag-wide sums stay local, sizes are
# local, only the gradient all-reduce crosses ranks and yields the SUM of the per-slide gradients
sg = mil.SlideGroup(dist.group.WORLD, grad_buckets=2)
assert sg.world == 2 and not sg.shares_bag and grp.shares_bag
assert sg.total(5 if rank == 0 else 3) == (5 if rank == 0 else 3)
loc = torch.full((4,), float(rank + 1), dtype=torch.float64)
sg.all_reduce_sum(loc)
assert torch.equal(loc, torch.full((4,), float(rank + 1), dtype=torch.float64))
torch.manual_seed(11 + rank)
want = torch.randperm(30)[:6]
torch.manual_seed(11 + rank)
got, kk = sg.subsample(30, 0.2)
assert kk == 6 and torch.equal(got, want)
gslide = torch.arange(500, dtype=torch.float32) * (10 ** rank)
sg.all_reduce_grads(gslide)
assert torch.equal(gslide, torch.arange(500, dtype=torch.float32) * 11)
dist.barrier()
print("rank", rank, "ok", flush=True)
dist.destroy_process_group()
"""


def test_bag_group_over_gloo_world_size_2(tmp_path, mil):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, pkg=G.PKG))
    import socket
    for attempt in range(2):      # one retry: a port can be taken between probing and rendezvous
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
        procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=240)[0] for p in procs]
        if all(f"rank {r} ok" in o for r, o in enumerate(outs)):
            return
    raise AssertionError("\n".join(outs))


def test_dla_writer_and_export_refuse_cpu(tmp_path, mil):
    """SURVEY 8f N3 host side: `col row weight` lines like the reference's write_map (gbm/classify.py:211), and the
    device-side scaling has no CPU fallback."""
    raster = np.array([[0, 300], [300, 0], [600, 900]])
    path = str(tmp_path / "m.dla")
    mil.write_dla(path, raster, [0.25, 0.5, 1.0])
    rows = np.loadtxt(path)
    assert rows.tolist() == [[300.0, 0.0, 0.25], [0.0, 300.0, 0.5], [900.0, 600.0, 1.0]]
    with pytest.raises(ValueError):
        mil.write_dla(path, raster, [1.0, 2.0])
    with pytest.raises(RuntimeError):
        mil.minmax_normalize(torch.rand(3, 5))
    with pytest.raises(RuntimeError):
        mil.flatten_parameters(mil.Attention(n_classes=3))      # CPU module: no CPU fallback for the fused optimizer


def test_sharded_subsample_bookkeeping(mil):
    """BagGroup.subsample over a 3-rank bag without a process group (the shard sizes are injected): every rank
    draws the same permutation, keeps the indices of its shard, and ALL ranks raise when some shard gets none."""
    sizes = [30, 50, 20]
    picks = []
    for rank in range(3):
        g = mil.BagGroup()
        g.world, g.rank = 3, rank
        g._gen = torch.Generator().manual_seed(5)
        idx, k = g.subsample(sizes[rank], 0.2, sizes=sizes)
        assert k == 20
        assert int(idx.min()) >= 0 and int(idx.max()) < sizes[rank]
        picks.append(idx + sum(sizes[:rank]))
    allp = torch.cat(picks)
    assert allp.numel() == 20 and allp.unique().numel() == 20
    ref = torch.randperm(100, generator=torch.Generator().manual_seed(5))[:20]
    assert sorted(allp.tolist()) == sorted(ref.tolist())
    sizes = [3, 3, 3]
    errs = 0
    for rank in range(3):
        g = mil.BagGroup()
        g.world, g.rank = 3, rank
        g._gen = torch.Generator().manual_seed(1)
        try:
            g.subsample(3, 0.2, sizes=sizes)                      # one tile for three ranks
        except ValueError as e:
            errs += 1
            assert "without a tile" in str(e)
    assert errs == 3


def test_set_stage_follows_the_reference_schedule(mil):
    """gbm/classify_combined.py:110-138: learning rate and train / eval mode per epoch."""
    import torch
    net = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(net.parameters(), lr=1.0)
    want = {0: ("Warmup", 0.0002 / 10, True), 9: ("Warmup", 0.0002, True), 10: ("Main", 0.0002, True),
            149: ("Main", 0.0002, True), 150: ("Check", 0.0001, True), 250: ("Freeze", 0.00002, True),
            339: ("Freeze", 0.00002, True)}
    for epoch, (stage, lr, training) in want.items():
        got = mil.set_stage(opt, net, epoch, verbose=False)
        assert got[0] == stage and abs(got[1] - lr) < 1e-12 and net.training is training, (epoch, got)
        assert all(abs(g["lr"] - lr) < 1e-12 for g in opt.param_groups)
    mil.set_stage(opt, net, 200, test=True, verbose=False)
    assert net.training is False                      # check / freeze stages evaluate when test=True
    mil.set_stage(opt, net, 20, test=True, verbose=False)
    assert net.training is True
    assert mil.set_stage(opt, net, 341, verbose=False)[0] == "Stop"


def test_checkpoint_file_round_trip_and_transfer_filter(tmp_path, mil):
    """gbm/classify_combined.py:468-474 (save) and :521-535 (--ckpt / --transfer): the file holds the reference's
    state-dict keys; `transfer` restores only the extractor's convolutions."""
    import torch
    torch.manual_seed(3)
    a = mil.Attention(n_classes=3)
    torch.manual_seed(4)
    b = mil.Attention(n_classes=3)
    path = mil.save_checkpoint(str(tmp_path / "train_step-007.model"), a)
    blob = torch.load(path, weights_only=False)
    assert set(blob) == {"classifier"} and "cnn.module.layer1.0.conv1.weight" in blob["classifier"]
    mil.load_checkpoint(path, b)
    for (k, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(p, q), k
    torch.manual_seed(5)
    c = mil.Attention(n_classes=3)
    before = {k: v.clone() for k, v in c.state_dict().items()}
    mil.load_checkpoint(path, c, transfer=True)
    for k, v in c.state_dict().items():
        if "cnn" in k and "conv" in k:
            assert torch.equal(v, a.state_dict()[k]), k
        else:
            assert torch.equal(v, before[k]), k
