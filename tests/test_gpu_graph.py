"""-m gpu: the CUDA-graph replay of a training step (graph.GraphedStep) against the eager path -- the same kernels in the
same order, so an eval-mode step (every tile through the CNN, no dropout: deterministic) must reproduce the eager
results and parameter updates bit for bit (reference step: gbm/classify_combined.py:432-454 with optim.Adam, :519)."""
import pytest
import torch

from oracle import synth
from tests import gpu_ops as G
from tests.helpers import golden_weights

pytestmark = pytest.mark.gpu


def make(precision="bf16", train=False):
    mil = G.pkg()
    net = mil.Attention(n_classes=3).cuda()
    net.load_state_dict(golden_weights())
    net.precision = precision
    net.train(train)
    return mil, net


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graphed_eval_steps_equal_eager_steps_bit_for_bit(precision):
    mil, a = make(precision)
    _, b = make(precision)
    oa, ob = mil.FusedAdam(a, lr=1e-3), mil.FusedAdam(b, lr=1e-3)
    bags = [torch.from_numpy(synth.make_bag(24, 64, seed=s)).cuda() for s in (3, 4, 5)]
    labels = [torch.tensor([s % 3]).cuda() for s in (3, 4, 5)]
    step = mil.GraphedStep(b, 24, 64, optimizer=ob)
    # the capture (warm-up steps on a zero bag included) must leave parameters and Adam state alone
    assert torch.equal(oa._flat, ob._flat) and float(ob._m.abs().sum()) == 0.0 and ob._t == 0
    for bag, y in zip(bags, labels):
        oa.zero_grad()
        out_a = a(bag, y)
        out_a["loss"].backward()
        oa.step()
        out_b = step(bag, y)
        for k in ("loss", "Aterm", "Mterm", "y_pred", "Fterm", "y_pred_hat", "l2"):
            assert torch.equal(out_a[k], out_b[k]), k
        assert torch.equal(oa._gflat, ob._gflat)
        assert torch.equal(oa._flat, ob._flat), "parameters after the step"
    assert oa._t == ob._t == 3
    assert torch.equal(oa._m, ob._m) and torch.equal(oa._v, ob._v)


def test_graphed_step_without_optimizer_leaves_gradients():
    mil, a = make()
    _, b = make()
    bag = torch.from_numpy(synth.make_bag(16, 96, seed=7)).cuda()
    y = torch.tensor([2]).cuda()
    step = mil.GraphedStep(b, 16, 96)
    out = a(bag, y)
    out["loss"].backward()
    step(bag, y)
    step(bag, y)        # a replay starts from zeroed gradients, it does not accumulate
    for (nm, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p.grad, q.grad), nm


def test_graphed_train_mode_redraws_subsample_and_dropout():
    mil, net = make(train=True)
    opt = mil.FusedAdam(net, lr=1e-4)
    bag = torch.from_numpy(synth.make_bag(60, 64, seed=9)).cuda()
    u8 = ((bag * 0.5 + 0.5) * 255).round().to(torch.uint8)
    step = mil.GraphedStep(net, 60, 64, optimizer=opt, bag_dtype=torch.uint8)
    seen, losses = set(), []
    for _ in range(4):
        out = step(u8, torch.tensor([1]).cuda())
        torch.cuda.synchronize()
        assert out["Fterm"].shape == (12, 80) and bool(torch.isfinite(out["loss"]))   # int(0.2 * 60) tiles
        seen.add(tuple(step.idx.tolist()))
        losses.append(float(out["loss"].detach()))
    assert len(seen) == 4, "every replay must see a fresh randperm subsample (gbm/model.py:193)"
    assert len(set(losses)) == 4 and opt._t == 4
    with pytest.raises(ValueError):
        step(u8[:59], None)


def test_graphed_step_rejects_other_optimizers():
    mil, net = make()
    with pytest.raises(TypeError):
        mil.GraphedStep(net, 8, 64, optimizer=torch.optim.Adam(net.parameters()))
