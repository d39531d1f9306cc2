"""Backward / optimiser parity on a B200: gradients from the hand-written backward kernels vs
torch autograd through the oracle, the AdamW update vs the oracle restatement of optax, and
micro-batch (= data-parallel shard) equivalence."""
import importlib

import numpy as np
import pytest
import torch

from oracle import model as om
from tests.helpers import SMALL_ARCH, make_inputs, product, rel_err, small_cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spa():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return product()


def _setup(spa, seed=3, B=2, N=9, Q=5):
    c = small_cfg()
    fields = {k: getattr(c, k) for k in om.Config3D.__dataclass_fields__}
    model = spa.TrackAutoEncoder3D(**fields)
    inp, noise = make_inputs(c, B=B, N=N, Q=Q, targets=True, seed=seed)
    inp["boundary_frame"] = np.array([c.num_output_frames - 1] + [c.num_output_frames] * (B - 1), np.int32)
    tree = model.init(seed, inp, arch=SMALL_ARCH)["params"]
    om._randomize(tree, np.random.RandomState(seed))
    return c, model, tree, inp, noise


def oracle_grads(c, tree, inp, noise, dtype=torch.float64):
    p = om.to_torch(tree, dtype, requires_grad=True)
    ci = om.cast_inputs(inp, dtype)
    res = om.forward_3d(p, c, ci, torch.as_tensor(noise).to(dtype), True)
    loss = om.compute_loss_3d(res, ci)
    loss["total_loss"].backward()
    return loss, {k: v.grad for k, v in om.flatten(p).items()}


def test_gradients_match_oracle_autograd_fp32(spa):
    """Real loss (train.py:96-129), fp32 path: every parameter leaf vs autograd through the oracle."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c, model, tree, inp, noise = _setup(spa)
    loss_ref, gref = oracle_grads(c, tree, inp, noise)
    store = te.ParamStore(tree, "fp32")
    eng = te.TrainEngine(model, store)
    denom = max(float(inp["query_tracks_visible"].sum()), 1.0)
    sums = eng.loss_and_backward(inp, noise, denom)
    pos, bce = float(sums[0]) / denom, float(sums[1]) / denom
    assert abs(pos - float(loss_ref["position_loss"])) < 1e-4 * float(loss_ref["position_loss"])
    assert abs(bce - float(loss_ref["visible_loss"])) < 1e-4 * float(loss_ref["visible_loss"])
    got = spa.params.flatten(store.grad_tree())
    assert got.keys() == gref.keys()
    bad = {k: rel_err(got[k], g) for k, g in gref.items() if rel_err(got[k], g) > 2e-3}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = b.detach().double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_backward_kernels_linear_functional(spa, precision):
    """Backward chain in isolation: d(sum(head_out * R))/dparams for a fixed random R.  (The L1 loss
    gradient is sign(pred - target): a bf16-sized forward error flips isolated signs, which says
    nothing about the backward kernels, so the bf16 path is checked with a smooth functional.)"""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c, model, tree, inp, noise = _setup(spa, seed=11)
    B, Q, T = inp["query_points"].shape[0], inp["query_points"].shape[1], c.num_output_frames
    R = torch.randn(B * Q, 4 * T, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    p = om.to_torch(tree, torch.float64, requires_grad=True)
    ci = om.cast_inputs(inp, torch.float64)
    res = om.forward_3d(p, c, ci, torch.as_tensor(noise).double(), False)
    head = torch.cat([res.tracks[..., 0], res.tracks[..., 1], res.tracks[..., 2], res.visible_logits[..., 0]], dim=-1)
    (head.reshape(B * Q, 4 * T) * R).sum().backward()
    gref = {k: v.grad for k, v in om.flatten(p).items()}
    store = te.ParamStore(tree, precision)
    eng = te.TrainEngine(model, store)
    with torch.enable_grad():
        out = eng.forward_train(inp, noise, discretize=False)
    out.backward(R.float().cuda())
    got = spa.params.flatten(store.grad_tree())
    if precision == "fp32":
        bad = {k: rel_err(got[k], g) for k, g in gref.items() if rel_err(got[k], g) > 1e-3}
    else:
        # bf16: ~0.4 % rounding per op over a ~40-op chain; judge direction and size of every leaf
        bad = {k: (_cos(got[k], g), rel_err(got[k], g)) for k, g in gref.items()
               if _cos(got[k], g) < 0.985 or rel_err(got[k], g) > 0.2}
    assert not bad, sorted(bad.items(), key=lambda kv: str(kv[1]))[:8]


def test_microbatching_equals_full_batch(spa):
    """Gradient accumulation over clips (what a DP shard does) equals the full-batch gradient."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c, model, tree, inp, noise = _setup(spa, seed=5, B=4)
    denom = max(float(inp["query_tracks_visible"].sum()), 1.0)
    s1 = te.ParamStore(tree, "fp32")
    te.TrainEngine(model, s1).loss_and_backward(inp, noise, denom)
    s2 = te.ParamStore(tree, "fp32")
    e2 = te.TrainEngine(model, s2)
    for s in range(0, 4, 1):
        mb = {k: v[s : s + 1] for k, v in inp.items()}
        e2.loss_and_backward(mb, noise[s : s + 1], denom)
    assert rel_err(s2.grad, s1.grad) < 1e-4


def test_train_step_matches_oracle_adamw(spa):
    """clip_by_global_norm(1.0) -> adamw(lr(t), wd=0.01) with the warm-up/cosine schedule
    (train.py:41-57,239-243): the device update, fed its own gradients, equals the oracle's optax
    restatement; losses and learning rates match the oracle forward."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c, model, tree, inp, noise = _setup(spa, seed=7)
    trainer = te.Trainer(model, tree, precision="fp32", base_lr=1e-3, warmup_steps=2, micro_batch=1)
    keys = sorted(om.flatten(tree))
    params = [torch.as_tensor(np.asarray(om.flatten(tree)[k])).double() for k in keys]
    m = [torch.zeros_like(t) for t in params]
    v = [torch.zeros_like(t) for t in params]
    for step in range(3):
        before = {k: t.clone() for k, t in zip(keys, params)}
        p = {k: t.clone() for k, t in zip(keys, params)}
        res = om.forward_3d(spa.params.unflatten(p), c, om.cast_inputs(inp, torch.float64), torch.as_tensor(noise).double(), True)
        loss = om.compute_loss_3d(res, om.cast_inputs(inp, torch.float64))
        log = trainer.train_step(inp, noise)
        lr = om.learning_rate(step, 1e-3, 2, 1000000)
        assert abs(log["learning_rate"] - lr) < 1e-12
        assert abs(log["total_loss"] - float(loss["total_loss"])) < 1e-3 * abs(float(loss["total_loss"]))
        g = spa.params.flatten(trainer.store.grad_tree())
        grads = [torch.as_tensor(np.asarray(g[k])).double() for k in keys]
        gn = om.adamw_step(params, grads, m, v, step + 1, lr)
        assert abs(log["grad_norm"] - float(gn)) < 1e-4 * float(gn)
        got = spa.params.flatten(trainer.store.tree())
        worst = max(rel_err(got[k], t) for k, t in zip(keys, params))
        assert worst < 1e-5, (step, worst)
        if step == 0:  # lr(0) = 0: the warm-up starts from zero, weights must not move
            assert all(np.array_equal(got[k], before[k].float().numpy()) for k in keys)
    moved = max(float(np.abs(got[k] - np.asarray(om.flatten(tree)[k])).max()) for k in keys)
    assert moved > 1e-5


@pytest.mark.parametrize("precision,real_widths", [("fp32", False), ("bf16", False), ("bf16", True)])
def test_gradient_regions_declared_final_are_never_written_again(spa, precision, real_widths):
    """Data-parallel overlap sends a bucket as soon as every gradient below the frontier is final.  Completion is tracked per
    backward block (ParamStore.finish); here, on one GPU, every region is snapshotted when it is declared final and must be
    bit-identical at the end of the backward pass, every parameter must have been reported, and the frontier must reach the
    end of the buffer.  Real widths: the bias-gradient reductions fused into LayerNorm backward kernels cross block borders."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    if real_widths:
        c = om.Config3D()
        inp, noise = make_inputs(c, B=2, N=6, Q=4, targets=True, seed=17)
        model = spa.TrackAutoEncoder3D()
        tree = model.init(17, inp)["params"]
    else:
        c, model, tree, inp, noise = _setup(spa, seed=13, B=2)
    tr = te.Trainer(model, tree, precision=precision, micro_batch=1)
    tr.debug_frontier = []
    tr.train_step(inp, noise)
    assert tr.unfinished == [], tr.unfinished
    regions = tr.debug_frontier
    assert len(regions) > 4 and regions[0][0] == 0 and regions[-1][1] == tr.store.total
    for (lo, hi, snap), nxt in zip(regions, regions[1:] + [None]):
        assert hi > lo and (nxt is None or nxt[0] == hi)
        assert torch.equal(tr.store.grad[lo:hi], snap), (lo, hi)


def test_training_with_a_feature_missing_from_the_batch(spa):
    """The tree has dino and depth projections, the batch has no depth features: that Dense is skipped
    (track_autoencoder_3d.py:140,145) - no bias, zero gradient for its kernel and bias; everything else as the oracle."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c, model, tree, inp, noise = _setup(spa, seed=19)
    inp = {k: v for k, v in inp.items() if k != "depth_features"}
    loss_ref, gref = oracle_grads(c, tree, inp, noise)
    store = te.ParamStore(tree, "fp32")
    eng = te.TrainEngine(model, store)
    denom = max(float(inp["query_tracks_visible"].sum()), 1.0)
    sums = eng.loss_and_backward(inp, noise, denom)
    assert abs(float(sums[0]) / denom - float(loss_ref["position_loss"])) < 1e-4 * float(loss_ref["position_loss"])
    got = spa.params.flatten(store.grad_tree())
    for k, g in gref.items():
        if g is None:
            assert k.startswith("depth_projection") and not np.asarray(got[k]).any(), k
        else:
            assert rel_err(got[k], g) < 2e-3, (k, rel_err(got[k], g))
