"""The ``spa3d::`` torch operator library (3dspa_code_b200/torch_ops.py): the framework registration of the kernels.

CPU part (no GPU needed): every operator is defined with a schema, has a fake implementation that propagates shapes / dtypes
on fake CUDA tensors, and ops.py's public names dispatch to it.  GPU part: torch.library.opcheck on the functional operators,
gradients of composed operators against float64 torch, and a TorchDispatchMode trace showing that the model's forward reaches
the kernels through the dispatcher.
"""
import importlib

import numpy as np
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from tests.helpers import SMALL_ARCH, make_inputs, product, rel_err, small_cfg

BF, F32 = torch.bfloat16, torch.float32


def test_every_operator_is_registered_with_a_schema():
    spa = product()
    t = importlib.import_module("3dspa_code_b200.torch_ops")
    for name in t.REGISTERED:
        op = getattr(torch.ops.spa3d, name).default
        assert str(op._schema).startswith(f"spa3d::{name}("), op._schema
    for name in ("gemm", "gemm_rmsnorm", "gemm_gelu", "gemm_gelu_bwd", "gemm_dw", "attention_fwd", "attention_bwd", "layernorm_fwd",
                 "layernorm_bwd", "embed_fused", "lift_sample", "loss_fwd", "loss_bwd"):
        assert "torch.ops.spa3d" in getattr(spa.ops, name).__doc__, name


def test_fake_tensor_propagation_without_a_gpu():
    """Shapes and dtypes of every functional operator on fake CUDA tensors (what torch.compile / export tracing sees)."""
    spa = product()
    ops = spa.ops
    with FakeTensorMode():
        dev = "cuda"
        a = torch.empty(309, 384, device=dev, dtype=BF)
        wt = torch.empty(2304, 384, device=dev, dtype=BF)
        bias = torch.empty(2304, device=dev)
        sc = torch.empty(96, device=dev)
        y = ops.gemm(a, wt, bias, out_dtype=F32)
        assert y.shape == (309, 2304) and y.dtype == F32 and y.device.type == "cuda"
        qkv, rstd = ops.gemm_rmsnorm(a, wt, 96, 768, 768, sc, sc, save_rstd=True)
        assert qkv.shape == (309, 2304) and qkv.dtype == BF and rstd.shape == (309, 16)
        z, h = ops.gemm_gelu(a, wt, bias)
        assert z.shape == h.shape == (309, 2304)
        dz = ops.gemm_gelu_bwd(h, torch.empty(384, 2304, device=dev, dtype=BF), a)
        assert dz.shape == (309, 384)
        x = torch.empty(309, 384, device=dev)
        xn, mean, rs = ops.layernorm_fwd(x, torch.empty(384, device=dev), BF, stats=True)
        assert xn.dtype == BF and mean.shape == rs.shape == (309,)
        xn0 = ops.layernorm_fwd(x, torch.empty(384, device=dev), BF, rows=3, ldx=103 * 384, d=384)
        assert xn0.shape == (3, 384)
        o, st = torch.ops.spa3d.attention(qkv[:302, :768], qkv[:302, 768:1536], qkv[:302, 1536:], None, 2, 8, 151, 151, 96)
        assert o.shape == (302, 768) and st.shape == (2, 8, 151, 2)
        tok, acat = torch.ops.spa3d.embed_fused(torch.empty(300, 3, device=dev), torch.empty(300, 768, device=dev), torch.empty(300, 256, device=dev),
                                                torch.empty(384, 1280, device=dev, dtype=BF), torch.empty(384, device=dev),
                                                torch.empty(1, 384, device=dev), 150, 32, 1.0)
        assert tok.shape == (302, 384) and acat.shape == (302, 1280) and acat.dtype == BF
        xyz, df, zf = ops.lift_sample(torch.empty(5, 7, 2, device=dev), depth=torch.empty(7, 30, 40, 1, device=dev),
                                      dino=torch.empty(7, 3, 4, 768, device=dev), video_hw=(30, 40))
        assert xyz.shape == (5, 7, 3) and df.shape == (5, 7, 768) and zf.shape == (5, 7, 256)
        sums = torch.ops.spa3d.loss_sums(torch.empty(8, 600, device=dev), torch.empty(8, 150, 3, device=dev), torch.empty(8, 150, 1, device=dev), 150)
        assert sums.shape == (3,)


def test_workspace_queries_and_dispatch_counters():
    spa = product()
    lib = importlib.import_module("3dspa_code_b200._lib").lib()
    assert lib.spa3d_gemm_workspace_bytes(1000, 384, 384) == 0
    assert lib.spa3d_sumsq_workspace_bytes() == 4096
    assert lib.spa3d_attention_bwd_workspace_bytes(2048, 8, 151) == 2048 * 8 * 151 * 4
    assert lib.spa3d_attention_stats_bytes(2, 8, 151) == 2 * 8 * 151 * 8
    assert lib.spa3d_layernorm_bwd_workspace_bytes(296, 384) == 296 * 384 * 4
    st = spa.ops.stats()
    assert {"gemm_tcgen05", "gemm_bf16_fallback", "attention_tcgen05", "attention_bf16_fallback", "gemm_dw_bf16_fallback"} <= set(st)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def spa():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return product()


@pytest.mark.gpu
def test_opcheck_functional_operators(spa):
    """torch.library.opcheck: schema correctness, fake-vs-real agreement, autograd registration, AOT dispatch."""
    from torch.library import opcheck

    torch.manual_seed(0)
    dev = "cuda"
    a = torch.randn(200, 64, device=dev).to(BF)
    wt = (torch.randn(128, 64, device=dev) / 8).to(BF)
    bias = torch.randn(128, device=dev)
    checks = ("test_schema", "test_autograd_registration", "test_faketensor")
    opcheck(torch.ops.spa3d.gemm.default, (a, wt, bias, 0, None, F32, 0), test_utils=checks)
    opcheck(torch.ops.spa3d.gemm.default, (a.requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True), 0, None, BF, 0),
            test_utils=checks)
    a = a.detach()
    opcheck(torch.ops.spa3d.gemm_gelu.default, (a, wt, bias, 0, False), test_utils=checks)
    sc = torch.rand(32, device=dev) + 0.5
    opcheck(torch.ops.spa3d.gemm_rmsnorm.default, (a, wt, 32, 64, 32, sc, sc, True, 0), test_utils=checks)
    x = torch.randn(200, 384, device=dev)
    opcheck(torch.ops.spa3d.layernorm_fwd.default, (x, torch.rand(384, device=dev) + 0.5, BF, -1, -1, -1, True), test_utils=checks)
    q = torch.randn(2 * 37, 128, device=dev).to(BF)
    opcheck(torch.ops.spa3d.attention.default, (q, q.clone(), q.clone(), None, 2, 2, 37, 37, 64), test_utils=checks)
    head = torch.randn(6, 40, device=dev)
    opcheck(torch.ops.spa3d.loss_sums.default, (head, torch.randn(6, 10, 3, device=dev), (torch.rand(6, 10, 1, device=dev) < 0.7).float(), 10),
            test_utils=checks)
    tr2 = torch.rand(5, 4, 2, device=dev) * 20
    opcheck(torch.ops.spa3d.lift_sample.default, (tr2, torch.rand(4, 30, 40, 1, device=dev) + 0.5, torch.randn(4, 3, 4, 16, device=dev), 30, 40,
                                                  None, F32, 256, True, True, True), test_utils=checks)


@pytest.mark.gpu
def test_composed_operators_differentiate_like_torch(spa):
    """A caller composing the registered operators directly (no training engine) gets the right gradients:
    y = LN(x) -> q,k,v with per-head RMSNorm -> attention -> out-projection + residual -> GELU MLP, fp32 path, vs float64 torch."""
    torch.manual_seed(1)
    dev = "cuda"
    ops = spa.ops
    B, L, H, Dh, d = 3, 20, 2, 32, 48
    A = H * Dh
    x = torch.randn(B * L, d, device=dev, requires_grad=True)
    P = {"ln": torch.rand(d, device=dev) + 0.5, "wqkv": torch.randn(3 * A, d, device=dev) / 7, "sq": torch.rand(Dh, device=dev) + 0.5,
         "sk": torch.rand(Dh, device=dev) + 0.5, "wo": torch.randn(d, A, device=dev) / 8, "bo": torch.randn(d, device=dev) * 0.1,
         "w1": torch.randn(96, d, device=dev) / 7, "b1": torch.randn(96, device=dev) * 0.1, "w2": torch.randn(d, 96, device=dev) / 10,
         "b2": torch.randn(d, device=dev) * 0.1}
    for v in P.values():
        v.requires_grad_(True)
    mask = (torch.rand(B, L, device=dev) < 0.8).to(torch.uint8)
    mask[:, 0] = 1

    xn, _, _ = ops.layernorm_fwd(x, P["ln"], F32, stats=True)
    qkv, _ = ops.gemm_rmsnorm(xn, P["wqkv"], Dh, A, A, P["sq"], P["sk"], save_rstd=True)
    o, _ = torch.ops.spa3d.attention(qkv[:, :A].contiguous(), qkv[:, A : 2 * A].contiguous(), qkv[:, 2 * A :].contiguous(), mask, B, H, L, L, Dh)
    a = ops.gemm(o, P["wo"], P["bo"], residual=x, out_dtype=F32)
    z, h = ops.gemm_gelu(a, P["w1"], P["b1"])
    y = ops.gemm(h, P["w2"], P["b2"], residual=a, out_dtype=F32)
    R = torch.randn_like(y)
    (y * R).sum().backward()
    got = {k: v.grad.clone() for k, v in P.items()}
    got["x"] = x.grad.clone()

    def ref_fn(x, P):
        mu = x.mean(-1, keepdim=True)
        var = (x * x).mean(-1, keepdim=True) - mu * mu
        xn = (x - mu) * torch.rsqrt(var + 1e-6) * P["ln"]
        qkv = xn @ P["wqkv"].t()
        q, k, v = (qkv[:, i * A : (i + 1) * A].reshape(B, L, H, Dh) for i in range(3))
        rms = lambda t, s: t * torch.rsqrt((t * t).mean(-1, keepdim=True) + 1e-6) * s
        q, k = rms(q, P["sq"]) / Dh ** 0.5, rms(k, P["sk"])
        w = torch.einsum("bqhd,bkhd->bhqk", q, k)
        w = torch.where(mask.view(B, 1, 1, L) != 0, w, torch.full_like(w, torch.finfo(torch.float32).min))
        o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(w, -1), v).reshape(B * L, A)
        a = o @ P["wo"].t() + P["bo"] + x
        z = a @ P["w1"].t() + P["b1"]
        h = 0.5 * z * (1 + torch.tanh(0.7978845608028654 * (z + 0.044715 * z ** 3)))
        return h @ P["w2"].t() + P["b2"] + a

    xd = x.detach().double().requires_grad_(True)
    Pd = {k: v.detach().double().requires_grad_(True) for k, v in P.items()}
    yd = ref_fn(xd, Pd)
    assert rel_err(y, yd) < 1e-5
    (yd * R.double()).sum().backward()
    for k, v in Pd.items():
        assert rel_err(got[k], v.grad) < 1e-4, (k, rel_err(got[k], v.grad))
    assert rel_err(got["x"], xd.grad) < 1e-4


@pytest.mark.gpu
def test_model_forward_goes_through_the_dispatcher(spa):
    """TorchDispatchMode sees the spa3d:: operators of a forward pass (i.e. the model calls the kernels as registered ops)."""
    from oracle import model as om
    from torch.utils._python_dispatch import TorchDispatchMode

    seen = {}

    class Log(TorchDispatchMode):
        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            name = str(func)
            if name.startswith("spa3d."):
                seen[name] = seen.get(name, 0) + 1
            return func(*args, **(kwargs or {}))

    c = small_cfg()
    model = spa.TrackAutoEncoder3D(**{k: getattr(c, k) for k in om.Config3D.__dataclass_fields__})
    inp, noise = make_inputs(c, B=1, N=6, Q=4)
    variables = model.init(3, inp, arch=SMALL_ARCH)
    ref = model.apply(variables, inp, noise=noise, precision="bf16")
    with Log():
        got = model.apply(variables, inp, noise=noise, precision="bf16")
    assert torch.equal(got.tracks, ref.tracks)
    # under no_grad the engine takes the out-variants (no autograd key to cross); with grad mode on, the functional operators
    for base in ("gemm", "gemm_rmsnorm", "layernorm_fwd"):
        assert seen.get(f"spa3d.{base}_out.default", 0) + seen.get(f"spa3d.{base}.default", 0) > 0, (base, seen)
    assert seen.get("spa3d.attention_fwd.default", 0) > 0, seen


@pytest.mark.gpu
def test_dispatch_counters_and_weight_shadows(spa):
    ops = spa.ops
    ops.stats(reset=True)
    a = torch.randn(256, 384, device="cuda").to(BF)
    wt = torch.randn(768, 384, device="cuda").to(BF)
    ops.gemm(a, wt)
    s = ops.stats()
    assert s["gemm_tcgen05"] == 1 and s["gemm_bf16_fallback"] == 0
    ops.gemm(a[:, :380], wt[:, :380])          # K = 380 breaks the 16-byte row rule: computed on SIMT and counted
    s = ops.stats(reset=True)
    assert s["gemm_simt"] == 1 and s["gemm_bf16_fallback"] == 1
    assert ops.stats()["gemm_simt"] == 0
    w = torch.randn(70, 133, device="cuda")
    d, dt_ = torch.empty(70, 133, device="cuda", dtype=BF), torch.empty(133, 70, device="cuda", dtype=BF)
    ops.shadow_weights(w, d, dt_)
    assert torch.equal(d, w.to(BF)) and torch.equal(dt_, w.to(BF).t())
    g = torch.ones(1000, device="cuda")
    ops.fill_zero(g)
    assert not g.any()
