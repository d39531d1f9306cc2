"""Per-kernel parity on a B200: every C-ABI entry point against the oracle / plain torch fp64.

Integer / index / float32-op-order work is compared bit-exactly; floating-point kernels with the
tolerance written next to each assert.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import lifting as olift
from oracle import model as om
from tests.helpers import product, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spa():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return product()


def dev(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)) if not isinstance(a, torch.Tensor) else a
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


# ---- K0 ------------------------------------------------------------------------------------
def _lift_case(spa, g, prefix=""):
    k = lambda n: g[prefix + n]
    depth, dino, tracks = k("depth"), k("dino"), k("tracks")
    intr = k("intrinsics")
    intr = None if np.isnan(intr).any() else tuple(float(v) for v in intr)
    T, H, W = depth.shape[:3]
    L = spa.lifting
    np.testing.assert_array_equal(L.lift_2d_to_3d(tracks, depth, intr), k("xyz"))
    np.testing.assert_array_equal(L.sample_dino_features_for_tracks(dino, tracks, (T, H, W, 3)), k("dino_feat"))
    np.testing.assert_array_equal(L.sample_depth_features_for_tracks(depth, tracks), k("depth_feat"))
    xyz, df, zf = L.lift_and_sample(tracks, depth, dino, (T, H, W, 3), intr)
    np.testing.assert_array_equal(xyz.cpu().numpy(), k("xyz"))
    np.testing.assert_array_equal(df.cpu().numpy(), k("dino_feat"))
    np.testing.assert_array_equal(zf.cpu().numpy(), k("depth_feat"))


def test_lift_sample_golden_bit_exact(spa, golden_dir):
    """CUDA gather vs fixtures produced by the reference's own NumPy loops (inference.py:287-447)."""
    _lift_case(spa, np.load(os.path.join(golden_dir, "lifting_small.npz")))
    g = np.load(os.path.join(golden_dir, "lifting_edges.npz"))
    for case in ("integer", "far_out", "one_px", "intrinsics"):
        _lift_case(spa, g, case + "/")


def test_lift_sample_config4_shape_vs_oracle(spa):
    """Larger case (oracle finishes in seconds): 512 tracks x 20 frames on a 518x518 video."""
    rs = np.random.RandomState(0)
    N, T, H, W = 512, 20, 518, 518
    depth = rs.uniform(0.5, 10, (T, H, W, 1)).astype(np.float32)
    dino = rs.standard_normal((T, 37, 37, 768)).astype(np.float32)
    tr = np.stack([rs.uniform(-20, W + 20, (N, T)), rs.uniform(-20, H + 20, (N, T))], -1).astype(np.float32)
    xyz, df, zf = spa.lifting.lift_and_sample(tr, depth, dino, (T, H, W, 3))
    np.testing.assert_array_equal(xyz.cpu().numpy(), olift.lift_2d_to_3d(tr, depth))
    np.testing.assert_array_equal(df.cpu().numpy(), olift.sample_dino_features_for_tracks(dino, tr, (T, H, W, 3)))
    np.testing.assert_array_equal(zf.cpu().numpy(), olift.sample_depth_features_for_tracks(depth, tr))
    # bf16 output variant (fast path): equals the rounded f32 result
    _, dfb, _ = spa.lifting.lift_and_sample(tr, depth, dino, (T, H, W, 3), out_dtype=torch.bfloat16)
    assert torch.equal(dfb, df.to(torch.bfloat16))


@pytest.mark.parametrize("D,Cd,out_dtype,with_depth", [(128, 256, torch.float32, True), (256, 4, torch.float32, True),
                                                       (384, 8, torch.bfloat16, True), (768, 256, torch.float32, False)])
def test_lift_sample_binned_equals_per_point_kernel_and_oracle(spa, D, Cd, out_dtype, with_depth):
    """The cell-binned gather (corner rows read once per occupied patch cell; spa3d_lift_sample_ws) against the per-point
    kernel (spa3d_lift_sample, no workspace) and the NumPy restatement, bit for bit: crowded cells (more than 32 points in one
    cell: several batches), empty cells, points far outside the frame (clamped corners), both output dtypes, no depth."""
    import ctypes
    rs = np.random.RandomState(D + Cd)
    N, T, H, W, Hp, Wp = 150, 5, 40, 56, 3, 4
    tr = np.stack([rs.uniform(-30, W + 30, (N, T)), rs.uniform(-30, H + 30, (N, T))], -1).astype(np.float32)
    tr[:70, :, 0] = rs.uniform(15, 27, (70, T))      # 70 points of every frame inside one cell (cell width 14 px)
    tr[:70, :, 1] = rs.uniform(14, 26, (70, T))
    tr[70:80] = np.float32([W - 1, H - 1])           # exactly on the last pixel
    depth = rs.uniform(0.5, 10, (T, H, W, 1)).astype(np.float32)
    dino = rs.standard_normal((T, Hp, Wp, D)).astype(np.float32)
    dev = torch.device("cuda")
    t_tr, t_dino = torch.as_tensor(tr, device=dev), torch.as_tensor(dino, device=dev)
    t_depth = torch.as_tensor(depth, device=dev) if with_depth else None
    xyz, df, zf = spa.ops.lift_sample(t_tr, t_depth, t_dino, (H, W), None, out_dtype, Cd)      # binned (D % 128 == 0)
    # per-point kernel through the plain entry point
    xyz2 = torch.full((N, T, 3), 7.0, device=dev) if with_depth else None
    df2 = torch.empty(N, T, D, device=dev, dtype=out_dtype)
    zf2 = torch.empty(N, T, Cd, device=dev, dtype=out_dtype) if with_depth else None
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    rc = spa._lib.lib().spa3d_lift_sample(P(t_tr), P(t_depth), P(t_dino), P(xyz2), P(df2), P(zf2), spa.ops.dt(df2), N, T,
                                          H if with_depth else 0, W if with_depth else 0, Hp, Wp, D, Cd, H, W, None,
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(df, df2)
    want = torch.as_tensor(olift.sample_dino_features_for_tracks(dino, tr, (T, H, W, 3)), device=dev)
    assert torch.equal(df, want.to(out_dtype))
    if with_depth:
        assert torch.equal(xyz, xyz2) and torch.equal(zf, zf2)
        np.testing.assert_array_equal(xyz.cpu().numpy(), olift.lift_2d_to_3d(tr, depth))
        if Cd == 256:
            np.testing.assert_array_equal(zf.float().cpu().numpy(), olift.sample_depth_features_for_tracks(depth, tr))
        # geometry only (the maps path: no per-point patch features, 4-wide depth feature): one thread per point, same bits
        xyz3, none3, zf3 = spa.ops.lift_sample(t_tr, t_depth, None, (H, W), None, torch.float32, 4, True, False, True)
        assert none3 is None and torch.equal(xyz3, xyz)
        want3 = olift.sample_depth_features_for_tracks(depth, tr)[..., :4]
        np.testing.assert_array_equal(zf3.cpu().numpy(), want3)
    else:
        assert xyz is None and zf is None


# ---- Fourier features ---------------------------------------------------------------------
def test_fourier_exact_matches_oracle(spa):
    rs = np.random.RandomState(1)
    x = rs.uniform(-1, 1, (1000, 3)).astype(np.float32)
    out = torch.empty(1000, 192, device="cuda")
    spa.ops.fourier_features(dev(x), out, 32, 1.0, exact=True)
    ref = om.sinusoidal_embedding(torch.from_numpy(x), 32)
    diff = (out.cpu() - ref).abs()
    # correctly-rounded sine on both sides: at most isolated 1-ulp differences from the two libms
    assert float(diff.max()) <= 1.2e-7 and float((diff > 0).float().mean()) < 1e-3
    # time coordinate + read-out row layout + tail zero
    T = 7
    xt = rs.uniform(-1, 1, (3 * T, 3)).astype(np.float32)
    out = torch.zeros(3 * (T + 1), 256, device="cuda")
    spa.ops.fourier_features(dev(xt), out, 32, 1.0, append_time=T, exact=True, out_row_group=T)
    fr = (torch.arange(T, dtype=torch.float32) / T).repeat(3)[:, None]
    ref = om.sinusoidal_embedding(torch.cat([torch.from_numpy(xt), fr], 1), 32)
    got = out.view(3, T + 1, 256)
    assert torch.all(got[:, 0] == 0)
    assert float((got[:, 1:].reshape(3 * T, 256).cpu() - ref).abs().max()) <= 1.2e-7
    out = torch.empty(5, 4 * 64, device="cuda")
    spa.ops.fourier_features(dev(x[:5]), out, 32, 1.0, tail_zero=True, exact=True)
    assert torch.all(out[:, 192:224] == 0) and torch.all(out[:, 224:] == 1)
    # fast sine (bf16 path) stays within float32 sinf accuracy
    outf = torch.empty(1000, 192, device="cuda")
    spa.ops.fourier_features(dev(x), outf, 32, 1.0, exact=False)
    assert float((outf.cpu() - om.sinusoidal_embedding(torch.from_numpy(x), 32)).abs().max()) < 1e-6


# ---- GEMMs -----------------------------------------------------------------------------------
def _gemm_ref(a, wt, bias, act, res):
    y = a.double() @ wt.double().t()
    if bias is not None:
        y = y + bias.double()
    if act:
        y = om.gelu_tanh(y)
    if res is not None:
        y = y + res.double()
    return y


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
def test_gemm_shapes(spa, impl):
    ops = spa.ops
    torch.manual_seed(0)
    dtype = torch.float32 if impl == "simt" else torch.bfloat16
    code = ops.GEMM_SIMT if impl == "simt" else ops.GEMM_TCGEN05
    shapes = [(1, 64, 96), (127, 96, 256), (128, 128, 384), (129, 192, 64), (1000, 256, 1280), (333, 600, 1280),
              (2048 + 77, 384, 768), (515, 2304, 384), (300, 1536, 384), (256, 1280, 12352), (700, 1152, 96),
              (150, 88, 200), (64, 512, 1152)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda").to(dtype)
        wt = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(dtype)
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")
        for variant in range(5):
            b = bias if variant in (1, 2, 3) else None
            act = ops.ACT_GELU if variant == 2 else ops.ACT_NONE
            r = res if variant == 3 else (res.to(dtype) if variant == 4 else None)  # 4: residual in the compute dtype (backward)
            odt = torch.float32 if variant in (0, 3) else dtype
            y = ops.gemm(a, wt, b, act, r, out_dtype=odt, impl=code)
            ref = _gemm_ref(a, wt, b, act, r)
            tol = 2e-5 if odt == torch.float32 else 6e-3  # fp32 accumulate of identical operands / bf16 store
            assert rel_err(y, ref) < tol, (impl, M, N, K, variant, rel_err(y, ref))


def test_gemm_auto_dispatch_and_views(spa):
    """bf16 operands with strided views (a column block of a wider matrix) take the tcgen05 path."""
    ops = spa.ops
    torch.manual_seed(1)
    big = torch.randn(500, 3 * 768, device="cuda").to(torch.bfloat16)
    a = big[:, 768:1536]
    wt = (torch.randn(384, 768, device="cuda") / 28).to(torch.bfloat16)
    y = ops.gemm(a, wt, out_dtype=torch.float32)
    assert rel_err(y, a.double() @ wt.double().t()) < 2e-5


def test_gemm_strided_backward_layouts(spa):
    ops = spa.ops
    torch.manual_seed(2)
    M, N, K = 300, 70, 130
    x = torch.randn(M, K, device="cuda")
    dy = torch.randn(M, N, device="cuda")
    # dW^T[N,K] = dY^T X : A(n,m) = dy[m,n] (sam=1, sak=N), B(m,k) = x[m,k]
    dwt = torch.zeros(N, K, device="cuda")
    ops.gemm_strided(dy, 1, N, x, K, 1, dwt, N, K, M)
    assert rel_err(dwt, dy.double().t() @ x.double()) < 1e-5
    ops.gemm_strided(dy, 1, N, x, K, 1, dwt, N, K, M, accumulate=True)
    assert rel_err(dwt, 2 * (dy.double().t() @ x.double())) < 1e-5


# ---- norms -----------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [48, 384, 512, 1024, 1152, 1280, 1536])
def test_layernorm_fwd_bwd(spa, d):
    ops = spa.ops
    torch.manual_seed(3)
    rows = 777
    x = (torch.randn(rows, d, device="cuda") * 2 + 0.5).requires_grad_(True)
    scale = (1 + 0.2 * torch.randn(d, device="cuda")).requires_grad_(True)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), scale.detach(), torch.float32, stats=True)
    ref = om.layer_norm(x.double(), scale.double())
    assert rel_err(y, ref) < 1e-5
    dy = torch.randn(rows, d, device="cuda")
    ref.backward(dy.double())
    dx = torch.empty_like(dy)
    dscale = ops.layernorm_bwd(x.detach(), scale.detach(), mean, rstd, dy, dx)
    assert rel_err(dx, x.grad) < 1e-4 and rel_err(dscale, scale.grad) < 1e-4
    # bf16 dy (what the bf16 backward feeds), accumulation into dx and the bf16 side copy of dx
    dx2 = torch.ones_like(dy)
    low = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dyb = dy.to(torch.bfloat16)
    x.grad = None
    scale.grad = None
    om.layer_norm(x.double(), scale.double()).backward(dyb.double())
    cs = torch.full((d,), 2.0, device="cuda") if ((d % 128 == 0 and d <= 512) or d in (1024, 1280, 1536)) else None   # fused column sums of the final dx (+=)
    dscale2 = ops.layernorm_bwd(x.detach(), scale.detach(), mean, rstd, dyb, dx2, accumulate=True, dx_lowp=low, dx_colsum=cs)
    assert rel_err(dx2, x.grad + 1) < 1e-4 and rel_err(dscale2, scale.grad) < 1e-4
    assert rel_err(low, x.grad + 1) < 6e-3
    if cs is not None:
        assert rel_err(cs, (x.grad + 1).sum(0) + 2.0) < 1e-5, rel_err(cs, (x.grad + 1).sum(0) + 2.0)
    # bf16 output and "first token of each sequence" addressing
    yb = ops.layernorm_fwd(x.detach(), scale.detach(), torch.bfloat16, rows=rows // 7, ldx=7 * d, d=d)
    assert rel_err(yb, ref[::7]) < 6e-3


def test_head_rmsnorm_fwd_bwd(spa):
    ops = spa.ops
    torch.manual_seed(4)
    rows, H, Dh = 500, 8, 96
    buf = torch.randn(rows, 3 * H * Dh, device="cuda")
    scale = (1 + 0.2 * torch.randn(Dh, device="cuda")).requires_grad_(True)
    x = buf[:, : H * Dh].clone().requires_grad_(True)
    ref = om.rms_norm(x.double().view(rows, H, Dh), scale.double()) / math.sqrt(Dh)
    q = buf[:, : H * Dh]
    rstd = ops.head_rmsnorm_fwd(q, scale.detach(), 1 / math.sqrt(Dh), H, Dh, save_rstd=True)
    assert rel_err(q, ref.reshape(rows, -1)) < 1e-5
    dy = torch.randn(rows, H * Dh, device="cuda")
    ref.backward(dy.double().view(rows, H, Dh))
    d_io = dy.clone()
    dscale = ops.head_rmsnorm_bwd(q, scale.detach(), 1 / math.sqrt(Dh), rstd, d_io, H, Dh)
    assert rel_err(d_io, x.grad) < 1e-4 and rel_err(dscale, scale.grad) < 1e-4


@pytest.mark.parametrize("Dh,H,rows", [(96, 8, 1003), (64, 8, 500), (32, 2, 77), (128, 4, 260)])
def test_head_rmsnorm_bwd_bf16_fast_path(spa, Dh, H, rows):
    """bf16 buffers (4 lanes per head, register-resident scale gradient) vs autograd in fp64 on the
    same bf16-rounded normalised output."""
    ops = spa.ops
    torch.manual_seed(40)
    A = H * Dh
    scale = (1 + 0.2 * torch.randn(Dh, device="cuda")).requires_grad_(True)
    x = torch.randn(rows, A, device="cuda").requires_grad_(True)
    mul = 1 / math.sqrt(Dh)
    ref = om.rms_norm(x.double().view(rows, H, Dh), scale.double()) * mul
    buf = torch.zeros(rows, 3 * A, device="cuda", dtype=torch.bfloat16)
    y = buf[:, A : 2 * A]
    y.copy_(ref.reshape(rows, A))
    rstd = torch.rsqrt((x.detach().view(rows, H, Dh) ** 2).mean(-1) + 1e-6).contiguous()
    dy = torch.randn(rows, A, device="cuda").to(torch.bfloat16)
    ref.backward(dy.double().view(rows, H, Dh))
    dbuf = torch.zeros(rows, 3 * A, device="cuda", dtype=torch.bfloat16)
    d_io = dbuf[:, A : 2 * A]
    d_io.copy_(dy)
    dscale = ops.head_rmsnorm_bwd(y, scale.detach(), mul, rstd, d_io, H, Dh)
    assert rel_err(d_io, x.grad) < 1.5e-2 and rel_err(dscale, scale.grad) < 1e-2
    assert float(dbuf[:, :A].abs().max()) == 0 and float(dbuf[:, 2 * A :].abs().max()) == 0


# ---- attention ---------------------------------------------------------------------------------
def _attn_ref(q, k, v, mask, batch, H, Lq, Lk, Dh):
    q4 = q.double().view(batch, Lq, H, Dh)
    k4 = k.double().view(batch, Lk, H, Dh)
    v4 = v.double().view(batch, Lk, H, Dh)
    w = torch.einsum("bqhd,bkhd->bhqk", q4, k4)
    if mask is not None:
        w = torch.where(mask.view(batch, 1, 1, Lk) != 0, w, torch.full_like(w, torch.finfo(torch.float32).min))
    w = torch.softmax(w, dim=-1)
    return torch.einsum("bhqk,bkhd->bqhd", w, v4).reshape(batch * Lq, H * Dh)


@pytest.mark.parametrize("dtype,Lq,Lk,Dh", [
    (torch.float32, 151, 151, 96), (torch.float32, 128, 300, 96), (torch.float32, 13, 13, 64),
    (torch.bfloat16, 151, 151, 96), (torch.bfloat16, 129, 129, 96), (torch.bfloat16, 128, 128, 64),
    (torch.bfloat16, 16, 16, 96), (torch.bfloat16, 37, 37, 64), (torch.bfloat16, 128, 2048, 96),
    (torch.bfloat16, 40, 300, 64), (torch.bfloat16, 128, 4096, 96),
])
def test_attention_fwd(spa, dtype, Lq, Lk, Dh):
    ops = spa.ops
    torch.manual_seed(5)
    batch, H = 5, 8
    A = H * Dh
    qkv = torch.randn(batch * max(Lq, Lk), 3 * A, device="cuda")
    q = (qkv[: batch * Lq, :A] / math.sqrt(Dh)).to(dtype)
    k = qkv[: batch * Lk, A : 2 * A].to(dtype)
    v = qkv[: batch * Lk, 2 * A :].to(dtype)
    mask = None
    if Lq == Lk:
        mask = (torch.rand(batch, Lk, device="cuda") < 0.8).to(torch.uint8)
        mask[:, 0] = 1
        mask[1] = 0  # one sequence entirely masked -> uniform weights (finfo.min semantics)
    o = torch.empty(batch * Lq, A, device="cuda", dtype=dtype)
    ops.attention_fwd(q, k, v, o, batch, H, Lq, Lk, Dh, mask)
    ref = _attn_ref(q, k, v, mask, batch, H, Lq, Lk, Dh)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(o, ref) < tol, rel_err(o, ref)


@pytest.mark.parametrize("dtype,Lq,Lk,Dh", [
    (torch.float32, 40, 40, 96), (torch.float32, 16, 70, 96), (torch.bfloat16, 151, 151, 96), (torch.bfloat16, 129, 129, 96),
    (torch.bfloat16, 128, 128, 64), (torch.bfloat16, 23, 23, 64), (torch.bfloat16, 128, 150, 96), (torch.bfloat16, 64, 300, 96),
    (torch.bfloat16, 128, 2048, 96), (torch.bfloat16, 40, 500, 64),
])
def test_attention_bwd(spa, dtype, Lq, Lk, Dh):
    ops = spa.ops
    torch.manual_seed(6)
    batch, H = 3, 4
    A = H * Dh
    q = (torch.randn(batch * Lq, A, device="cuda") / math.sqrt(Dh)).to(dtype)
    k = torch.randn(batch * Lk, A, device="cuda").to(dtype)
    v = torch.randn(batch * Lk, A, device="cuda").to(dtype)
    mask = None
    if Lq == Lk:
        mask = (torch.rand(batch, Lk, device="cuda") < 0.7).to(torch.uint8)
        mask[:, 0] = 1
        mask[2] = 0  # a fully masked sequence: uniform weights, value gradient only
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qd, kd, vd, mask, batch, H, Lq, Lk, Dh)
    d_o = torch.randn(batch * Lq, A, device="cuda").to(dtype)
    ref.backward(d_o.double())
    o = torch.empty(batch * Lq, A, device="cuda", dtype=dtype)
    stats = ops.attention_fwd(q, k, v, o, batch, H, Lq, Lk, Dh, mask, save_stats=True)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_bwd(q, k, v, o, d_o, dq, dk, dv, stats, batch, H, Lq, Lk, Dh, mask)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for got, want in ((dq, qd.grad), (dk, kd.grad), (dv, vd.grad)):
        assert rel_err(got, want) < tol, rel_err(got, want)


@pytest.mark.parametrize("L,Dh", [(151, 96), (129, 96), (152, 96), (137, 96), (145, 64), (128, 96), (160, 96)])
def test_attention_many_items_per_cta(spa, L, Dh):
    """Several (sequence, head) items per persistent CTA: the cross-item pipelines of the tcgen05 kernels (operand release, lagging
    epilogues, double-buffered accumulators) only show up with more items than SMs."""
    ops = spa.ops
    torch.manual_seed(11)
    batch, H = 160, 4   # 640 items over 148 CTAs
    A = H * Dh
    q = (torch.randn(batch * L, A, device="cuda") / math.sqrt(Dh)).to(torch.bfloat16)
    k = torch.randn(batch * L, A, device="cuda").to(torch.bfloat16)
    v = torch.randn(batch * L, A, device="cuda").to(torch.bfloat16)
    mask = (torch.rand(batch, L, device="cuda") < 0.8).to(torch.uint8)
    mask[:, 0] = 1
    mask[5] = 0
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qd, kd, vd, mask, batch, H, L, L, Dh)
    d_o = torch.randn(batch * L, A, device="cuda").to(torch.bfloat16)
    ref.backward(d_o.double())
    o = torch.empty(batch * L, A, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):   # twice: the second launch must not depend on state the first one left behind
        o.zero_()
        stats = ops.attention_fwd(q, k, v, o, batch, H, L, L, Dh, mask, save_stats=True)
        assert rel_err(o, ref.detach()) < 1e-2, rel_err(o, ref.detach())
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ops.attention_bwd(q, k, v, o, d_o, dq, dk, dv, stats, batch, H, L, L, Dh, mask)
        for got, want in ((dq, qd.grad), (dk, kd.grad), (dv, vd.grad)):
            assert rel_err(got, want) < 2e-2, rel_err(got, want)
    # per-sequence check: an error confined to one item would be averaged away above
    per_seq = ((o.float() - ref.detach().float()).reshape(batch, -1).norm(dim=1) / ref.detach().float().reshape(batch, -1).norm(dim=1)).max().item()
    assert per_seq < 2e-2, per_seq


@pytest.mark.parametrize("Lk,Dh", [(151, 96), (129, 96), (160, 64), (7, 64), (1, 96)])
def test_attention_one_query(spa, Lk, Dh):
    """Lq = 1 (the pruned last layers): dedicated kernel, key mask incl. a fully masked sequence, forward + backward."""
    ops = spa.ops
    torch.manual_seed(16)
    batch, H, dtype = 37, 8, torch.bfloat16
    A = H * Dh
    q = (torch.randn(batch, A, device="cuda") / math.sqrt(Dh)).to(dtype)
    kv = torch.randn(batch * Lk, 2 * A, device="cuda").to(dtype)     # k and v as column slices of one buffer (ld = 2A)
    k, v = kv[:, :A], kv[:, A:]
    mask = (torch.rand(batch, Lk, device="cuda") < 0.7).to(torch.uint8)
    mask[:, 0] = 1
    mask[3] = 0
    qd, kd, vd = (t.double().clone().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qd, kd, vd, mask, batch, H, 1, Lk, Dh)
    d_o = torch.randn(batch, A, device="cuda").to(dtype)
    ref.backward(d_o.double())
    o = torch.empty(batch, A, device="cuda", dtype=dtype)
    stats = ops.attention_fwd(q, k, v, o, batch, H, 1, Lk, Dh, mask, save_stats=True)
    assert rel_err(o, ref) < 1e-2, rel_err(o, ref)
    dq = torch.empty_like(q)
    dkv = torch.full_like(kv, float("nan"))
    ops.attention_bwd(q, k, v, o, d_o, dq, dkv[:, :A], dkv[:, A:], stats, batch, H, 1, Lk, Dh, mask)
    for got, want in ((dq, qd.grad), (dkv[:, :A], kd.grad), (dkv[:, A:], vd.grad)):
        assert rel_err(got, want) < 2e-2, rel_err(got, want)


# ---- small kernels --------------------------------------------------------------------------------
def test_key_mask_matches_oracle(spa):
    rs = np.random.RandomState(7)
    B, N, T = 3, 5, 11
    vis = (rs.uniform(size=(B, N, T, 1)) < 0.7).astype(np.float32)
    bd = np.array([T, T - 4, 0], np.int32)
    m = spa.ops.build_key_mask(dev(vis), dev(bd), True)
    ref = om.key_mask_3d(torch.from_numpy(vis), torch.from_numpy(bd))
    assert torch.equal(m.cpu().bool(), ref)
    m2 = spa.ops.build_key_mask(dev(vis), dev(bd), False)
    assert torch.equal(m2.cpu().bool(), ref[..., 1:])


def test_quantize_bit_exact(spa):
    rs = np.random.RandomState(8)
    x = (rs.standard_normal((4, 128, 96)) * 0.8).astype(np.float32)
    x.reshape(-1)[:6] = [0.5 / 128, 1.5 / 128, 2.5 / 128, -0.5 / 128, 1.0, -1.0]  # ties + clip edges
    noise = rs.uniform(size=x.shape).astype(np.float32)
    y, mask = spa.ops.quantize_fwd(dev(x), dev(noise), True, save_mask=True)
    ref = om.quantize_latents(torch.from_numpy(x), torch.from_numpy(noise), True)
    assert torch.equal(y.cpu(), ref)
    assert torch.equal(mask.cpu().bool(), torch.from_numpy((x >= -1) & (x <= 1)))
    y2 = spa.ops.quantize_fwd(dev(x), None, False)
    assert torch.equal(y2.cpu(), torch.clamp(torch.from_numpy(x), -1, 1))


def test_decoder_tokens_fwd_bwd(spa):
    ops = spa.ops
    torch.manual_seed(9)
    B, Q, L, C = 2, 5, 7, 200
    D = C + 128
    lat = torch.randn(B, L, C, device="cuda")
    qe = torch.randn(B * Q, D, device="cuda")
    qf = torch.tensor([[0, 3, 14, 20, 39], [1, 2, 40, 0, 7]], dtype=torch.int32, device="cuda")  # 5*40+127 > C: zero fill
    tok = torch.empty(B * Q * (L + 1), D, device="cuda")
    ops.decoder_tokens_fwd(lat, qe, qf, tok, B, Q, L, C)
    latq = lat[:, None].expand(B, Q, L, C).cpu()
    ref = torch.cat([qe.view(B, Q, 1, D).cpu(), om.append_time_feat(latq, qf.cpu())], dim=2)
    assert torch.equal(tok.view(B, Q, L + 1, D).cpu(), ref)
    g = torch.randn_like(tok)
    d_lat = torch.empty(B, L, C, device="cuda")
    d_qe = torch.empty(B * Q, D, device="cuda")
    ops.decoder_tokens_bwd(g, qf, d_lat, d_qe, B, Q, L, C)
    lat_r = lat.cpu().double().requires_grad_(True)
    qe_r = qe.cpu().double().requires_grad_(True)
    ref2 = torch.cat([qe_r.view(B, Q, 1, D), om.append_time_feat(lat_r[:, None].expand(B, Q, L, C), qf.cpu())], dim=2)
    ref2.backward(g.view(B, Q, L + 1, D).cpu().double())
    assert rel_err(d_lat, lat_r.grad) < 1e-5 and rel_err(d_qe, qe_r.grad) < 1e-6
    # many queries: the reduction over queries is split across blocks (atomic partial sums into a zeroed buffer)
    Q2 = 83
    qf2 = torch.randint(0, 41, (B, Q2), dtype=torch.int32, device="cuda")
    g2 = torch.randn(B * Q2 * (L + 1), D, device="cuda")
    d_lat2 = torch.full((B, L, C), float("nan"), device="cuda")
    d_qe2 = torch.empty(B * Q2, D, device="cuda")
    ops.decoder_tokens_bwd(g2, qf2, d_lat2, d_qe2, B, Q2, L, C)
    lat_r = lat.cpu().double().requires_grad_(True)
    qe2_r = torch.zeros(B * Q2, D, dtype=torch.float64, requires_grad=True)
    ref3 = torch.cat([qe2_r.view(B, Q2, 1, D), om.append_time_feat(lat_r[:, None].expand(B, Q2, L, C), qf2.cpu())], dim=2)
    ref3.backward(g2.view(B, Q2, L + 1, D).cpu().double())
    assert rel_err(d_lat2, lat_r.grad) < 1e-5 and rel_err(d_qe2, qe2_r.grad) < 1e-6


def test_split_and_loss(spa):
    ops = spa.ops
    torch.manual_seed(10)
    rows, T = 37, 12
    ho = torch.randn(rows, 4 * T, device="cuda")
    tracks, vis, cert = ops.split_outputs(ho, T, 3)
    h = ho.cpu()
    assert torch.equal(tracks.cpu(), torch.stack([h[:, :T], h[:, T : 2 * T], h[:, 2 * T : 3 * T]], -1))
    assert torch.equal(vis.cpu()[..., 0], h[:, 3 * T :]) and not cert.any()
    tt = torch.randn(rows, T, 3, device="cuda")
    tv = (torch.rand(rows, T, 1, device="cuda") < 0.6).float()
    sums = torch.zeros(3, device="cuda")
    ops.loss_fwd(ho, tt, tv, sums, T)
    hod = h.double().requires_grad_(True)
    pred = om.Results(torch.stack([hod[:, :T], hod[:, T : 2 * T], hod[:, 2 * T : 3 * T]], -1)[None],
                      hod[:, 3 * T :, None][None], None)
    ref = om.compute_loss_3d(pred, {"query_tracks": tt.cpu().double()[None], "query_tracks_visible": tv.cpu().double()[None]})
    nv = float(sums[2])
    assert nv == float(tv.sum())
    assert abs(float(sums[0]) / max(nv, 1) - float(ref["position_loss"])) < 1e-5 * float(ref["position_loss"])
    assert abs(float(sums[1]) / max(nv, 1) - float(ref["visible_loss"])) < 1e-5 * float(ref["visible_loss"])
    ref["total_loss"].backward()
    d = ops.loss_bwd(ho, tt, tv, 5000.0, 1e-8, 1.0 / max(nv, 1), T)
    assert rel_err(d, hod.grad) < 1e-5


def test_gelu_colsum_adamw(spa):
    ops = spa.ops
    torch.manual_seed(11)
    pre = torch.randn(300, 200, device="cuda").requires_grad_(True)
    dy = torch.randn(300, 200, device="cuda")
    om.gelu_tanh(pre.double()).backward(dy.double())
    dx = torch.empty_like(dy)
    ops.gelu_bwd(pre.detach(), dy, dx)
    assert rel_err(dx, pre.grad) < 1e-5
    y = ops.gelu_fwd(pre.detach(), torch.empty_like(dy))
    assert rel_err(y, om.gelu_tanh(pre.detach().double())) < 1e-6
    out = torch.empty(200, device="cuda")
    ops.colsum(dy, out)
    assert rel_err(out, dy.double().sum(0)) < 1e-5
    # AdamW + global-norm clip vs the oracle restatement of optax (train.py:239-243)
    p = torch.randn(5000, device="cuda")
    g = torch.randn(5000, device="cuda") * 0.05
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pr, mr, vr = [p.cpu().clone()], [torch.zeros(5000)], [torch.zeros(5000)]
    for step in (1, 2, 3):
        ss = torch.zeros(1, device="cuda")
        ops.sumsq(g, ss)
        ops.adamw_step(p, g, m, v, ss, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
        om.adamw_step(pr, [g.cpu().clone()], mr, vr, step, 1e-3)
    assert rel_err(p, pr[0]) < 1e-5


def test_masked_mean(spa):
    torch.manual_seed(12)
    S, T, W = 9, 14, 50
    tok = torch.randn(S * T, W, device="cuda")
    vis = (torch.rand(S, T, device="cuda") < 0.5).float()
    vis[2] = 0
    out = spa.ops.masked_mean_fwd(tok, vis, S, T, torch.float32)
    ref = (tok.view(S, T, W) * vis[..., None]).sum(1) / torch.clamp(vis.sum(1, keepdim=True), min=1.0)
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("impl,Dh", [("tcgen05", 96), ("tcgen05", 64), ("tcgen05", 32), ("simt", 96)])
def test_gemm_fused_head_rmsnorm(spa, impl, Dh):
    """QKV projection with the per-head RMSNorm in the GEMM epilogue (attention.py:154-173)."""
    ops = spa.ops
    torch.manual_seed(13)
    H = 8 if Dh != 32 else 2
    A, d, M = H * Dh, 384, 1000 + 77
    dtype = torch.bfloat16 if impl == "tcgen05" else torch.float32
    code = ops.GEMM_TCGEN05 if impl == "tcgen05" else ops.GEMM_SIMT
    x = torch.randn(M, d, device="cuda").to(dtype)
    wt = (torch.randn(3 * A, d, device="cuda") / math.sqrt(d)).to(dtype)
    sq = 1 + 0.2 * torch.randn(Dh, device="cuda")
    sk = 1 + 0.2 * torch.randn(Dh, device="cuda")
    raw = x.double() @ wt.double().t()
    q = om.rms_norm(raw[:, :A].view(M, H, Dh), sq.double()).reshape(M, A) / math.sqrt(Dh)
    k = om.rms_norm(raw[:, A : 2 * A].view(M, H, Dh), sk.double()).reshape(M, A)
    ref = torch.cat([q, k, raw[:, 2 * A :]], dim=1)
    tol = 2e-5 if impl == "simt" else 6e-3
    out, rstd = ops.gemm_rmsnorm(x, wt, Dh, A, A, sq, sk, save_rstd=True, impl=code)
    assert rel_err(out[:, :A], ref[:, :A]) < tol and rel_err(out[:, A : 2 * A], ref[:, A : 2 * A]) < tol
    assert rel_err(out[:, 2 * A :], ref[:, 2 * A :]) < tol
    rref = torch.rsqrt((raw[:, : 2 * A].view(M, 2 * H, Dh) ** 2).mean(-1) + 1e-6)
    assert rel_err(rstd, rref) < 1e-4
    # cross-attention splits: queries only / keys+values only, on a strided "token 0" view of A
    kv = ops.gemm_rmsnorm(x, wt[A:], Dh, 0, A, sq, sk, impl=code)
    assert rel_err(kv[:, :A], ref[:, A : 2 * A]) < tol and rel_err(kv[:, A:], ref[:, 2 * A :]) < tol
    rows = M // 7
    q0 = ops.gemm_rmsnorm(x[: rows * 7].view(rows, 7 * d)[:, :d], wt[:A], Dh, A, 0, sq, sk, impl=code)
    assert rel_err(q0, ref[: rows * 7 : 7, :A]) < tol


def test_gemm_persistent_tile_counts(spa):
    """Persistent tile loop: fewer / more tiles than SMs, odd M tails."""
    ops = spa.ops
    torch.manual_seed(14)
    for M in (129, 256, 257, 128 * 7 + 5, 128 * 300 + 1):
        a = torch.randn(M, 384, device="cuda").to(torch.bfloat16)
        wt = (torch.randn(768, 384, device="cuda") / 20).to(torch.bfloat16)
        y = ops.gemm(a, wt, out_dtype=torch.float32, impl=ops.GEMM_TCGEN05)
        assert rel_err(y, a.double() @ wt.double().t()) < 2e-5, M


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_gemm_dw_weight_gradient(spa, impl):
    """dW[N,K] += dY[M,N]^T X[M,K] (backward of every Dense): token-major operands on tcgen05 with
    split-K atomics vs fp64 of the same bf16 operands; ragged M / N / K, strided views, accumulation."""
    ops = spa.ops
    torch.manual_seed(21)
    code = ops.GEMM_TCGEN05 if impl == "tcgen05" else ops.GEMM_SIMT
    dtype = torch.bfloat16 if impl == "tcgen05" else torch.float32
    shapes = [(300, 72, 136), (5000, 384, 1280), (1000, 600, 1280), (777, 2304, 384), (4096 + 13, 96, 512),
              (2000, 1152, 96), (64, 128, 64), (20000, 1536, 384), (129, 8, 8)]
    for (M, N, K) in shapes:
        big = torch.randn(M, N + 16, device="cuda").to(dtype)
        dy = big[:, 8 : 8 + N]                      # strided view (a column block of a wider buffer)
        x = torch.randn(M, K, device="cuda").to(dtype)
        ref = dy.double().t() @ x.double()
        dw = torch.full((N, K), 3.0, device="cuda")
        ops.gemm_dw(dy, x, dw, accumulate=False, impl=code)
        assert rel_err(dw, ref) < 5e-5, (impl, M, N, K, rel_err(dw, ref))
        ops.gemm_dw(dy, x, dw, accumulate=True, impl=code)
        assert rel_err(dw, 2 * ref) < 5e-5, (impl, M, N, K, "accumulate")


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_gemm_gelu_fused_forward_and_backward(spa, impl):
    """MLP_in + tanh-GELU with the pre-activation as a side output, and the backward through MLP_out
    and the activation in one epilogue (attention.py:103-108 under autodiff)."""
    ops = spa.ops
    torch.manual_seed(31)
    dtype = torch.bfloat16 if impl == "tcgen05" else torch.float32
    code = ops.GEMM_TCGEN05 if impl == "tcgen05" else ops.GEMM_SIMT
    for (M, d, Mh) in [(1000 + 13, 384, 1536), (300, 1280, 1536), (129, 64, 200)]:
        a = torch.randn(M, d, device="cuda").to(dtype)
        w1t = (torch.randn(Mh, d, device="cuda") / math.sqrt(d)).to(dtype)
        b1 = torch.randn(Mh, device="cuda")
        z, h = ops.gemm_gelu(a, w1t, b1, impl=code)
        zref = a.double() @ w1t.double().t() + b1.double()
        href = torch.nn.functional.gelu(zref, approximate="tanh")
        tol = 6e-3 if impl == "tcgen05" else 2e-5
        assert rel_err(z, zref) < tol and rel_err(h, href) < tol, (M, d, Mh)
        dy = torch.randn(M, d, device="cuda").to(dtype)
        w2 = (torch.randn(Mh, d, device="cuda") / math.sqrt(Mh)).to(dtype)   # MLP_out kernel [Mh, d] as Flax stores it
        dz = ops.gemm_gelu_bwd(dy, w2, z, impl=code)
        zz = z.double().requires_grad_(True)
        torch.nn.functional.gelu(zz, approximate="tanh").backward(dy.double() @ w2.double().t())
        assert rel_err(dz, zz.grad) < (1e-2 if impl == "tcgen05" else 2e-5), (M, d, Mh, rel_err(dz, zz.grad))
        if impl == "tcgen05":
            # saved-derivative form: the forward writes gelu'(z) (sharing tanh(u) with the activation), the backward multiplies
            g, h2 = ops.gemm_gelu(a, w1t, b1, impl=code, save_grad=True)
            assert torch.equal(h2, h)
            z64 = zref.clone().requires_grad_(True)
            torch.nn.functional.gelu(z64, approximate="tanh").sum().backward()
            assert rel_err(g, z64.grad) < 6e-3, rel_err(g, z64.grad)
            bsum = torch.full((Mh,), 0.5, device="cuda")
            dz2 = ops.gemm_gelu_bwd(dy, w2, g, impl=code, z_is_grad=True, dz_colsum=bsum)
            want = z64.grad * (dy.double() @ w2.double().t())
            assert rel_err(dz2, want) < 1e-2
            assert rel_err(bsum, want.sum(0) + 0.5) < 1e-2, rel_err(bsum, want.sum(0) + 0.5)   # bias gradient from the epilogue (+=)


@pytest.mark.parametrize("N,T,Dd,Dz,W", [(40, 7, 768, 256, 384), (3, 150, 768, 256, 384), (33, 12, 64, 0, 256), (9, 5, 0, 128, 192)])
def test_embed_fused_matches_unfused_math(spa, N, T, Dd, Dz, W):
    """K1 fused embedding (Fourier + fp32->bf16 features + first projection in one tcgen05 kernel)
    vs the oracle's SinusoidalEmbedding + the stacked Dense in fp64 on bf16-rounded operands
    (track_autoencoder_3d.py:123-149); read-out slots must be left untouched."""
    ops = spa.ops
    torch.manual_seed(50)
    R = N * T
    tracks = (torch.rand(R, 3, device="cuda") * 2 - 1)
    dino = torch.randn(R, Dd, device="cuda") if Dd else None
    depth = torch.randn(R, Dz, device="cuda") if Dz else None
    K = 256 + Dd + Dz
    wt = (torch.randn(W, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(W, device="cuda")
    assert ops.embed_fused_applicable(W, K, Dd, Dz, 3)
    out = torch.full((N * (T + 1), W), -7.0, device="cuda")
    a_cat = torch.full((N * (T + 1), K), -3.0, device="cuda", dtype=torch.bfloat16)
    ops.embed_fused(tracks, dino, depth, wt, bias, out, T, 32, 1.0, a_cat=a_cat)
    t = (torch.arange(R, device="cuda") % T).float() / T
    feats = om.sinusoidal_embedding(torch.cat([tracks, t[:, None]], -1).cpu(), 32).cuda()
    parts = [feats.float().to(torch.bfloat16).double()]
    if Dd:
        parts.append(dino.to(torch.bfloat16).double())
    if Dz:
        parts.append(depth.to(torch.bfloat16).double())
    ref = torch.cat(parts, -1) @ wt.double().t() + bias.double()
    got = out.view(N, T + 1, W)
    assert float((got[:, 0] + 7.0).abs().max()) == 0.0          # read-out slots untouched
    assert rel_err(got[:, 1:].reshape(R, W), ref) < 3e-3, rel_err(got[:, 1:].reshape(R, W), ref)
    # side output for training: the bf16 concatenated features at the same (remapped) rows
    ac = a_cat.view(N, T + 1, K)
    assert float((ac[:, 0].float() + 3.0).abs().max()) == 0.0
    cat = torch.cat(parts, -1)
    assert rel_err(ac[:, 1:, :256].reshape(R, 256), cat[:, :256]) < 1e-2          # sin.approx + bf16 rounding
    assert torch.equal(ac[:, 1:, 256:].reshape(R, K - 256).double(), cat[:, 256:])   # features: the same bf16 rounding


@pytest.mark.parametrize("Lq,Lk,Dh,batch", [(128, 2048, 96, 2), (128, 4096, 96, 1), (128, 24, 96, 3), (100, 130, 96, 2), (37, 300, 64, 2), (128, 129, 64, 1)])
def test_cross_attention_tcgen05_fwd_bwd(spa, Lq, Lk, Dh, batch):
    """Latents<-tracks cross-attention (track_autoencoder_3d.py:200-201) on tcgen05: key chunks split across CTAs, flash-style
    merge of the chunks' (max, sum, O) states, dQ summed over the chunks; with and without a key mask (a fully masked sequence
    is uniform, finfo.min semantics), ragged Lq / Lk; the dispatch counter shows the tensor-core kernel ran."""
    ops = spa.ops
    torch.manual_seed(31)
    H = 4
    A = H * Dh
    for masked in (False, True):
        q = (torch.randn(batch * Lq, A, device="cuda") / math.sqrt(Dh)).to(torch.bfloat16)
        k = torch.randn(batch * Lk, A, device="cuda").to(torch.bfloat16)
        v = torch.randn(batch * Lk, A, device="cuda").to(torch.bfloat16)
        mask = None
        if masked:
            mask = (torch.rand(batch, Lk, device="cuda") < 0.7).to(torch.uint8)
            mask[:, 0] = 1
            mask[batch - 1] = 0
        qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
        ref = _attn_ref(qd, kd, vd, mask, batch, H, Lq, Lk, Dh)
        d_o = torch.randn(batch * Lq, A, device="cuda").to(torch.bfloat16)
        ref.backward(d_o.double())
        ops.stats(reset=True)
        o = torch.empty(batch * Lq, A, device="cuda", dtype=torch.bfloat16)
        stats = ops.attention_fwd(q, k, v, o, batch, H, Lq, Lk, Dh, mask, save_stats=True)
        assert rel_err(o, ref) < 1e-2, rel_err(o, ref)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ops.attention_bwd(q, k, v, o, d_o, dq, dk, dv, stats, batch, H, Lq, Lk, Dh, mask)
        st = ops.stats()
        assert st["attention_cross_tcgen05"] == 2 and st["attention_simt"] == 0, st
        for got, want in ((dq, qd.grad), (dk, kd.grad), (dv, vd.grad)):
            assert rel_err(got, want) < 2e-2, (masked, rel_err(got, want))


def test_gemm_bf16x3_accurate_mode(spa):
    """"bf16 x 3": fp32 operands split into three bf16 terms, six tcgen05 products per K block into one fp32 accumulator - the
    tensor-core form of the accurate mode.  Against float64 of the SAME fp32 operands: ~1e-6, i.e. fp32-GEMM quality (the
    north_star's 1e-4 path), with bias, residual, GELU, strided rows, ragged M / N."""
    ops = spa.ops
    torch.manual_seed(41)
    for (M, K, N) in [(300, 384, 2304), (1000, 1280, 600), (129, 64, 8), (5000, 1536, 384), (77, 12352, 1280), (256, 512, 96)]:
        big = torch.randn(M, K + 8, device="cuda")
        a = big[:, 4 : 4 + K]                      # row pitch K + 8, 16-byte aligned start
        w = torch.randn(N, K, device="cuda") / math.sqrt(K)
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")
        w3 = ops.split3(w)
        parts = w3.float().view(N, 3, K)
        assert torch.equal(parts.sum(1), w) or float((parts.sum(1) - w).abs().max()) < 2e-7 * float(w.abs().max())
        ops.stats(reset=True)
        y = ops.gemm(a, w3, bias, residual=res, out_dtype=torch.float32)
        assert ops.stats()["gemm_x3"] == 1
        ref = a.double() @ w.double().t() + bias.double() + res.double()
        tol = 3e-6 * max(1.0, K / 2048)     # fp32 accumulation over 6 K / 64 tensor-core K blocks (K = 12352: the query encoder)
        assert rel_err(y, ref) < tol, (M, K, N, rel_err(y, ref))
        h = ops.gemm(a, w3, bias, act=ops.ACT_GELU)
        z = a.double() @ w.double().t() + bias.double()
        gref = 0.5 * z * (1 + torch.tanh(0.7978845608028654 * (z + 0.044715 * z ** 3)))
        assert rel_err(h, gref) < tol, (M, K, N, "gelu", rel_err(h, gref))
    # fused-form projection + per-head RMSNorm on split weights
    M, K, H, Dh = 200, 384, 8, 96
    A = H * Dh
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(3 * A, K, device="cuda") / math.sqrt(K)
    sq, sk = torch.rand(Dh, device="cuda") + 0.5, torch.rand(Dh, device="cuda") + 0.5
    out, rstd = ops.gemm_rmsnorm(a, ops.split3(w), Dh, A, A, sq, sk, save_rstd=True)
    z = (a.double() @ w.double().t()).view(M, 3, H, Dh)
    rms = lambda t, s: t * torch.rsqrt((t * t).mean(-1, keepdim=True) + 1e-6) * s.double()
    ref = torch.stack([rms(z[:, 0], sq) / math.sqrt(Dh), rms(z[:, 1], sk), z[:, 2]], 1).reshape(M, 3 * A)
    assert rel_err(out, ref) < 3e-6 and rstd.shape == (M, 2 * H)


@pytest.mark.parametrize("M,Hd", [(128, 128), (1000, 256), (4531, 1536), (128 * 150 + 77, 1536)])
def test_mlp_fused_matches_unfused_and_fp64(spa, M, Hd):
    """out = residual + gelu_tanh(a W1 + b1) W2 + b2 in one kernel (hidden activation never in HBM) against float64 of the same bf16
    operands (with the hidden activation rounded to bf16, as both CUDA paths do) and against the two-GEMM path; ragged M."""
    ops = spa.ops
    torch.manual_seed(51)
    D = 384
    a = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(Hd, D, device="cuda") / math.sqrt(D)).to(torch.bfloat16)
    w2 = (torch.randn(D, Hd, device="cuda") / math.sqrt(Hd)).to(torch.bfloat16)
    b1, b2 = torch.randn(Hd, device="cuda") * 0.3, torch.randn(D, device="cuda") * 0.3
    res = torch.randn(M, D, device="cuda")
    guard = torch.full((M + 2, D), 7.0, device="cuda")
    out = guard[1 : M + 1]
    ops._torch_ops._raw["mlp_fused"](a, w1, b1, w2, b2, res, out)
    assert float(guard[0].min()) == 7.0 and float(guard[M + 1].min()) == 7.0        # neighbours untouched
    z = a.double() @ w1.double().t() + b1.double()
    h = (0.5 * z * (1 + torch.tanh(0.7978845608028654 * (z + 0.044715 * z ** 3)))).to(torch.bfloat16).double()
    ref = h @ w2.double().t() + b2.double() + res.double()
    assert rel_err(out, ref) < 4e-3, rel_err(out, ref)
    hh = ops.gemm(a, w1, b1, act=ops.ACT_GELU)
    two = ops.gemm(hh, w2, b2, residual=res, out_dtype=torch.float32)
    assert rel_err(out, two) < 4e-3, rel_err(out, two)
    y = ops.mlp_fused(a, w1, b1, w2, b2, res)                                            # through the dispatcher
    assert torch.equal(y, out)
