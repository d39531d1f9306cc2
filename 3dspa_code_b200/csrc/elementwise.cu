// Bandwidth-bound kernels of the 3DSPA hot path: Fourier features, norms, masks, quantiser,
// decoder token assembly, output split, loss, optimiser.  Each function cites the reference
// lines it replaces in include/spa3d_b200.h.
#include <cmath>

#include "common.cuh"

namespace spa3d {

// ------------------------------------------------------------------------------------------
// SinusoidalEmbedding (track_autoencoder.py:18-38)
// ------------------------------------------------------------------------------------------
struct FourierScales {
  float s[64];
};

template <typename TO, bool EXACT>
__global__ void fourier_kernel(const float* __restrict__ x, int64_t ldx, TO* __restrict__ out,
                               int64_t ldo, int64_t rows, int C, int Ctot, int F, float scale_factor,
                               int append_time, int out_group, FourierScales sc) {
  // one thread per (row, coord, freq): writes the sin and the cos(=sin(.+pi/2)) feature
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = rows * Ctot * F;
  if (idx >= total) return;
  int f = (int)(idx % F);
  int c = (int)((idx / F) % Ctot);
  int64_t r = idx / ((int64_t)F * Ctot);
  float v;
  if (c < C) {
    v = x[r * ldx + c];
  } else if (append_time > 0 && c == C) {
    // fr_id = arange(T)/T in float32 (track_autoencoder_3d.py:126)
    v = __fdiv_rn((float)(r % append_time), (float)append_time);
  } else {
    v = 0.f;  // query_frame // 150.0 == 0 (track_autoencoder_3d.py:268-269)
  }
  v = __fdiv_rn(v, scale_factor);
  float a0 = __fmul_rn(v, sc.s[f]);
  float a1 = __fadd_rn(a0, 1.57079632679489661923f);  // float32(0.5*pi)
  float s0, s1;
  if (EXACT) {
    s0 = (float)sin((double)a0);
    s1 = (float)sin((double)a1);
  } else {
    s0 = sinf(a0);
    s1 = sinf(a1);
  }
  int64_t ro = out_group > 0 ? r + r / out_group + 1 : r;
  TO* o = out + ro * ldo + (int64_t)c * 2 * F;
  stf<TO>(o + f, s0);
  stf<TO>(o + F + f, s1);
}

template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst,
                               int64_t ldd, int64_t rows, int cols, int out_group) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols;
  int c = (int)(idx % cols);
  int64_t ro = out_group > 0 ? r + r / out_group + 1 : r;
  stf<TD>(dst + ro * ldd + c, ldf<TS>(src + r * lds + c));
}

// four elements per thread, 128-bit accesses on the fp32 side (cols % 4 == 0, 16 / 8-byte aligned rows)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) convert_vec4_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd,
                                                           int64_t rows, int cols4, int out_group) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols4) return;
  const int64_t r = idx / cols4;
  const int c = (int)(idx % cols4) * 4;
  const int64_t ro = out_group > 0 ? r + r / out_group + 1 : r;
  float4 v;
  if constexpr (sizeof(TS) == 4) {
    v = __ldcs(reinterpret_cast<const float4*>(src + r * lds + c));
  } else {
    const uint2 u = __ldcs(reinterpret_cast<const uint2*>(src + r * lds + c));
    v = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  }
  if constexpr (sizeof(TD) == 4) {
    *reinterpret_cast<float4*>(dst + ro * ldd + c) = v;
  } else {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + ro * ldd + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  }
}

template <typename TD>
__global__ void set_rows_kernel(TD* __restrict__ dst, int64_t ld, int64_t row_stride,
                                const float* __restrict__ vec, int64_t rows, int cols) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols;
  int c = (int)(idx % cols);
  stf<TD>(dst + r * row_stride * ld + c, vec[c]);
}

// ------------------------------------------------------------------------------------------
// LayerNorm (scale only, eps 1e-6, fast variance) - one warp per row
// ------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void layernorm_fwd_kernel(const TX* __restrict__ x, int64_t ldx,
                                     const float* __restrict__ scale, TY* __restrict__ y,
                                     int64_t ldy, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, int64_t rows, int d) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const TX* xr = x + row * ldx;
  float s = 0.f, ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    float v = ldf<TX>(xr + i);
    s += v;
    ss += v * v;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float mean = s / d;
  float var = fmaxf(ss / d - mean * mean, 0.f);
  float rstd = rsqrtf(var + kNormEps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  TY* yr = y + row * ldy;
  for (int i = lane; i < d; i += 32) {
    float v = ldf<TX>(xr + i);
    stf<TY>(yr + i, (v - mean) * rstd * scale[i]);
  }
}

// Bandwidth-shaped forward: rows of d = 128*J values (fp32, or bf16 when the residual stream is kept in bf16), one warp per row, the
// whole row in registers (J float4 per lane, read once with 128 / 64-bit loads), bf16 or fp32 output with 8 / 16-byte stores.
template <typename TX>
__device__ __forceinline__ float4 ln_load4(const TX* p);
template <>
__device__ __forceinline__ float4 ln_load4<float>(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 ln_load4<bf16>(const bf16* p) {
  const uint2 h = __ldcs(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16), __uint_as_float(h.y & 0xffff0000u));
}

template <int J, typename TY, typename TX = float>
__global__ void __launch_bounds__(256) layernorm_fwd_fast_kernel(const TX* __restrict__ x, int64_t ldx,
                                                                 const float* __restrict__ scale, TY* __restrict__ y,
                                                                 int64_t ldy, float* __restrict__ mean_out,
                                                                 float* __restrict__ rstd_out, int64_t rows) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int d = 128 * J;
  const TX* xr = x + row * ldx;
  float4 v[J];
#pragma unroll
  for (int j = 0; j < J; ++j) v[j] = ln_load4<TX>(xr + (j * 32 + lane) * 4);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float mean = s * (1.f / d);
  const float var = fmaxf(ss * (1.f / d) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + kNormEps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* sc = reinterpret_cast<const float4*>(scale);
  TY* yr = y + row * ldy;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const float4 g = __ldg(sc + j * 32 + lane);
    const float o0 = (v[j].x - mean) * rstd * g.x, o1 = (v[j].y - mean) * rstd * g.y;
    const float o2 = (v[j].z - mean) * rstd * g.z, o3 = (v[j].w - mean) * rstd * g.w;
    if constexpr (sizeof(TY) == 2) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
      *reinterpret_cast<uint2*>(yr + (j * 32 + lane) * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    } else {
      *reinterpret_cast<float4*>(yr + (j * 32 + lane) * 4) = make_float4(o0, o1, o2, o3);
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*scale;  dscale += dy*xhat
template <typename TX, typename TDY, typename TDX>
__global__ void layernorm_bwd_kernel(const TX* __restrict__ x, int64_t ldx,
                                     const float* __restrict__ scale,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const TDY* __restrict__ dy, int64_t lddy, TDX* __restrict__ dx,
                                     int64_t lddx, int dx_accumulate,
                                     float* __restrict__ dscale_partial, int ds_accum, int64_t rows, int d) {
  // grid-stride over rows, one warp per row; per-block partial dscale accumulated in smem
  extern __shared__ float s_ds[];  // [d]
  for (int i = threadIdx.x; i < d; i += blockDim.x) s_ds[i] = 0.f;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    const TX* xr = x + row * ldx;
    const TDY* dyr = dy + row * lddy;
    float mu = mean[row], rs = rstd[row];
    float sg = 0.f, sgx = 0.f;
    for (int i = lane; i < d; i += 32) {
      float xh = (ldf<TX>(xr + i) - mu) * rs;
      float dyv = ldf<TDY>(dyr + i);
      float g = dyv * scale[i];
      sg += g;
      sgx += g * xh;
      atomicAdd(&s_ds[i], dyv * xh);
    }
    sg = warp_sum(sg) / d;
    sgx = warp_sum(sgx) / d;
    TDX* dxr = dx + row * lddx;
    for (int i = lane; i < d; i += 32) {
      float xh = (ldf<TX>(xr + i) - mu) * rs;
      float g = ldf<TDY>(dyr + i) * scale[i];
      float v = rs * (g - sg - xh * sgx);
      if (dx_accumulate) v += ldf<TDX>(dxr + i);
      stf<TDX>(dxr + i, v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    if (ds_accum) atomicAdd(&dscale_partial[i], s_ds[i]);   // straight into the gradient of the scale
    else dscale_partial[(int64_t)blockIdx.x * d + i] = s_ds[i];
  }
}


// Bandwidth version for fp32 x / dx with d = 128*J: one warp per row, lane owns the float4 chunks
// lane + 32j (128-bit loads), the row stays in registers between the statistics and the output
// pass, the scale gradient is accumulated in per-lane registers across the grid-stride loop.
// dx_lowp (optional) receives a bf16 copy of the final dx - the operand of the next backward GEMMs.
template <int J, typename TDY>
__global__ void __launch_bounds__(256, (J >= 8 ? 2 : 1))
layernorm_bwd_fast_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ scale,
                          const float* __restrict__ mean, const float* __restrict__ rstd,
                          const TDY* __restrict__ dy, int64_t lddy, float* __restrict__ dx, int64_t lddx,
                          int dx_accumulate, bf16* __restrict__ dx_lowp, int64_t ldl,
                          float* __restrict__ dscale_partial, int ds_accum, float* __restrict__ dcol, int64_t rows) {
  constexpr int D = 128 * J;
  constexpr bool DYF = sizeof(TDY) == 4;
  constexpr bool CS = J <= 4;   // column sums of the final dx (the bias gradient of the layer that produced x): narrow rows only
  __shared__ float s_ds[D];
  __shared__ float s_cs[CS ? D : 1];
  __shared__ __align__(16) float s_sc[D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    s_ds[i] = 0.f;
    if constexpr (CS) s_cs[i] = 0.f;
    s_sc[i] = scale[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float4 acc[J];
  float4 csum[CS ? J : 1];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < (CS ? J : 1); ++j) csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto unpack = [](const uint4& w) -> float4 {   // dy chunk as loaded (fp32: 4 floats; bf16: 4 values in .x,.y)
    if constexpr (DYF) {
      return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
    } else {
      const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&w.x);
      const __nv_bfloat162 h1 = *reinterpret_cast<const __nv_bfloat162*>(&w.y);
      return make_float4(__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1));
    }
  };
  // Narrow rows (J <= 4: a row is 12-16 floats per lane) are processed two at a time and the old dx of an accumulating call
  // is loaded together with x and dy: every warp then keeps two rows x three streams in flight instead of paying one HBM
  // round trip for x / dy and a second, dependent one for dx per row.  Wide rows (J >= 8) stay at one row (register budget).
  constexpr int R = CS ? 2 : 1;
  constexpr bool KEEPX = true;
  const int64_t stride = (int64_t)gridDim.x * nwarp;
  for (int64_t row0 = (int64_t)blockIdx.x * nwarp + warp; row0 < rows; row0 += stride * R) {
    float4 xv[R][KEEPX ? J : 1];
    uint2 dp[R][J];
    uint2 dq[R][DYF ? J : 1];
    float4 ov[R][CS ? J : 1];
    float mu[R], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r * stride;
      const bool ok = row < rows;
      const int64_t rr = ok ? row : row0;   // clamped: loads stay in bounds, results of a phantom row are discarded
      const float* xr = x + rr * ldx;
      const TDY* dyr = dy + rr * lddy;
      mu[r] = mean[rr];
      rs[r] = rstd[rr];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = (lane + 32 * j) * 4;
        if constexpr (KEEPX) xv[r][j] = *reinterpret_cast<const float4*>(xr + c);
        if constexpr (DYF) {
          const uint4 w = *reinterpret_cast<const uint4*>(dyr + c);
          dp[r][j] = make_uint2(w.x, w.y);
          dq[r][j] = make_uint2(w.z, w.w);
        } else {
          dp[r][j] = *reinterpret_cast<const uint2*>(dyr + c);
        }
        if constexpr (CS) {
          ov[r][j] = dx_accumulate ? *reinterpret_cast<const float4*>(dx + rr * lddx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r * stride;
      if (row >= rows) break;
      float sg = 0.f, sgx = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = (lane + 32 * j) * 4;
        const float4 dv = unpack(make_uint4(dp[r][j].x, dp[r][j].y, DYF ? dq[r][DYF ? j : 0].x : 0u, DYF ? dq[r][DYF ? j : 0].y : 0u));
        const float4 sc = *reinterpret_cast<const float4*>(s_sc + c);
        const float4 xq = KEEPX ? xv[r][KEEPX ? j : 0] : *reinterpret_cast<const float4*>(x + row * ldx + c);
        const float4 xh = make_float4((xq.x - mu[r]) * rs[r], (xq.y - mu[r]) * rs[r], (xq.z - mu[r]) * rs[r], (xq.w - mu[r]) * rs[r]);
        const float4 g = make_float4(dv.x * sc.x, dv.y * sc.y, dv.z * sc.z, dv.w * sc.w);
        sg += (g.x + g.y) + (g.z + g.w);
        sgx += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
        acc[j].x = fmaf(dv.x, xh.x, acc[j].x);
        acc[j].y = fmaf(dv.y, xh.y, acc[j].y);
        acc[j].z = fmaf(dv.z, xh.z, acc[j].z);
        acc[j].w = fmaf(dv.w, xh.w, acc[j].w);
      }
      sg = warp_sum(sg) * (1.f / D);
      sgx = warp_sum(sgx) * (1.f / D);
      float* dxr = dx + row * lddx;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = (lane + 32 * j) * 4;
        const float4 dv = unpack(make_uint4(dp[r][j].x, dp[r][j].y, DYF ? dq[r][DYF ? j : 0].x : 0u, DYF ? dq[r][DYF ? j : 0].y : 0u));
        const float4 sc = *reinterpret_cast<const float4*>(s_sc + c);
        const float4 xq = KEEPX ? xv[r][KEEPX ? j : 0] : *reinterpret_cast<const float4*>(x + row * ldx + c);
        const float4 xh = make_float4((xq.x - mu[r]) * rs[r], (xq.y - mu[r]) * rs[r], (xq.z - mu[r]) * rs[r], (xq.w - mu[r]) * rs[r]);
        float4 v = make_float4(rs[r] * (dv.x * sc.x - sg - xh.x * sgx), rs[r] * (dv.y * sc.y - sg - xh.y * sgx),
                               rs[r] * (dv.z * sc.z - sg - xh.z * sgx), rs[r] * (dv.w * sc.w - sg - xh.w * sgx));
        if constexpr (CS) {
          v.x += ov[r][j].x; v.y += ov[r][j].y; v.z += ov[r][j].z; v.w += ov[r][j].w;
        } else if (dx_accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(dxr + c);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *reinterpret_cast<float4*>(dxr + c) = v;
        if constexpr (CS) {
          csum[j].x += v.x; csum[j].y += v.y; csum[j].z += v.z; csum[j].w += v.w;
        }
        if (dx_lowp) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
          *reinterpret_cast<uint2*>(dx_lowp + row * ldl + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (lane + 32 * j) * 4;
    atomicAdd(&s_ds[c], acc[j].x);
    atomicAdd(&s_ds[c + 1], acc[j].y);
    atomicAdd(&s_ds[c + 2], acc[j].z);
    atomicAdd(&s_ds[c + 3], acc[j].w);
    if constexpr (CS) {
      if (dcol != nullptr) {
        atomicAdd(&s_cs[c], csum[j].x);
        atomicAdd(&s_cs[c + 1], csum[j].y);
        atomicAdd(&s_cs[c + 2], csum[j].z);
        atomicAdd(&s_cs[c + 3], csum[j].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (ds_accum) atomicAdd(&dscale_partial[i], s_ds[i]);   // straight into the gradient of the scale
    else dscale_partial[(int64_t)blockIdx.x * D + i] = s_ds[i];
    if constexpr (CS) {
      if (dcol != nullptr) atomicAdd(&dcol[i], s_cs[i]);      // sum over rows of the final dx (+= into the caller's buffer)
    }
  }
}

// Wide rows (d = 256*JH: 1024, 1280, 1536): two warps per row, each owning one half of the columns, so a thread holds
// 4*JH floats of x, dy, the old dx and the two column accumulators (no spills under 128 registers, two blocks per SM);
// the two row statistics are exchanged through shared memory under a 64-thread named barrier per warp pair.
template <int JH, typename TDY>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_wide_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ scale,
                          const float* __restrict__ mean, const float* __restrict__ rstd,
                          const TDY* __restrict__ dy, int64_t lddy, float* __restrict__ dx, int64_t lddx,
                          int dx_accumulate, bf16* __restrict__ dx_lowp, int64_t ldl,
                          float* __restrict__ dscale_partial, int ds_accum, float* __restrict__ dcol, int64_t rows) {
  constexpr int D = 256 * JH, DH_ = 128 * JH;
  constexpr bool DYF = sizeof(TDY) == 4;
  __shared__ float s_ds[D];
  __shared__ float s_cs[D];
  __shared__ __align__(16) float s_sc[D];
  __shared__ float s_red[2][8][2];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    s_ds[i] = 0.f;
    s_cs[i] = 0.f;
    s_sc[i] = scale[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1, col0 = (warp & 1) * DH_;
  float4 acc[JH], csum[JH];
#pragma unroll
  for (int j = 0; j < JH; ++j) acc[j] = csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  int it = 0;
  for (int64_t row = (int64_t)blockIdx.x * 4 + pair; row < rows; row += (int64_t)gridDim.x * 4, ++it) {
    const float* xr = x + row * ldx + col0;
    const TDY* dyr = dy + row * lddy + col0;
    float* dxr = dx + row * lddx + col0;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[JH], dv[JH], ov[JH];
#pragma unroll
    for (int j = 0; j < JH; ++j) {
      const int c = (lane + 32 * j) * 4;
      xh[j] = *reinterpret_cast<const float4*>(xr + c);
      if constexpr (DYF) {
        dv[j] = *reinterpret_cast<const float4*>(dyr + c);
      } else {
        const uint2 w = *reinterpret_cast<const uint2*>(dyr + c);
        dv[j] = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                            __uint_as_float(w.y & 0xffff0000u));
      }
      ov[j] = dx_accumulate ? *reinterpret_cast<const float4*>(dxr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < JH; ++j) {
      const int c = col0 + (lane + 32 * j) * 4;
      const float4 sc = *reinterpret_cast<const float4*>(s_sc + c);
      xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
      acc[j].x = fmaf(dv[j].x, xh[j].x, acc[j].x);
      acc[j].y = fmaf(dv[j].y, xh[j].y, acc[j].y);
      acc[j].z = fmaf(dv[j].z, xh[j].z, acc[j].z);
      acc[j].w = fmaf(dv[j].w, xh[j].w, acc[j].w);
      dv[j] = make_float4(dv[j].x * sc.x, dv[j].y * sc.y, dv[j].z * sc.z, dv[j].w * sc.w);   // g = dy * scale
      sg += (dv[j].x + dv[j].y) + (dv[j].z + dv[j].w);
      sgx += (dv[j].x * xh[j].x + dv[j].y * xh[j].y) + (dv[j].z * xh[j].z + dv[j].w * xh[j].w);
    }
    sg = warp_sum(sg);
    sgx = warp_sum(sgx);
    if (lane == 0) {
      s_red[it & 1][warp][0] = sg;
      s_red[it & 1][warp][1] = sgx;
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
    sg = (sg + s_red[it & 1][warp ^ 1][0]) * (1.f / D);
    sgx = (sgx + s_red[it & 1][warp ^ 1][1]) * (1.f / D);
#pragma unroll
    for (int j = 0; j < JH; ++j) {
      const int c = (lane + 32 * j) * 4;
      float4 v = make_float4(rs * (dv[j].x - sg - xh[j].x * sgx), rs * (dv[j].y - sg - xh[j].y * sgx),
                             rs * (dv[j].z - sg - xh[j].z * sgx), rs * (dv[j].w - sg - xh[j].w * sgx));
      v.x += ov[j].x; v.y += ov[j].y; v.z += ov[j].z; v.w += ov[j].w;
      *reinterpret_cast<float4*>(dxr + c) = v;
      csum[j].x += v.x; csum[j].y += v.y; csum[j].z += v.z; csum[j].w += v.w;
      if (dx_lowp) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(dx_lowp + row * ldl + col0 + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < JH; ++j) {
    const int c = col0 + (lane + 32 * j) * 4;
    atomicAdd(&s_ds[c], acc[j].x);
    atomicAdd(&s_ds[c + 1], acc[j].y);
    atomicAdd(&s_ds[c + 2], acc[j].z);
    atomicAdd(&s_ds[c + 3], acc[j].w);
    if (dcol != nullptr) {
      atomicAdd(&s_cs[c], csum[j].x);
      atomicAdd(&s_cs[c + 1], csum[j].y);
      atomicAdd(&s_cs[c + 2], csum[j].z);
      atomicAdd(&s_cs[c + 3], csum[j].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (ds_accum) atomicAdd(&dscale_partial[i], s_ds[i]);
    else dscale_partial[(int64_t)blockIdx.x * D + i] = s_ds[i];
    if (dcol != nullptr) atomicAdd(&dcol[i], s_cs[i]);
  }
}

// ------------------------------------------------------------------------------------------
// per-head RMSNorm (attention.py:166-167), in place; one warp per (row, head)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void head_rmsnorm_fwd_kernel(T* __restrict__ buf, int64_t ld, const float* __restrict__ scale,
                                        float out_mul, float* __restrict__ rstd_out, int64_t rstd_ld,
                                        int64_t rows, int heads, int Dh) {
  int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= rows * heads) return;
  int lane = threadIdx.x & 31;
  int64_t r = wid / heads;
  int h = (int)(wid % heads);
  T* p = buf + r * ld + (int64_t)h * Dh;
  float ss = 0.f;
  for (int i = lane; i < Dh; i += 32) {
    float v = ldf<T>(p + i);
    ss += v * v;
  }
  ss = warp_sum(ss);
  float rstd = rsqrtf(ss / Dh + kNormEps);
  if (lane == 0 && rstd_out) rstd_out[r * rstd_ld + h] = rstd;
  for (int i = lane; i < Dh; i += 32) stf<T>(p + i, ldf<T>(p + i) * rstd * scale[i] * out_mul);
}

// y = x*rstd*scale*mul.  Given dy: g = dy*scale*mul; dx = rstd*(g - xhat*mean(g*xhat)), xhat = x*rstd.
// xhat is recovered from the saved output: xhat = y/(scale*mul) (scale==0 => that channel's
// output carries no information and its g is 0 as well).
template <typename TY, typename TD>
__global__ void head_rmsnorm_bwd_kernel(const TY* __restrict__ y, int64_t ldy,
                                        const float* __restrict__ scale, float out_mul,
                                        const float* __restrict__ rstd, int64_t rstd_ld,
                                        TD* __restrict__ d_io, int64_t ldd,
                                        float* __restrict__ dscale_partial, int ds_accum, int64_t rows, int heads,
                                        int Dh) {
  extern __shared__ float s_ds[];  // [Dh]
  for (int i = threadIdx.x; i < Dh; i += blockDim.x) s_ds[i] = 0.f;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int64_t total = rows * heads;
  for (int64_t wid = (int64_t)blockIdx.x * nwarp + warp; wid < total; wid += (int64_t)gridDim.x * nwarp) {
    int64_t r = wid / heads;
    int h = (int)(wid % heads);
    const TY* yp = y + r * ldy + (int64_t)h * Dh;
    TD* dp = d_io + r * ldd + (int64_t)h * Dh;
    float rs = rstd[r * rstd_ld + h];
    float sgx = 0.f;
    for (int i = lane; i < Dh; i += 32) {
      float sm = scale[i] * out_mul;
      float xh = sm != 0.f ? ldf<TY>(yp + i) / sm : 0.f;
      float dyv = ldf<TD>(dp + i);
      sgx += dyv * sm * xh;
      atomicAdd(&s_ds[i], dyv * xh * out_mul);
    }
    sgx = warp_sum(sgx) / Dh;
    for (int i = lane; i < Dh; i += 32) {
      float sm = scale[i] * out_mul;
      float xh = sm != 0.f ? ldf<TY>(yp + i) / sm : 0.f;
      float g = ldf<TD>(dp + i) * sm;
      stf<TD>(dp + i, rs * (g - xh * sgx));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Dh; i += blockDim.x) {
    if (ds_accum) atomicAdd(&dscale_partial[i], s_ds[i]);   // straight into the gradient of the scale
    else dscale_partial[(int64_t)blockIdx.x * Dh + i] = s_ds[i];
  }
}


// Bandwidth version of the above for bf16 buffers: 4 lanes per (row, head), each lane owns CPL =
// Dh/32 16-byte chunks (chunks lane, lane+4, ...), the head's statistics are two shuffles, the
// scale gradient lives in per-lane registers for the whole grid-stride loop (no atomics inside).
template <int CPL>
__global__ void __launch_bounds__(256, 2)
head_rmsnorm_bwd_fast_kernel(const bf16* __restrict__ y, int64_t ldy, const float* __restrict__ scale,
                             float out_mul, const float* __restrict__ rstd, int64_t rstd_ld,
                             bf16* __restrict__ d_io, int64_t ldd, float* __restrict__ dscale_partial, int ds_accum,
                             int64_t rows, int heads) {
  constexpr int DH = CPL * 32;
  __shared__ float s_ds[DH];
  __shared__ __align__(16) float s_sm[DH], s_ism[DH];   // scale*mul and its reciprocal
  for (int i = threadIdx.x; i < DH; i += blockDim.x) {
    s_ds[i] = 0.f;
    const float v = scale[i] * out_mul;
    s_sm[i] = v;
    s_ism[i] = v != 0.f ? 1.f / v : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
  float acc[CPL][8];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  const int64_t total = rows * heads;
  const int64_t gstride = (int64_t)gridDim.x * (blockDim.x >> 2);
  for (int64_t wb = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8; wb < total; wb += gstride) {
    // a warp takes 8 consecutive (row, head) pairs; groups beyond the end idle but stay in the shuffles
    const int64_t wid = wb + grp;
    const bool live = wid < total;
    const int64_t r = live ? wid / heads : 0;
    const int h = live ? (int)(wid % heads) : 0;
    const bf16* yp = y + r * ldy + (int64_t)h * DH;
    bf16* dp = d_io + r * ldd + (int64_t)h * DH;
    uint4 yv[CPL], dv[CPL];   // the row stays packed (bf16) in registers between the two passes
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      yv[c] = make_uint4(0, 0, 0, 0);
      dv[c] = make_uint4(0, 0, 0, 0);
      if (live) {
        yv[c] = *reinterpret_cast<const uint4*>(yp + (sub + 4 * c) * 8);
        dv[c] = *reinterpret_cast<const uint4*>(dp + (sub + 4 * c) * 8);
      }
    }
    float sgx = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const uint32_t yw[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w}, dw[4] = {dv[c].x, dv[c].y, dv[c].z, dv[c].w};
      const float* smp = s_sm + (sub + 4 * c) * 8;
      const float* isp = s_ism + (sub + 4 * c) * 8;
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const __nv_bfloat162 yh = *reinterpret_cast<const __nv_bfloat162*>(&yw[e2]);
        const __nv_bfloat162 dh = *reinterpret_cast<const __nv_bfloat162*>(&dw[e2]);
        const float d0 = __low2float(dh), d1 = __high2float(dh);
        const float x0 = __low2float(yh) * isp[2 * e2], x1 = __high2float(yh) * isp[2 * e2 + 1];
        sgx = fmaf(d0 * smp[2 * e2], x0, sgx);
        sgx = fmaf(d1 * smp[2 * e2 + 1], x1, sgx);
        acc[c][2 * e2] = fmaf(d0 * out_mul, x0, acc[c][2 * e2]);
        acc[c][2 * e2 + 1] = fmaf(d1 * out_mul, x1, acc[c][2 * e2 + 1]);
      }
    }
    sgx += __shfl_xor_sync(0xffffffffu, sgx, 1);
    sgx += __shfl_xor_sync(0xffffffffu, sgx, 2);
    sgx *= 1.f / DH;
    if (live) {
      const float rs = rstd[r * rstd_ld + h];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const uint32_t yw[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w}, dw[4] = {dv[c].x, dv[c].y, dv[c].z, dv[c].w};
        const float* smp = s_sm + (sub + 4 * c) * 8;
        const float* isp = s_ism + (sub + 4 * c) * 8;
        uint32_t w[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const __nv_bfloat162 yh = *reinterpret_cast<const __nv_bfloat162*>(&yw[e2]);
          const __nv_bfloat162 dh = *reinterpret_cast<const __nv_bfloat162*>(&dw[e2]);
          const float x0 = __low2float(yh) * isp[2 * e2], x1 = __high2float(yh) * isp[2 * e2 + 1];
          const float g0 = __low2float(dh) * smp[2 * e2], g1 = __high2float(dh) * smp[2 * e2 + 1];
          __nv_bfloat162 o = __floats2bfloat162_rn(rs * (g0 - x0 * sgx), rs * (g1 - x1 * sgx));
          w[e2] = *reinterpret_cast<uint32_t*>(&o);
        }
        *reinterpret_cast<uint4*>(dp + (sub + 4 * c) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  // reduce the per-lane scale gradients: across the 8 groups of a warp, then across warps
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[c][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (grp == 0) atomicAdd(&s_ds[(sub + 4 * c) * 8 + e], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < DH; i += blockDim.x) {
    if (ds_accum) atomicAdd(&dscale_partial[i], s_ds[i]);   // straight into the gradient of the scale
    else dscale_partial[(int64_t)blockIdx.x * DH + i] = s_ds[i];
  }
}

// ------------------------------------------------------------------------------------------
// key mask (R1), quantiser, decoder tokens, output split, loss
// ------------------------------------------------------------------------------------------
__global__ void key_mask_kernel(const float* __restrict__ visible, const int32_t* __restrict__ boundary,
                                uint8_t* __restrict__ mask, int B, int N, int T, int ro) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int L = T + ro;
  int64_t total = (int64_t)B * N * L;
  if (idx >= total) return;
  int j = (int)(idx % L);
  int64_t bn = idx / L;
  int b = (int)(bn / N);
  uint8_t m;
  if (ro && j == 0) {
    m = 1;
  } else {
    int t = j - ro;
    m = (visible[bn * T + t] != 0.f) && (t < boundary[b]);
  }
  mask[idx] = m;
}

__global__ void quantize_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                float* __restrict__ y, uint8_t* __restrict__ pass, int64_t n,
                                int discretize) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (pass) pass[i] = (v >= -1.f && v <= 1.f);
  v = fminf(fmaxf(v, -1.f), 1.f);
  if (discretize) {
    float d = __fdiv_rn(rintf(__fmul_rn(v, 128.f)), 128.f);
    d = __fsub_rn(__fadd_rn(d, __fdiv_rn(noise[i], 128.f)), 1.0f / 256.0f);
    // straight-through: latents - stop_gradient(latents - disc)
    v = __fsub_rn(v, __fsub_rn(v, d));
  }
  y[i] = v;
}

__global__ void quantize_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ pass,
                                    float* __restrict__ dx, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = pass[i] ? dy[i] : 0.f;
}

template <typename TL, typename TQ, typename TT>
__global__ void decoder_tokens_fwd_kernel(const TL* __restrict__ lat, const TQ* __restrict__ qe,
                                          const int32_t* __restrict__ qframe, TT* __restrict__ tok,
                                          int B, int Q, int L, int C) {
  // one block per (b, q, token); threads over channels
  int64_t row = blockIdx.x;  // (b*Q + q)*(L+1) + tkn
  int tkn = (int)(row % (L + 1));
  int64_t bq = row / (L + 1);
  int b = (int)(bq / Q);
  int D = C + 128;
  TT* o = tok + row * D;
  if (tkn == 0) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) stf<TT>(o + c, ldf<TQ>(qe + bq * D + c));
    return;
  }
  const TL* l = lat + ((int64_t)b * L + (tkn - 1)) * C;
  int off = qframe[bq] * 5;
  auto value = [&](int c) -> float {
    if (c < C) return ldf<TL>(l + c);
    const int src = c - C + off;
    return (src >= 0 && src < C) ? ldf<TL>(l + src) : 0.f;
  };
  if constexpr (sizeof(TT) == 4) {
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(tok) & 15) == 0) {   // 128-bit stores (the latents are small and cached)
      for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
        *reinterpret_cast<float4*>(o + c) = make_float4(value(c), value(c + 1), value(c + 2), value(c + 3));
      return;
    }
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) stf<TT>(o + c, value(c));
}

// d_lat[b,n,c] = sum_q d_tok[b,q,1+n,c] + sum_q [0 <= c-5t_q < 128] d_tok[b,q,1+n,C + c-5t_q]
// The sum over the Q queries is split QS ways across blocks (B*L alone is 128 blocks of strictly serial loads) and the
// partial sums are added atomically into the zeroed d_lat.
template <typename TT>
__global__ void decoder_tokens_bwd_kernel(const TT* __restrict__ d_tok, const int32_t* __restrict__ qframe,
                                          float* __restrict__ d_lat, float* __restrict__ d_qe, int B,
                                          int Q, int L, int C, int QS) {
  int D = C + 128;
  const int64_t lat_blocks = (int64_t)B * L * QS;
  if ((int64_t)blockIdx.x >= lat_blocks) {   // query rows: d_qe[b,q,:] = d_tok[b,q,0,:]
    int64_t bq = (int64_t)blockIdx.x - lat_blocks;
    const TT* src = d_tok + bq * (L + 1) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) d_qe[bq * D + c] = ldf<TT>(src + c);
    return;
  }
  const int64_t bn = blockIdx.x / QS;   // b*L + n
  const int split = (int)(blockIdx.x % QS);
  const int qper = (Q + QS - 1) / QS;
  const int q0 = split * qper, q1 = min(Q, q0 + qper);
  int b = (int)(bn / L), n = (int)(bn % L);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
    for (int q = q0; q < q1; ++q) {
      int64_t bq = (int64_t)b * Q + q;
      const TT* src = d_tok + (bq * (L + 1) + 1 + n) * D;
      acc0 += ldf<TT>(src + c);
      int w = c - qframe[bq] * 5;
      if (w >= 0 && w < 128) acc1 += ldf<TT>(src + C + w);
    }
    atomicAdd(&d_lat[bn * C + c], acc0 + acc1);
  }
}

__global__ void split_outputs_kernel(const float* __restrict__ ho, float* __restrict__ tracks,
                                     float* __restrict__ vis, float* __restrict__ cert, int64_t rows,
                                     int T, int coords) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * T) return;
  int64_t r = idx / T;
  int t = (int)(idx % T);
  const float* h = ho + r * 4 * T;
  for (int c = 0; c < coords; ++c) tracks[idx * coords + c] = h[c * T + t];
  vis[idx] = h[coords * T + t];
  if (cert) cert[idx] = (coords == 3) ? 0.f : h[3 * T + t];
}

__device__ __forceinline__ float log_sigmoid(float x) {
  // stable: min(x,0) - log1p(exp(-|x|))
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

__global__ void loss_fwd_kernel(const float* __restrict__ ho, const float* __restrict__ tt,
                                const float* __restrict__ tv, float* __restrict__ sums, int64_t rows,
                                int T) {
  float pos = 0.f, bce = 0.f, nv = 0.f;
  int64_t total = rows * T;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx / T;
    int t = (int)(idx % T);
    const float* h = ho + r * 4 * T;
    float v = tv[idx];
    float e = fabsf(h[t] - tt[idx * 3]) + fabsf(h[T + t] - tt[idx * 3 + 1]) + fabsf(h[2 * T + t] - tt[idx * 3 + 2]);
    pos += e * v;
    float l = h[3 * T + t];
    bce += -v * log_sigmoid(l) - (1.f - v) * log_sigmoid(-l);
    nv += v;
  }
  __shared__ float sh[3][32];
  pos = warp_sum(pos); bce = warp_sum(bce); nv = warp_sum(nv);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = pos; sh[1][w] = bce; sh[2][w] = nv; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    pos = lane < nw ? sh[0][lane] : 0.f;
    bce = lane < nw ? sh[1][lane] : 0.f;
    nv = lane < nw ? sh[2][lane] : 0.f;
    pos = warp_sum(pos); bce = warp_sum(bce); nv = warp_sum(nv);
    if (lane == 0) { atomicAdd(sums, pos); atomicAdd(sums + 1, bce); atomicAdd(sums + 2, nv); }
  }
}

__global__ void loss_bwd_kernel(const float* __restrict__ ho, const float* __restrict__ tt,
                                const float* __restrict__ tv, float* __restrict__ dho, float l1w,
                                float bcew, float inv_denom, int64_t rows, int T) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * T) return;
  int64_t r = idx / T;
  int t = (int)(idx % T);
  const float* h = ho + r * 4 * T;
  float* d = dho + r * 4 * T;
  float v = tv[idx];
  float s = l1w * inv_denom * v;
  for (int c = 0; c < 3; ++c) {
    float diff = h[c * T + t] - tt[idx * 3 + c];
    d[c * T + t] = diff > 0.f ? s : (diff < 0.f ? -s : 0.f);
  }
  float l = h[3 * T + t];
  float sig = 1.f / (1.f + expf(-l));
  d[3 * T + t] = bcew * inv_denom * (sig - v);
}

template <typename TP, typename TDY, typename TDX>
__global__ void gelu_bwd_kernel(const TP* __restrict__ pre, int64_t ldp, const TDY* __restrict__ dy,
                                int64_t lddy, TDX* __restrict__ dx, int64_t lddx, int64_t rows, int cols) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols;
  int c = (int)(idx % cols);
  stf<TDX>(dx + r * lddx + c, ldf<TDY>(dy + r * lddy + c) * gelu_tanh_grad(ldf<TP>(pre + r * ldp + c)));
}

// column sums: block handles 32 columns x a slab of rows; atomics across slabs
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out,
                              int64_t rows, int cols, int64_t rows_per_block) {
  __shared__ float sh[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int ry = threadIdx.x >> 5;  // 0..7
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += ldf<T>(x + r * ldx + c);
  sh[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x & 31];
    atomicAdd(out + c, s);
  }
}

// 16-byte loads: thread = one 16-byte column group (4 fp32 / 8 bf16) x every 8th row of a slab, four
// independent accumulation streams in flight; 512 contiguous bytes per warp instruction
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t rows, int cols,
                  int64_t rows_per_block) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float sh[8][32][V + 1];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * V;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[4][V];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[u][e] = 0.f;
  if (c < cols) {
    for (int64_t r = r0 + ry; r < r1; r += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + 8 * u;
        if (rr < r1) {
          const uint4 w = *reinterpret_cast<const uint4*>(x + rr * ldx + c);
          if constexpr (sizeof(T) == 4) {
            acc[u][0] += __uint_as_float(w.x); acc[u][1] += __uint_as_float(w.y);
            acc[u][2] += __uint_as_float(w.z); acc[u][3] += __uint_as_float(w.w);
          } else {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&ww[e]);
              acc[u][2 * e] += __low2float(h);
              acc[u][2 * e + 1] += __high2float(h);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) sh[ry][lane][e] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
  __syncthreads();
  if (ry == 0 && c < cols) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sh[i][lane][e];
      atomicAdd(out + c + e, t);
    }
  }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += alpha * x[i];
}

__global__ void zero_kernel(float* __restrict__ p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// sum of squares in a FIXED order (per-block partials, then one block adds them in index order):
// every data-parallel replica must derive the identical clip scale from the identical all-reduced
// gradient, or the replicas drift apart bit by bit.
__global__ void sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    acc += v * v;
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = acc;
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    acc = lane < nw ? sh[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) partial[blockIdx.x] = acc;
  }
}
__global__ void sumsq_final_kernel(const float* __restrict__ partial, int nb, float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  acc = warp_sum(acc);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = acc;
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    acc = lane < nw ? sh[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) *out += acc;
  }
}

// optax.chain(clip_by_global_norm(c), adamw(lr, wd)) - train.py:239-243
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, const float* __restrict__ sumsq,
                             float clip_norm, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gn = sqrtf(*sumsq);
  float sc = gn < clip_norm ? 1.f : clip_norm / gn;
  float gi = g[i] * sc;
  float mi = b1 * m[i] + (1.f - b1) * gi;
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float mh = mi / bc1, vh = vi / bc2;
  float pi = p[i];
  p[i] = pi - lr * (mh / (sqrtf(vh) + eps) + wd * pi);
}

template <typename TT, typename TO>
__global__ void masked_mean_kernel(const TT* __restrict__ tok, int64_t ldt, const float* __restrict__ vis,
                                   TO* __restrict__ out, int64_t ldo, int T, int W) {
  int64_t s = blockIdx.x;
  __shared__ float cnt_s;
  if (threadIdx.x < 32) {
    float c = 0.f;
    for (int t = threadIdx.x; t < T; t += 32) c += vis[s * T + t] != 0.f ? 1.f : 0.f;
    c = warp_sum(c);
    if (threadIdx.x == 0) cnt_s = fmaxf(c, 1.f);
  }
  __syncthreads();
  float cnt = cnt_s;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t)
      if (vis[s * T + t] != 0.f) acc += ldf<TT>(tok + (s * T + t) * ldt + c);
    stf<TO>(out + s * ldo + c, acc / cnt);
  }
}

template <typename TX, typename TY>
__global__ void gelu_fwd_kernel(const TX* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy,
                                int64_t rows, int cols) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols;
  int c = (int)(idx % cols);
  stf<TY>(y + r * ldy + c, gelu_tanh(ldf<TX>(x + r * ldx + c)));
}

// compute-dtype shadows of one fp32 master matrix in one pass: dst[r,c] = T(src[r,c]) and dst_t[c,r] = T(src[r,c])
// (the [out,in] operand of the forward / weight-gradient GEMMs and the [in,out] operand of dX = dY . W); 32 x 32 tiles through
// shared memory so both stores are coalesced.
template <typename T>
__global__ void shadow_weights_kernel(const float* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd,
                                      T* __restrict__ dst_t, int64_t ldt, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[(int64_t)r * lds + c];
      if (dst != nullptr) stf<T>(dst + (int64_t)r * ldd + c, v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
  if (dst_t == nullptr) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;   // transposed: row index of dst_t = column of src
    if (r < rows && c < cols) stf<T>(dst_t + (int64_t)c * ldt + r, tile[tx][ty + 8 * i]);
  }
}

// x = hi + mid + lo with three bf16 terms (24 mantissa bits): dst[r, p*K + c] = part p of src[r, c]
__global__ void split3_kernel(const float* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd, int64_t rows, int K4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, group of 4 columns)
  if (i >= rows * K4) return;
  const int64_t r = i / K4;
  const int c4 = (int)(i % K4);
  const float4 x = *reinterpret_cast<const float4*>(src + r * lds + c4 * 4);
  const float xs[4] = {x.x, x.y, x.z, x.w};
  __nv_bfloat16 part[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h = __float2bfloat16_rn(xs[e]);
    const float r1 = xs[e] - __bfloat162float(h);               // exact
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);                  // exact
    part[0][e] = h; part[1][e] = m; part[2][e] = __float2bfloat16_rn(r2);
  }
  const int64_t K = (int64_t)K4 * 4;
#pragma unroll
  for (int p = 0; p < 3; ++p)
    *reinterpret_cast<uint2*>(dst + r * ldd + p * K + c4 * 4) = *reinterpret_cast<const uint2*>(part[p]);
}

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int head_rmsnorm_fwd_impl(void* buf, int64_t ld, int dtype, const float* scale, float out_mul,
                          float* rstd_out, int64_t rstd_ld, int64_t rows, int heads, int Dh,
                          cudaStream_t st) {
  if (rows == 0 || heads == 0) return 0;
  SPA3D_DISPATCH(dtype, T, {
    head_rmsnorm_fwd_kernel<T><<<blocks_for(rows * heads, 8), 256, 0, st>>>((T*)buf, ld, scale, out_mul, rstd_out, rstd_ld, rows, heads, Dh);
  });
  return check_launch("head_rmsnorm_fwd");
}

// evaluate_tapvid3d.py:39-59: predictions [Q,T,C] / logits [Q,T] -> TAPVid-3D order [T,Q,C], occluded = logit <= 0;
// optional per-point reconstruction error |pred - target|_2 in the same [T,Q] order (the visualiser's coords_score).
__global__ void to_tapvid3d_kernel(const float* __restrict__ tracks, const float* __restrict__ logits,
                                   const float* __restrict__ target, float* __restrict__ out_tracks,
                                   uint8_t* __restrict__ out_occ, float* __restrict__ out_score, int64_t Q, int T, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output order: t-major
  if (idx >= Q * T) return;
  const int64_t t = idx / Q, q = idx % Q;
  const int64_t src = q * T + t;
  float err = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = tracks[src * C + c];
    out_tracks[idx * C + c] = v;
    if (out_score != nullptr) {
      const float d = v - target[src * C + c];
      err = fmaf(d, d, err);
    }
  }
  out_occ[idx] = logits[src] <= 0.f ? 1 : 0;
  if (out_score != nullptr) out_score[idx] = sqrtf(err);
}

}  // namespace spa3d

using namespace spa3d;

extern "C" {

int spa3d_fourier_features(const float* x, int64_t ldx, void* out, int64_t ldo, int out_dtype,
                           int64_t rows, int C, int num_freq, float scale_factor, int append_time,
                           int tail_zero, int exact, int out_row_group, void* stream) {
  SPA3D_REQUIRE(num_freq > 0 && num_freq <= 64, "fourier: num_freq must be in 1..64");
  if (rows == 0) return 0;
  FourierScales sc;
  for (int i = 0; i < num_freq; ++i) sc.s[i] = (float)pow(2.0, (double)i / 3.0);
  int Ctot = C + (append_time > 0 ? 1 : 0) + (tail_zero ? 1 : 0);
  int64_t total = rows * Ctot * num_freq;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(out_dtype, TO, {
    if (exact)
      fourier_kernel<TO, true><<<blocks_for(total, 256), 256, 0, st>>>(x, ldx, (TO*)out, ldo, rows, C, Ctot, num_freq, scale_factor, append_time, out_row_group, sc);
    else
      fourier_kernel<TO, false><<<blocks_for(total, 256), 256, 0, st>>>(x, ldx, (TO*)out, ldo, rows, C, Ctot, num_freq, scale_factor, append_time, out_row_group, sc);
  });
  return check_launch("fourier_features");
}

int spa3d_set_rows(void* dst, int64_t ld, int dst_dtype, int64_t row_stride, const float* vec,
                   int64_t rows, int cols, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(dst_dtype, TD, {
    set_rows_kernel<TD><<<blocks_for(rows * cols, 256), 256, 0, st>>>((TD*)dst, ld, row_stride, vec, rows, cols);
  });
  return check_launch("set_rows");
}

int spa3d_convert(const void* src, int64_t lds, int src_dtype, void* dst, int64_t ldd, int dst_dtype,
                  int64_t rows, int cols, int out_row_group, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    auto al = [](const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const size_t ss = src_dtype == SPA3D_F32 ? 4 : 2, ds = dst_dtype == SPA3D_F32 ? 4 : 2;
    if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && al(src, 4 * ss) && al(dst, 4 * ds) &&
        (src_dtype == SPA3D_F32 || src_dtype == SPA3D_BF16) && (dst_dtype == SPA3D_F32 || dst_dtype == SPA3D_BF16)) {
      SPA3D_DISPATCH(src_dtype, TS, SPA3D_DISPATCH(dst_dtype, TD, {
        convert_vec4_kernel<TS, TD><<<blocks_for(rows * (cols / 4), 256), 256, 0, st>>>((const TS*)src, lds, (TD*)dst, ldd, rows, cols / 4, out_row_group);
      }));
      return check_launch("convert_vec4");
    }
  }
  SPA3D_DISPATCH(src_dtype, TS, SPA3D_DISPATCH(dst_dtype, TD, {
    convert_kernel<TS, TD><<<blocks_for(rows * cols, 256), 256, 0, st>>>((const TS*)src, lds, (TD*)dst, ldd, rows, cols, out_row_group);
  }));
  return check_launch("convert");
}

int spa3d_layernorm_fwd(const void* x, int64_t ldx, int x_dtype, const float* scale, void* y,
                        int64_t ldy, int y_dtype, float* mean_out, float* rstd_out, int64_t rows,
                        int d, void* stream) {
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const bool ybf = y_dtype == SPA3D_BF16;
    const bool xbf = x_dtype == SPA3D_BF16;
    if ((x_dtype == SPA3D_F32 || (xbf && ybf)) && (ybf || y_dtype == SPA3D_F32) && d % 128 == 0 && d <= 1536 && ldx % 4 == 0 && ldy % 4 == 0 &&
        al(x, xbf ? 8 : 16) && al(scale, 16) && al(y, ybf ? 8 : 16)) {
#define SPA3D_LN_FWD_FAST(J)                                                                                              \
  case J:                                                                                                                 \
    if (xbf) layernorm_fwd_fast_kernel<J, bf16, bf16><<<blocks_for(rows, 8), 256, 0, st>>>((const bf16*)x, ldx, scale, (bf16*)y, ldy, mean_out, rstd_out, rows); \
    else if (ybf) layernorm_fwd_fast_kernel<J, bf16><<<blocks_for(rows, 8), 256, 0, st>>>((const float*)x, ldx, scale, (bf16*)y, ldy, mean_out, rstd_out, rows); \
    else layernorm_fwd_fast_kernel<J, float><<<blocks_for(rows, 8), 256, 0, st>>>((const float*)x, ldx, scale, (float*)y, ldy, mean_out, rstd_out, rows);   \
    return check_launch("layernorm_fwd_fast");
      switch (d / 128) {
        SPA3D_LN_FWD_FAST(1) SPA3D_LN_FWD_FAST(2) SPA3D_LN_FWD_FAST(3) SPA3D_LN_FWD_FAST(4) SPA3D_LN_FWD_FAST(8)
        SPA3D_LN_FWD_FAST(9) SPA3D_LN_FWD_FAST(10) SPA3D_LN_FWD_FAST(12)
        default: break;
      }
#undef SPA3D_LN_FWD_FAST
    }
  }
  SPA3D_DISPATCH(x_dtype, TX, SPA3D_DISPATCH(y_dtype, TY, {
    layernorm_fwd_kernel<TX, TY><<<blocks_for(rows, 8), 256, 0, st>>>((const TX*)x, ldx, scale, (TY*)y, ldy, mean_out, rstd_out, rows, d);
  }));
  return check_launch("layernorm_fwd");
}

int spa3d_layernorm_bwd(const void* x, int64_t ldx, int x_dtype, const float* scale,
                        const float* mean, const float* rstd, const void* dy, int64_t lddy,
                        int dy_dtype, void* dx, int64_t lddx, int dx_dtype, int dx_accumulate,
                        void* dx_lowp, int64_t ldl, float* dscale_partial, int num_partials, int accumulate_dscale,
                        float* dx_colsum, int64_t rows, int d, void* stream) {
  SPA3D_REQUIRE(num_partials > 0, "layernorm_bwd: num_partials must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool fast = x_dtype == SPA3D_F32 && dx_dtype == SPA3D_F32 && d % 128 == 0 && d <= 1536 && ldx % 4 == 0 &&
                    lddx % 4 == 0 && lddy % 4 == 0 && al16(x) && al16(dx) && al16(dy) && al16(scale) &&
                    (!dx_lowp || (ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(dx_lowp) & 7) == 0));
  const bool wide = fast && (d == 1024 || d == 1280 || d == 1536);
  SPA3D_REQUIRE(dx_colsum == nullptr || (fast && (d <= 512 || wide)), "layernorm_bwd: dx_colsum needs the fp32 fast path with d <= 512 or d in {1024, 1280, 1536}");
  if (wide) {
#define SPA3D_LN_BWD_WIDE(JH)                                                                                  \
  case JH:                                                                                                     \
    SPA3D_DISPATCH(dy_dtype, TDY, {                                                                            \
      layernorm_bwd_wide_kernel<JH, TDY><<<num_partials, 256, 0, st>>>((const float*)x, ldx, scale, mean, rstd, \
          (const TDY*)dy, lddy, (float*)dx, lddx, dx_accumulate, (bf16*)dx_lowp, ldl, dscale_partial, accumulate_dscale,   \
          dx_colsum, rows);                                                                                    \
    });                                                                                                        \
    return check_launch("layernorm_bwd_wide");
    switch (d / 256) {
      SPA3D_LN_BWD_WIDE(4) SPA3D_LN_BWD_WIDE(5) SPA3D_LN_BWD_WIDE(6)
      default: break;
    }
#undef SPA3D_LN_BWD_WIDE
  }
  if (fast) {
#define SPA3D_LN_BWD_FAST(J)                                                                                   \
  case J:                                                                                                      \
    SPA3D_DISPATCH(dy_dtype, TDY, {                                                                            \
      layernorm_bwd_fast_kernel<J, TDY><<<num_partials, 256, 0, st>>>((const float*)x, ldx, scale, mean, rstd, \
          (const TDY*)dy, lddy, (float*)dx, lddx, dx_accumulate, (bf16*)dx_lowp, ldl, dscale_partial, accumulate_dscale,   \
          dx_colsum, rows);                                                                                    \
    });                                                                                                        \
    return check_launch("layernorm_bwd_fast");
    switch (d / 128) {
      SPA3D_LN_BWD_FAST(1) SPA3D_LN_BWD_FAST(2) SPA3D_LN_BWD_FAST(3) SPA3D_LN_BWD_FAST(4) SPA3D_LN_BWD_FAST(8)
      SPA3D_LN_BWD_FAST(9) SPA3D_LN_BWD_FAST(10) SPA3D_LN_BWD_FAST(12)
      default: break;
    }
#undef SPA3D_LN_BWD_FAST
  }
  SPA3D_DISPATCH(x_dtype, TX, SPA3D_DISPATCH(dy_dtype, TDY, SPA3D_DISPATCH(dx_dtype, TDX, {
    layernorm_bwd_kernel<TX, TDY, TDX><<<num_partials, 256, d * sizeof(float), st>>>(
        (const TX*)x, ldx, scale, mean, rstd, (const TDY*)dy, lddy, (TDX*)dx, lddx, dx_accumulate, dscale_partial, accumulate_dscale, rows, d);
  })));
  int rc = check_launch("layernorm_bwd");
  if (rc || !dx_lowp) return rc;
  return spa3d_convert(dx, lddx, dx_dtype, dx_lowp, ldl, SPA3D_BF16, rows, d, 0, stream);
}

int spa3d_head_rmsnorm_fwd(void* buf, int64_t ld, int dtype, const float* scale, float out_mul,
                           float* rstd_out, int64_t rstd_ld, int64_t rows, int heads, int Dh,
                           void* stream) {
  return spa3d::head_rmsnorm_fwd_impl(buf, ld, dtype, scale, out_mul, rstd_out, rstd_ld, rows, heads, Dh,
                                      (cudaStream_t)stream);
}

int spa3d_head_rmsnorm_bwd(const void* y, int64_t ldy, int y_dtype, const float* scale, float out_mul,
                           const float* rstd, int64_t rstd_ld, void* dy_inout, int64_t ldd, int d_dtype,
                           float* dscale_partial, int num_partials, int accumulate_dscale, int64_t rows, int heads, int Dh,
                           void* stream) {
  SPA3D_REQUIRE(num_partials > 0, "head_rmsnorm_bwd: num_partials must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (y_dtype == SPA3D_BF16 && d_dtype == SPA3D_BF16 && Dh % 32 == 0 && Dh <= 128 && ldy % 8 == 0 && ldd % 8 == 0 &&
      al16(y) && al16(dy_inout)) {
#define SPA3D_RMS_BWD_FAST(CPL)                                                                              \
  head_rmsnorm_bwd_fast_kernel<CPL><<<num_partials, 256, 0, st>>>((const bf16*)y, ldy, scale, out_mul, rstd, \
                                                                  rstd_ld, (bf16*)dy_inout, ldd, dscale_partial, accumulate_dscale, rows, heads)
    switch (Dh / 32) {
      case 1: SPA3D_RMS_BWD_FAST(1); break;
      case 2: SPA3D_RMS_BWD_FAST(2); break;
      case 3: SPA3D_RMS_BWD_FAST(3); break;
      default: SPA3D_RMS_BWD_FAST(4); break;
    }
#undef SPA3D_RMS_BWD_FAST
    return check_launch("head_rmsnorm_bwd_fast");
  }
  SPA3D_DISPATCH(y_dtype, TY, SPA3D_DISPATCH(d_dtype, TD, {
    head_rmsnorm_bwd_kernel<TY, TD><<<num_partials, 256, Dh * sizeof(float), st>>>(
        (const TY*)y, ldy, scale, out_mul, rstd, rstd_ld, (TD*)dy_inout, ldd, dscale_partial, accumulate_dscale, rows, heads, Dh);
  }));
  return check_launch("head_rmsnorm_bwd");
}

int spa3d_masked_mean_fwd(const void* tok, int64_t ldt, int tok_dtype, const float* visible, void* out,
                          int64_t ldo, int out_dtype, int64_t S, int T, int W, void* stream) {
  if (S == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(tok_dtype, TT, SPA3D_DISPATCH(out_dtype, TO, {
    masked_mean_kernel<TT, TO><<<(unsigned)S, 128, 0, st>>>((const TT*)tok, ldt, visible, (TO*)out, ldo, T, W);
  }));
  return check_launch("masked_mean_fwd");
}

int spa3d_gelu_fwd(const void* x, int64_t ldx, int x_dtype, void* y, int64_t ldy, int y_dtype,
                   int64_t rows, int cols, void* stream) {
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(x_dtype, TX, SPA3D_DISPATCH(y_dtype, TY, {
    gelu_fwd_kernel<TX, TY><<<blocks_for(rows * cols, 256), 256, 0, st>>>((const TX*)x, ldx, (TY*)y, ldy, rows, cols);
  }));
  return check_launch("gelu_fwd");
}

int spa3d_build_key_mask(const float* visible, const int32_t* boundary_frame, uint8_t* mask, int B,
                         int N, int T, int has_readout, void* stream) {
  int64_t total = (int64_t)B * N * (T + (has_readout ? 1 : 0));
  if (total == 0) return 0;
  key_mask_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(visible, boundary_frame, mask, B, N, T, has_readout ? 1 : 0);
  return check_launch("build_key_mask");
}

int spa3d_quantize_fwd(const float* x, const float* noise, float* y, uint8_t* pass_mask, int64_t n,
                       int discretize, void* stream) {
  SPA3D_REQUIRE(!discretize || noise, "quantize: discretize needs a noise tensor");
  if (n == 0) return 0;
  quantize_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, noise, y, pass_mask, n, discretize);
  return check_launch("quantize_fwd");
}

int spa3d_quantize_bwd(const float* dy, const uint8_t* pass_mask, float* dx, int64_t n, void* stream) {
  if (n == 0) return 0;
  quantize_bwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, pass_mask, dx, n);
  return check_launch("quantize_bwd");
}

int spa3d_decoder_tokens_fwd(const void* lat, int lat_dtype, const void* query_emb, int qe_dtype,
                             const int32_t* query_frame, void* tokens, int tok_dtype, int B, int Q,
                             int L, int C, void* stream) {
  int64_t rows = (int64_t)B * Q * (L + 1);
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(lat_dtype, TL, SPA3D_DISPATCH(qe_dtype, TQ, SPA3D_DISPATCH(tok_dtype, TT, {
    decoder_tokens_fwd_kernel<TL, TQ, TT><<<(unsigned)rows, 256, 0, st>>>((const TL*)lat, (const TQ*)query_emb, query_frame, (TT*)tokens, B, Q, L, C);
  })));
  return check_launch("decoder_tokens_fwd");
}

int spa3d_decoder_tokens_bwd(const void* d_tokens, int tok_dtype, const int32_t* query_frame,
                             float* d_lat, float* d_query_emb, int B, int Q, int L, int C,
                             void* stream) {
  const int QS = Q >= 64 ? 8 : 1;   // query-axis splits of the latent-gradient reduction
  int64_t blocks = (int64_t)B * L * QS + (int64_t)B * Q;
  if (blocks == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t me = cudaMemsetAsync(d_lat, 0, (size_t)B * L * C * sizeof(float), st);
  SPA3D_REQUIRE(me == cudaSuccess, "decoder_tokens_bwd: memset: %s", cudaGetErrorString(me));
  SPA3D_DISPATCH(tok_dtype, TT, {
    decoder_tokens_bwd_kernel<TT><<<(unsigned)blocks, 256, 0, st>>>((const TT*)d_tokens, query_frame, d_lat, d_query_emb, B, Q, L, C, QS);
  });
  return check_launch("decoder_tokens_bwd");
}

int spa3d_split_outputs(const float* head_out, float* tracks, float* visible_logits, int64_t rows,
                        int T, int coords, float* certain_logits, void* stream) {
  if (rows == 0) return 0;
  SPA3D_REQUIRE(coords == 2 || coords == 3, "split_outputs: coords must be 2 or 3");
  split_outputs_kernel<<<blocks_for(rows * T, 256), 256, 0, (cudaStream_t)stream>>>(head_out, tracks, visible_logits, certain_logits, rows, T, coords);
  return check_launch("split_outputs");
}

int spa3d_to_tapvid3d(const float* tracks, const float* visible_logits, const float* target_tracks, float* out_tracks,
                      uint8_t* out_occluded, float* out_score, int64_t Q, int T, int coords, void* stream) {
  if (Q == 0 || T == 0) return 0;
  SPA3D_REQUIRE(tracks && visible_logits && out_tracks && out_occluded, "to_tapvid3d: NULL operand");
  SPA3D_REQUIRE(coords >= 1 && coords <= 4, "to_tapvid3d: coords must be 1..4");
  SPA3D_REQUIRE((out_score == nullptr) || (target_tracks != nullptr), "to_tapvid3d: a score needs target tracks");
  spa3d::to_tapvid3d_kernel<<<blocks_for(Q * T, 256), 256, 0, (cudaStream_t)stream>>>(tracks, visible_logits, target_tracks, out_tracks,
                                                                                     out_occluded, out_score, Q, T, coords);
  return check_launch("to_tapvid3d");
}

int spa3d_loss_fwd(const float* head_out, const float* target_tracks, const float* target_vis,
                   float* sums, int64_t rows, int T, void* stream) {
  if (rows == 0) return 0;
  int64_t total = rows * T;
  unsigned nb = blocks_for(total, 256);
  unsigned cap = (unsigned)num_sms() * 8;
  if (nb > cap) nb = cap;
  loss_fwd_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(head_out, target_tracks, target_vis, sums, rows, T);
  return check_launch("loss_fwd");
}

int spa3d_loss_bwd(const float* head_out, const float* target_tracks, const float* target_vis,
                   float* d_head_out, float l1_w, float bce_w, float inv_denom, int64_t rows, int T,
                   void* stream) {
  if (rows == 0) return 0;
  loss_bwd_kernel<<<blocks_for(rows * T, 256), 256, 0, (cudaStream_t)stream>>>(head_out, target_tracks, target_vis, d_head_out, l1_w, bce_w, inv_denom, rows, T);
  return check_launch("loss_bwd");
}

int spa3d_gelu_bwd(const void* pre, int64_t ldp, int p_dtype, const void* dy, int64_t lddy,
                   int dy_dtype, void* dx, int64_t lddx, int dx_dtype, int64_t rows, int cols,
                   void* stream) {
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(p_dtype, TP, SPA3D_DISPATCH(dy_dtype, TDY, SPA3D_DISPATCH(dx_dtype, TDX, {
    gelu_bwd_kernel<TP, TDY, TDX><<<blocks_for(rows * cols, 256), 256, 0, st>>>((const TP*)pre, ldp, (const TDY*)dy, lddy, (TDX*)dx, lddx, rows, cols);
  })));
  return check_launch("gelu_bwd");
}

int spa3d_colsum(const void* x, int64_t ldx, int dtype, float* out, int accumulate, int64_t rows,
                 int cols, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) zero_kernel<<<blocks_for(cols, 256), 256, 0, st>>>(out, cols);
  if (rows == 0) return check_launch("colsum");
  int64_t slabs = (rows + 1023) / 1024;
  if (slabs > 4096) slabs = 4096;
  int64_t rpb = (rows + slabs - 1) / slabs;
  const int vec = dtype == SPA3D_F32 ? 4 : 8;
  if (cols % vec == 0 && ldx % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    dim3 gridv((cols / vec + 31) / 32, (unsigned)slabs);
    SPA3D_DISPATCH(dtype, T, {
      colsum_vec_kernel<T><<<gridv, 256, 0, st>>>((const T*)x, ldx, out, rows, cols, rpb);
    });
    return check_launch("colsum_vec");
  }
  dim3 grid((cols + 31) / 32, (unsigned)slabs);
  SPA3D_DISPATCH(dtype, T, {
    colsum_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ldx, out, rows, cols, rpb);
  });
  return check_launch("colsum");
}

int spa3d_shadow_weights(const float* src, int64_t lds, void* dst, int64_t ldd, void* dst_t, int64_t ldt, int dtype,
                         int rows, int cols, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  SPA3D_REQUIRE(src != nullptr && (dst != nullptr || dst_t != nullptr), "shadow_weights: NULL operand");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  SPA3D_DISPATCH(dtype, T, {
    shadow_weights_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, (T*)dst, ldd, (T*)dst_t, ldt, rows, cols);
  });
  return check_launch("shadow_weights");
}

int spa3d_split3(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int K, void* stream) {
  if (rows == 0 || K == 0) return 0;
  SPA3D_REQUIRE(K % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                "split3: K and the row pitches must be multiples of 4, src 16-byte and dst 8-byte aligned");
  split3_kernel<<<blocks_for(rows * (K / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, rows, K / 4);
  return check_launch("split3");
}

int spa3d_fill_zero(void* p, int64_t bytes, void* stream) {
  if (bytes == 0) return 0;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  SPA3D_REQUIRE(e == cudaSuccess, "fill_zero: %s", cudaGetErrorString(e));
  return 0;
}

int spa3d_axpy(float* y, const float* x, float alpha, int64_t n, void* stream) {
  if (n == 0) return 0;
  axpy_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y, x, alpha, n);
  return check_launch("axpy");
}

int spa3d_sumsq(const float* g, int64_t n, float* sumsq, float* workspace, void* stream) {
  if (n == 0) return 0;
  SPA3D_REQUIRE(workspace != nullptr, "sumsq: workspace of SPA3D_SUMSQ_WORKSPACE floats required");
  unsigned nb = blocks_for(n, 1024);
  if (nb > SPA3D_SUMSQ_WORKSPACE) nb = SPA3D_SUMSQ_WORKSPACE;
  sumsq_partial_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(g, n, workspace);
  sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, (int)nb, sumsq);
  return check_launch("sumsq");
}

int spa3d_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                     float clip_norm, float lr, float b1, float b2, float eps, float wd, int step,
                     void* stream) {
  if (n == 0) return 0;
  float bc1 = (float)(1.0 - pow((double)b1, (double)step)), bc2 = (float)(1.0 - pow((double)b2, (double)step));
  adamw_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sumsq, clip_norm, lr, b1, b2, eps, wd, bc1, bc2);
  return check_launch("adamw_step");
}

}  // extern "C"
