// softmax(q k^T [+ key mask]) v and its backward - fp32 SIMT kernels with online softmax.
//
// These are the accurate-mode ("fp32-accumulate") attention cores and the generic fallback for
// shapes the tensor-core kernels do not take (e.g. the 128 x N latents<-tracks cross-attention,
// track_autoencoder_3d.py:200-201, 0.2 % of FLOPs).  Semantics follow flax
// nn.dot_product_attention as used at attention.py:175: masked logits are replaced by
// finfo(float32).min (NOT -inf), so a fully masked row gets uniform weights.
//
// Layout: q/k/v/o are token-major matrices; head h occupies columns [h*Dh, (h+1)*Dh).
// q is expected to be RMS-normalised and already divided by sqrt(Dh).
#include "common.cuh"

namespace spa3d {

constexpr int AT = 32;       // tile of query rows / keys per block iteration
constexpr int AMAXI = 4;     // Dh <= 128 : up to 4 channels per lane

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_simt(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
              const T* __restrict__ v, int64_t ldv, T* __restrict__ o, int64_t ldo,
              const uint8_t* __restrict__ mask, float* __restrict__ lse, int heads, int Lq, int Lk,
              int Dh, int nqb) {
  extern __shared__ float sm[];
  float* Qs = sm;                     // [Dh][AT]   (transposed: row index fastest)
  float* Ks = Qs + Dh * AT;           // [AT][Dh+1]
  float* Vs = Ks + AT * (Dh + 1);     // [AT][Dh+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qb = blockIdx.x % nqb;
  const int64_t bh = blockIdx.x / nqb;
  const int h = (int)(bh % heads);
  const int64_t b = bh / heads;
  const int q0 = qb * AT;
  const int D1 = Dh + 1;
  const int ni = (Dh + 31) >> 5;

  for (int idx = tid; idx < AT * Dh; idx += 256) {
    int r = idx / Dh, d = idx % Dh;
    int qr = q0 + r;
    Qs[d * AT + r] = qr < Lq ? ldf<T>(q + (b * Lq + qr) * ldq + (int64_t)h * Dh + d) : 0.f;
  }
  float m[4], l[4], acc[4][AMAXI];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) acc[r][i] = 0.f;
  }

  for (int k0 = 0; k0 < Lk; k0 += AT) {
    __syncthreads();
    for (int idx = tid; idx < AT * Dh; idx += 256) {
      int j = idx / Dh, d = idx % Dh;
      int kj = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < Lk) {
        kv = ldf<T>(k + (b * Lk + kj) * ldk + (int64_t)h * Dh + d);
        vv = ldf<T>(v + (b * Lk + kj) * ldv + (int64_t)h * Dh + d);
      }
      Ks[j * D1 + d] = kv;
      Vs[j * D1 + d] = vv;
    }
    __syncthreads();
    const int kj = k0 + lane;
    const bool valid = kj < Lk;
    const bool keep = valid && (mask == nullptr || mask[b * Lk + kj] != 0);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < Dh; ++d) {
      float kd = Ks[lane * D1 + d];
      float4 q4 = *reinterpret_cast<const float4*>(&Qs[d * AT + warp * 4]);
      s[0] = fmaf(q4.x, kd, s[0]);
      s[1] = fmaf(q4.y, kd, s[1]);
      s[2] = fmaf(q4.z, kd, s[2]);
      s[3] = fmaf(q4.w, kd, s[3]);
    }
    float p[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float sr = valid ? (keep ? s[r] : -FLT_MAX) : -INFINITY;
      float mn = fmaxf(m[r], warp_max(sr));
      float corr = expf(m[r] - mn);  // m = -inf on the first tile -> 0
      p[r] = valid ? expf(sr - mn) : 0.f;
      l[r] = l[r] * corr + warp_sum(p[r]);
      m[r] = mn;
#pragma unroll
      for (int i = 0; i < AMAXI; ++i) acc[r][i] *= corr;
    }
    for (int j = 0; j < AT; ++j) {
      float p0 = __shfl_sync(0xffffffffu, p[0], j), p1 = __shfl_sync(0xffffffffu, p[1], j);
      float p2 = __shfl_sync(0xffffffffu, p[2], j), p3 = __shfl_sync(0xffffffffu, p[3], j);
#pragma unroll
      for (int i = 0; i < AMAXI; ++i) {
        if (i < ni) {
          int d = lane + 32 * i;
          float vv = d < Dh ? Vs[j * D1 + d] : 0.f;
          acc[0][i] = fmaf(p0, vv, acc[0][i]);
          acc[1][i] = fmaf(p1, vv, acc[1][i]);
          acc[2][i] = fmaf(p2, vv, acc[2][i]);
          acc[3][i] = fmaf(p3, vv, acc[3][i]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int qr = q0 + warp * 4 + r;
    if (qr >= Lq) continue;
    float inv = 1.f / l[r];
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) {
      int d = lane + 32 * i;
      if (i < ni && d < Dh) stf<T>(o + (b * Lq + qr) * ldo + (int64_t)h * Dh + d, acc[r][i] * inv);
    }
    if (lse != nullptr && lane == 0) {  // softmax stats for the backward: (row max, 1/row sum)
      lse[((b * heads + h) * Lq + qr) * 2] = m[r];
      lse[((b * heads + h) * Lq + qr) * 2 + 1] = inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: delta = rowsum(dO * O); dQ kernel (per q-tile); dK/dV kernel (per key-tile)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ o, int64_t ldo, const T* __restrict__ d_o,
                                  int64_t lddo, float* __restrict__ delta, int64_t rows, int heads,
                                  int Lq, int Dh) {
  // one warp per (token row, head); delta layout [batch, heads, Lq]
  int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= rows * heads) return;
  int lane = threadIdx.x & 31;
  int64_t r = wid / heads;
  int h = (int)(wid % heads);
  float s = 0.f;
  for (int d = lane; d < Dh; d += 32)
    s += ldf<T>(o + r * ldo + (int64_t)h * Dh + d) * ldf<T>(d_o + r * lddo + (int64_t)h * Dh + d);
  s = warp_sum(s);
  if (lane == 0) {
    int64_t b = r / Lq;
    int qi = (int)(r % Lq);
    delta[(b * heads + h) * Lq + qi] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_dq_simt(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                 const T* __restrict__ v, int64_t ldv, const T* __restrict__ d_o, int64_t lddo,
                 T* __restrict__ dq, int64_t lddq, const uint8_t* __restrict__ mask,
                 const float* __restrict__ lse, const float* __restrict__ delta, int heads, int Lq,
                 int Lk, int Dh, int nqb) {
  extern __shared__ float sm[];
  float* Qs = sm;                      // [Dh][AT]
  float* dOs = Qs + Dh * AT;           // [Dh][AT]
  float* Ks = dOs + Dh * AT;           // [AT][Dh+1]
  float* Vs = Ks + AT * (Dh + 1);      // [AT][Dh+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qb = blockIdx.x % nqb;
  const int64_t bh = blockIdx.x / nqb;
  const int h = (int)(bh % heads);
  const int64_t b = bh / heads;
  const int q0 = qb * AT, D1 = Dh + 1, ni = (Dh + 31) >> 5;
  for (int idx = tid; idx < AT * Dh; idx += 256) {
    int r = idx / Dh, d = idx % Dh;
    int qr = q0 + r;
    bool ok = qr < Lq;
    Qs[d * AT + r] = ok ? ldf<T>(q + (b * Lq + qr) * ldq + (int64_t)h * Dh + d) : 0.f;
    dOs[d * AT + r] = ok ? ldf<T>(d_o + (b * Lq + qr) * lddo + (int64_t)h * Dh + d) : 0.f;
  }
  float lrow[4], irow[4], drow[4], acc[4][AMAXI];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int qr = q0 + warp * 4 + r;
    lrow[r] = qr < Lq ? lse[((b * heads + h) * Lq + qr) * 2] : 0.f;
    irow[r] = qr < Lq ? lse[((b * heads + h) * Lq + qr) * 2 + 1] : 0.f;
    drow[r] = qr < Lq ? delta[(b * heads + h) * Lq + qr] : 0.f;
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) acc[r][i] = 0.f;
  }
  for (int k0 = 0; k0 < Lk; k0 += AT) {
    __syncthreads();
    for (int idx = tid; idx < AT * Dh; idx += 256) {
      int j = idx / Dh, d = idx % Dh;
      int kj = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < Lk) {
        kv = ldf<T>(k + (b * Lk + kj) * ldk + (int64_t)h * Dh + d);
        vv = ldf<T>(v + (b * Lk + kj) * ldv + (int64_t)h * Dh + d);
      }
      Ks[j * D1 + d] = kv;
      Vs[j * D1 + d] = vv;
    }
    __syncthreads();
    const int kj = k0 + lane;
    const bool valid = kj < Lk;
    const bool keep = valid && (mask == nullptr || mask[b * Lk + kj] != 0);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < Dh; ++d) {
      float kd = Ks[lane * D1 + d], vd = Vs[lane * D1 + d];
      float4 q4 = *reinterpret_cast<const float4*>(&Qs[d * AT + warp * 4]);
      float4 g4 = *reinterpret_cast<const float4*>(&dOs[d * AT + warp * 4]);
      s[0] = fmaf(q4.x, kd, s[0]); s[1] = fmaf(q4.y, kd, s[1]);
      s[2] = fmaf(q4.z, kd, s[2]); s[3] = fmaf(q4.w, kd, s[3]);
      dp[0] = fmaf(g4.x, vd, dp[0]); dp[1] = fmaf(g4.y, vd, dp[1]);
      dp[2] = fmaf(g4.z, vd, dp[2]); dp[3] = fmaf(g4.w, vd, dp[3]);
    }
    float ds[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float pr = keep ? expf(s[r] - lrow[r]) * irow[r] : 0.f;  // masked logits are constants: no gradient
      ds[r] = pr * (dp[r] - drow[r]);
    }
    for (int j = 0; j < AT; ++j) {
      float d0 = __shfl_sync(0xffffffffu, ds[0], j), d1 = __shfl_sync(0xffffffffu, ds[1], j);
      float d2 = __shfl_sync(0xffffffffu, ds[2], j), d3 = __shfl_sync(0xffffffffu, ds[3], j);
#pragma unroll
      for (int i = 0; i < AMAXI; ++i) {
        if (i < ni) {
          int d = lane + 32 * i;
          float kk = d < Dh ? Ks[j * D1 + d] : 0.f;
          acc[0][i] = fmaf(d0, kk, acc[0][i]); acc[1][i] = fmaf(d1, kk, acc[1][i]);
          acc[2][i] = fmaf(d2, kk, acc[2][i]); acc[3][i] = fmaf(d3, kk, acc[3][i]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int qr = q0 + warp * 4 + r;
    if (qr >= Lq) continue;
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) {
      int d = lane + 32 * i;
      if (i < ni && d < Dh) stf<T>(dq + (b * Lq + qr) * lddq + (int64_t)h * Dh + d, acc[r][i]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_dkv_simt(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                  const T* __restrict__ v, int64_t ldv, const T* __restrict__ d_o, int64_t lddo,
                  T* __restrict__ dk, int64_t lddk, T* __restrict__ dv, int64_t lddv,
                  const uint8_t* __restrict__ mask, const float* __restrict__ lse,
                  const float* __restrict__ delta, int heads, int Lq, int Lk, int Dh, int nkb) {
  extern __shared__ float sm[];
  float* Ks = sm;                      // [Dh][AT]  (keys of this block, transposed)
  float* Vs = Ks + Dh * AT;            // [Dh][AT]
  float* Qs = Vs + Dh * AT;            // [AT][Dh+1] (query tile)
  float* dOs = Qs + AT * (Dh + 1);     // [AT][Dh+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kb = blockIdx.x % nkb;
  const int64_t bh = blockIdx.x / nkb;
  const int h = (int)(bh % heads);
  const int64_t b = bh / heads;
  const int k0 = kb * AT, D1 = Dh + 1, ni = (Dh + 31) >> 5;
  for (int idx = tid; idx < AT * Dh; idx += 256) {
    int j = idx / Dh, d = idx % Dh;
    int kj = k0 + j;
    bool ok = kj < Lk;
    Ks[d * AT + j] = ok ? ldf<T>(k + (b * Lk + kj) * ldk + (int64_t)h * Dh + d) : 0.f;
    Vs[d * AT + j] = ok ? ldf<T>(v + (b * Lk + kj) * ldv + (int64_t)h * Dh + d) : 0.f;
  }
  bool keep[4];
  float accK[4][AMAXI], accV[4][AMAXI];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int kj = k0 + warp * 4 + r;
    keep[r] = kj < Lk && (mask == nullptr || mask[b * Lk + kj] != 0);
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) accK[r][i] = accV[r][i] = 0.f;
  }
  for (int q0 = 0; q0 < Lq; q0 += AT) {
    __syncthreads();
    for (int idx = tid; idx < AT * Dh; idx += 256) {
      int r = idx / Dh, d = idx % Dh;
      int qr = q0 + r;
      bool ok = qr < Lq;
      Qs[r * D1 + d] = ok ? ldf<T>(q + (b * Lq + qr) * ldq + (int64_t)h * Dh + d) : 0.f;
      dOs[r * D1 + d] = ok ? ldf<T>(d_o + (b * Lq + qr) * lddo + (int64_t)h * Dh + d) : 0.f;
    }
    __syncthreads();
    const int qi = q0 + lane;
    const bool qvalid = qi < Lq;
    const float lq = qvalid ? lse[((b * heads + h) * Lq + qi) * 2] : 0.f;
    const float il = qvalid ? lse[((b * heads + h) * Lq + qi) * 2 + 1] : 0.f;
    const float dl = qvalid ? delta[(b * heads + h) * Lq + qi] : 0.f;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < Dh; ++d) {
      float qd = Qs[lane * D1 + d], gd = dOs[lane * D1 + d];
      float4 k4 = *reinterpret_cast<const float4*>(&Ks[d * AT + warp * 4]);
      float4 v4 = *reinterpret_cast<const float4*>(&Vs[d * AT + warp * 4]);
      s[0] = fmaf(k4.x, qd, s[0]); s[1] = fmaf(k4.y, qd, s[1]);
      s[2] = fmaf(k4.z, qd, s[2]); s[3] = fmaf(k4.w, qd, s[3]);
      dp[0] = fmaf(v4.x, gd, dp[0]); dp[1] = fmaf(v4.y, gd, dp[1]);
      dp[2] = fmaf(v4.z, gd, dp[2]); dp[3] = fmaf(v4.w, gd, dp[3]);
    }
    float p[4], ds[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int kj = k0 + warp * 4 + r;
      // probability as in the forward: masked logits are -FLT_MAX (non-zero weight only when the
      // whole row is masked); the value path always gets p, the logit path only if kept.
      float sr = keep[r] ? s[r] : -FLT_MAX;
      p[r] = (qvalid && kj < Lk) ? expf(sr - lq) * il : 0.f;
      ds[r] = keep[r] ? p[r] * (dp[r] - dl) : 0.f;
    }
    for (int i2 = 0; i2 < AT; ++i2) {
      float p0 = __shfl_sync(0xffffffffu, p[0], i2), p1 = __shfl_sync(0xffffffffu, p[1], i2);
      float p2 = __shfl_sync(0xffffffffu, p[2], i2), p3 = __shfl_sync(0xffffffffu, p[3], i2);
      float d0 = __shfl_sync(0xffffffffu, ds[0], i2), d1 = __shfl_sync(0xffffffffu, ds[1], i2);
      float d2 = __shfl_sync(0xffffffffu, ds[2], i2), d3 = __shfl_sync(0xffffffffu, ds[3], i2);
#pragma unroll
      for (int i = 0; i < AMAXI; ++i) {
        if (i < ni) {
          int d = lane + 32 * i;
          float g = d < Dh ? dOs[i2 * D1 + d] : 0.f;
          float qq = d < Dh ? Qs[i2 * D1 + d] : 0.f;
          accV[0][i] = fmaf(p0, g, accV[0][i]); accV[1][i] = fmaf(p1, g, accV[1][i]);
          accV[2][i] = fmaf(p2, g, accV[2][i]); accV[3][i] = fmaf(p3, g, accV[3][i]);
          accK[0][i] = fmaf(d0, qq, accK[0][i]); accK[1][i] = fmaf(d1, qq, accK[1][i]);
          accK[2][i] = fmaf(d2, qq, accK[2][i]); accK[3][i] = fmaf(d3, qq, accK[3][i]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int kj = k0 + warp * 4 + r;
    if (kj >= Lk) continue;
#pragma unroll
    for (int i = 0; i < AMAXI; ++i) {
      int d = lane + 32 * i;
      if (i < ni && d < Dh) {
        stf<T>(dk + (b * Lk + kj) * lddk + (int64_t)h * Dh + d, accK[r][i]);
        stf<T>(dv + (b * Lk + kj) * lddv + (int64_t)h * Dh + d, accV[r][i]);
      }
    }
  }
}

int attention_fwd_simt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                       float* lse_out, int64_t batch, int heads, int Lq, int Lk, int Dh,
                       cudaStream_t st) {
  SPA3D_REQUIRE(Dh > 0 && Dh <= 128, "attention: head dim %d not in 1..128", Dh);
  int nqb = (Lq + AT - 1) / AT;
  int64_t blocks = batch * heads * nqb;
  SPA3D_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  size_t smem = sizeof(float) * (Dh * AT + 2 * AT * (Dh + 1));
  SPA3D_DISPATCH(dtype, T, {
    attn_fwd_simt<T><<<(unsigned)blocks, 256, smem, st>>>((const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (T*)o, ldo, key_mask, lse_out, heads, Lq, Lk, Dh, nqb);
  });
  return check_launch("attention_fwd_simt");
}

int attention_delta(const void* o, int64_t ldo, const void* d_o, int64_t lddo, int dtype, float* delta_ws,
                    int64_t batch, int heads, int Lq, int Dh, cudaStream_t st) {
  int64_t rows = batch * Lq;
  SPA3D_DISPATCH(dtype, T, {
    attn_delta_kernel<T><<<(unsigned)((rows * heads + 7) / 8), 256, 0, st>>>((const T*)o, ldo, (const T*)d_o, lddo, delta_ws, rows, heads, Lq, Dh);
  });
  return check_launch("attention_delta");
}

int attention_bwd_simt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                       void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                       int dtype, const uint8_t* key_mask, const float* lse, float* delta_ws,
                       int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st) {
  SPA3D_REQUIRE(Dh > 0 && Dh <= 128, "attention: head dim %d not in 1..128", Dh);
  int nqb = (Lq + AT - 1) / AT, nkb = (Lk + AT - 1) / AT;
  int64_t rows = batch * Lq;
  size_t smem_dq = sizeof(float) * (2 * Dh * AT + 2 * AT * (Dh + 1));
  SPA3D_DISPATCH(dtype, T, {
    attn_delta_kernel<T><<<(unsigned)((rows * heads + 7) / 8), 256, 0, st>>>((const T*)o, ldo, (const T*)d_o, lddo, delta_ws, rows, heads, Lq, Dh);
    cudaFuncSetAttribute(attn_bwd_dq_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);
    cudaFuncSetAttribute(attn_bwd_dkv_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);
    attn_bwd_dq_simt<T><<<(unsigned)(batch * heads * nqb), 256, smem_dq, st>>>((const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (const T*)d_o, lddo, (T*)dq, lddq, key_mask, lse, delta_ws, heads, Lq, Lk, Dh, nqb);
    attn_bwd_dkv_simt<T><<<(unsigned)(batch * heads * nkb), 256, smem_dq, st>>>((const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (const T*)d_o, lddo, (T*)dk, lddk, (T*)dv, lddv, key_mask, lse, delta_ws, heads, Lq, Lk, Dh, nkb);
  });
  return check_launch("attention_bwd_simt");
}

}  // namespace spa3d
