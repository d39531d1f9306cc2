// Attention with ONE query per sequence (bf16): the pruned last layer of the per-track and read-out
// transformers - only token 0 is read downstream (track_autoencoder_3d.py:186-188, :286-287), so its
// query attends to all T+1 (or 129) keys while no other query of that layer is ever needed.
//
// The work is a bandwidth problem (K and V of every sequence are read once, 2*Lk*Dh flops per byte pair):
// one warp per (sequence, head), 4 lanes per key row (16-byte loads, each lane owns Dh/4 dimensions), 8 keys
// per warp instruction, scores of up to 160 keys in registers.  No shared memory, no tensor cores.
//   forward : s_j = q.k_j (+ key mask), p = softmax(s), o = sum_j p_j v_j; saves (max, 1/sum) like the other kernels
//   backward: recomputes p; dP_j = dO.v_j, delta = sum_j p_j dP_j, dS_j = p_j (dP_j - delta);
//             dq = sum_j dS_j k_j, dK_j = dS_j q, dV_j = p_j dO  (K is re-read for dq: an L2 hit)
#include "common.cuh"

namespace spa3d {
namespace aq1 {

constexpr int MAXI = 20;   // keys per lane group: Lk <= 160

template <int N>
__device__ __forceinline__ void load_chunk(const bf16* p, float (&f)[N]) {
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + c * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[c * 8 + 2 * e] = __uint_as_float(w[e] << 16);
      f[c * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
}

template <int N>
__device__ __forceinline__ void store_chunk(bf16* p, const float (&f)[N], float mul) {
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[c * 8 + 2 * e] * mul, f[c * 8 + 2 * e + 1] * mul);
      w[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ float group_sum4(float x) {   // over the 4 lanes of a key group
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}
__device__ __forceinline__ float across_groups_sum(float x) {   // over the 8 key groups (same dimension slice)
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 8);
  x += __shfl_xor_sync(0xffffffffu, x, 16);
  return x;
}
__device__ __forceinline__ float across_groups_max(float x) {
  x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
  x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
  x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
  return x;
}

template <int DH, bool BWD>
__global__ void __launch_bounds__(128)
attn_q1_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
               const bf16* __restrict__ v, int64_t ldv, bf16* __restrict__ o, int64_t ldo,
               const bf16* __restrict__ d_o, int64_t lddo, bf16* __restrict__ dq, int64_t lddq,
               bf16* __restrict__ dk, int64_t lddk, bf16* __restrict__ dv, int64_t lddv,
               const uint8_t* __restrict__ mask, float* __restrict__ stats, int heads, int Lk, int64_t items) {
  constexpr int DL = DH / 4;   // dimensions per lane
  const int64_t item = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (item >= items) return;
  const int lane = threadIdx.x & 31, sub = lane & 3, kg = lane >> 2;
  const int64_t b = item / heads;
  const int h = (int)(item % heads);
  const int64_t col = (int64_t)h * DH + sub * DL;
  const bf16* kb = k + b * Lk * ldk + col;
  const bf16* vb = v + b * Lk * ldv + col;
  const uint8_t* mb = mask ? mask + b * Lk : nullptr;

  float qf[DL];
  load_chunk<DL>(q + b * ldq + col, qf);
  float dof[BWD ? DL : 1];
  if constexpr (BWD) load_chunk<DL>(d_o + b * lddo + col, dof);

  float sc[MAXI], dp[BWD ? MAXI : 1];
  float acc[DL];   // forward: o partial
#pragma unroll
  for (int d = 0; d < DL; ++d) acc[d] = 0.f;
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = kg + 8 * i;
    sc[i] = -INFINITY;
    if (BWD) dp[i] = 0.f;
    if (j < Lk) {
      float kf[DL];
      load_chunk<DL>(kb + (int64_t)j * ldk, kf);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DL; ++d) s = fmaf(qf[d], kf[d], s);
      if constexpr (BWD) {
        float vf[DL];
        load_chunk<DL>(vb + (int64_t)j * ldv, vf);
        float t = 0.f;
#pragma unroll
        for (int d = 0; d < DL; ++d) t = fmaf(dof[d], vf[d], t);
        dp[i] = t;
      }
      sc[i] = s;
    }
  }
  // the group reductions are outside the divergent branch (all 32 lanes take part)
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = kg + 8 * i;
    float s = group_sum4(j < Lk ? sc[i] : 0.f);
    if (BWD) dp[i] = group_sum4(dp[i]);
    if (j < Lk) {
      if (mb != nullptr && mb[j] == 0) s = masked_logit_bf16();
      sc[i] = s;
      m = fmaxf(m, s);
    }
  }
  m = across_groups_max(m);
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const float p = (kg + 8 * i < Lk) ? __expf(sc[i] - m) : 0.f;
    sc[i] = p;
    l += p;
  }
  l = across_groups_sum(l);
  const float inv = 1.f / l;

  if constexpr (!BWD) {
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int j = kg + 8 * i;
      if (j < Lk) {
        float vf[DL];
        load_chunk<DL>(vb + (int64_t)j * ldv, vf);
        const float p = sc[i];
#pragma unroll
        for (int d = 0; d < DL; ++d) acc[d] = fmaf(p, vf[d], acc[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < DL; ++d) acc[d] = across_groups_sum(acc[d]);
    if (kg == 0) store_chunk<DL>(o + b * ldo + col, acc, inv);
    if (stats != nullptr && lane == 0) {
      stats[item * 2] = m;
      stats[item * 2 + 1] = inv;
    }
  } else {
    float delta = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      sc[i] *= inv;                 // p_j
      delta = fmaf(sc[i], dp[i], delta);
    }
    delta = across_groups_sum(delta);
    bf16* dkb = dk + b * Lk * lddk + col;
    bf16* dvb = dv + b * Lk * lddv + col;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int j = kg + 8 * i;
      if (j < Lk) {
        const float p = sc[i];
        const bool masked = mb != nullptr && mb[j] == 0;   // a masked logit is a constant: no gradient through it
        const float ds = masked ? 0.f : p * (dp[i] - delta);
        float kf[DL];
        load_chunk<DL>(kb + (int64_t)j * ldk, kf);
#pragma unroll
        for (int d = 0; d < DL; ++d) acc[d] = fmaf(ds, kf[d], acc[d]);
        store_chunk<DL>(dkb + (int64_t)j * lddk, qf, ds);
        store_chunk<DL>(dvb + (int64_t)j * lddv, dof, p);
      }
    }
#pragma unroll
    for (int d = 0; d < DL; ++d) acc[d] = across_groups_sum(acc[d]);
    if (kg == 0) store_chunk<DL>(dq + b * lddq + col, acc, 1.f);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace aq1

bool attention_q1_applicable(int dtype, int Lq, int Lk, int Dh) {
  static const bool enabled = [] { const char* e = getenv("SPA3D_ATTN_Q1"); return !(e && atoi(e) == 0); }();
  return enabled && dtype == SPA3D_BF16 && Lq == 1 && Lk >= 1 && Lk <= 8 * aq1::MAXI && (Dh == 64 || Dh == 96);
}

int attention_q1_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int Lk, int Dh,
                     cudaStream_t st) {
  using namespace aq1;
  SPA3D_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                "attention_q1: operands must be 16-byte aligned");
  const int64_t items = batch * heads;
  const unsigned grid = (unsigned)((items + 3) / 4);
  if (Dh == 96)
    attn_q1_kernel<96, false><<<grid, 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, nullptr, 0,
                                                   nullptr, 0, nullptr, 0, nullptr, 0, key_mask, stats, heads, Lk, items);
  else
    attn_q1_kernel<64, false><<<grid, 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, nullptr, 0,
                                                   nullptr, 0, nullptr, 0, nullptr, 0, key_mask, stats, heads, Lk, items);
  return check_launch("attention_q1_fwd");
}

int attention_q1_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, int64_t batch, int heads, int Lk, int Dh, cudaStream_t st) {
  using namespace aq1;
  SPA3D_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(d_o) && aligned16(dq) && aligned16(dk) && aligned16(dv) &&
                    ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                "attention_q1: operands must be 16-byte aligned");
  const int64_t items = batch * heads;
  const unsigned grid = (unsigned)((items + 3) / 4);
  if (Dh == 96)
    attn_q1_kernel<96, true><<<grid, 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, nullptr, 0, (const bf16*)d_o,
                                                  lddo, (bf16*)dq, lddq, (bf16*)dk, lddk, (bf16*)dv, lddv, key_mask, nullptr, heads, Lk, items);
  else
    attn_q1_kernel<64, true><<<grid, 128, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, nullptr, 0, (const bf16*)d_o,
                                                  lddo, (bf16*)dq, lddq, (bf16*)dk, lddk, (bf16*)dv, lddv, key_mask, nullptr, heads, Lk, items);
  return check_launch("attention_q1_bwd");
}

}  // namespace spa3d
