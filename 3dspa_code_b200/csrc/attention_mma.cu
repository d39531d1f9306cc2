// Short-sequence attention core on tensor cores (bf16 in, fp32 softmax/accumulate):
//   o = softmax(q k^T [+ key mask]) v   per (sequence, head), whole sequence in one CTA.
//
// Used for the per-track temporal self-attention (L = T+1 = 151, track_autoencoder_3d.py:182-184),
// the decoder read-out attention (L = 129, :285) and the latent self-attentions (L = 128).
// One CTA owns one (sequence, head): Q, K, V ([L, Dh] each) are staged once in shared memory with
// cp.async (rows padded by 16 B so ldmatrix is bank-conflict free), each warp owns 16 query rows
// and walks the keys in tiles of 32 with an online softmax, so the L x L score matrix never
// leaves registers.  S = Q K^T and O = P V run on mma.sync.m16n8k16 (bf16 -> fp32).
//
// Why not tcgen05 here: L = 151/129 rows would pad to 256 (or 192 with M=64 tiles) in a UMMA
// tile, the core is 2-9 % of the layer FLOPs, and at Dh = 96 it is bound by the HBM traffic of
// q/k/v/o, not by the tensor pipe.  The dense projections around it use tcgen05.
//
// Mask semantics (attention.py:175, flax dot_product_attention): masked logits become one huge
// negative constant (finfo.min there, bf16(-1e30) here, see common.cuh), so an all-masked row is
// uniform; keys beyond Lk do not exist (-inf).
#include <cuda_pipeline.h>

#include "common.cuh"

namespace spa3d {

namespace am {

constexpr int KT = 32;  // keys per online-softmax step

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int DH>
__global__ void __launch_bounds__(512)
attn_fwd_mma_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
                    const bf16* __restrict__ v, int64_t ldv, bf16* __restrict__ o, int64_t ldo,
                    const uint8_t* __restrict__ mask, float* __restrict__ stats, int heads, int Lq,
                    int Lk, int QR, int KC, int q_tiles) {
  // QR = query rows per CTA (16 per warp), KC = keys staged per chunk (multiple of KT).
  // Lk <= KC: whole sequence in one tile (self-attention, L = 151/129/128).  Lk > KC: the keys are
  // streamed chunk by chunk under the same online softmax (latents <- tracks cross-attention,
  // 128 queries x N keys, track_autoencoder_3d.py:200-201).
  constexpr int LDS = DH + 8;    // padded smem row, elements (16 B pad)
  constexpr int CH = DH / 8;     // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + (size_t)QR * LDS;
  bf16* Vs = Ks + (size_t)KC * LDS;
  uint8_t* Ms = reinterpret_cast<uint8_t*>(Vs + (size_t)KC * LDS);  // 1 keep, 0 masked, 2 absent

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int qt = blockIdx.x % q_tiles;
  const int64_t bh = blockIdx.x / q_tiles;
  const int64_t b = bh / heads;
  const int h = (int)(bh % heads);
  const int q0 = qt * QR;
  const bf16* qg = q + (b * Lq + q0) * ldq + (int64_t)h * DH;
  const bf16* kg = k + (b * Lk) * ldk + (int64_t)h * DH;
  const bf16* vg = v + (b * Lk) * ldv + (int64_t)h * DH;

  for (int idx = tid; idx < QR * CH; idx += nthr) {
    int r = idx / CH, c = idx % CH;
    bf16* dst = Qs + r * LDS + c * 8;
    if (q0 + r < Lq) __pipeline_memcpy_async(dst, qg + (int64_t)r * ldq + c * 8, 16);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }

  const int r0 = warp * 16;
  const bool active = r0 < QR && q0 + r0 < Lq;
  const int g = lane >> 2, tq = lane & 3;

  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;

  // ldmatrix source coordinates (see fragment layouts of mma.m16n8k16)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;       // A (Q), x4
  const int kb_row = (lane & 7) + (lane >> 4) * 8, kb_col = ((lane >> 3) & 1) * 8;     // B from K
  const int vb_row = (lane & 7) + ((lane >> 3) & 1) * 8, vb_col = (lane >> 4) * 8;     // B from V (.trans)
  constexpr float LOG2E = 1.4426950408889634f;

  for (int kc0 = 0; kc0 < Lk; kc0 += KC) {
    if (kc0 > 0) __syncthreads();  // every warp is done with the previous chunk
    const int kn = min(KC, (Lk - kc0 + KT - 1) / KT * KT);  // staged rows of this chunk (padded to KT)
    for (int idx = tid; idx < kn * CH; idx += nthr) {
      int r = idx / CH, c = idx % CH;
      bf16* dk = Ks + r * LDS + c * 8;
      bf16* dv = Vs + r * LDS + c * 8;
      if (kc0 + r < Lk) {
        __pipeline_memcpy_async(dk, kg + (int64_t)(kc0 + r) * ldk + c * 8, 16);
        __pipeline_memcpy_async(dv, vg + (int64_t)(kc0 + r) * ldv + c * 8, 16);
      } else {
        *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
      }
    }
    for (int j = tid; j < kn; j += nthr)
      Ms[j] = kc0 + j < Lk ? (mask == nullptr ? 1 : (mask[b * Lk + kc0 + j] != 0 ? 1 : 0)) : 2;
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    if (!active) continue;

  for (int kt = 0; kt < kn; kt += KT) {
    float s[KT / 8][4];
#pragma unroll
    for (int i = 0; i < KT / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a[4];
      ldsm_x4(a, Qs + (r0 + a_row) * LDS + kk * 16 + a_col);
#pragma unroll
      for (int n2 = 0; n2 < KT / 16; ++n2) {
        uint32_t bb[4];
        ldsm_x4(bb, Ks + (kt + n2 * 16 + kb_row) * LDS + kk * 16 + kb_col);
        mma_bf16(s[n2 * 2], a, bb[0], bb[1]);
        mma_bf16(s[n2 * 2 + 1], a, bb[2], bb[3]);
      }
    }
    // mask + online softmax (rows g and g+8 of this warp's 16)
    float tmax0 = -INFINITY, tmax1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        uint8_t mk = Ms[kt + nt * 8 + tq * 2 + e];
        float lo = mk == 1 ? s[nt][e] : (mk == 0 ? masked_logit_bf16() : -INFINITY);
        float hi = mk == 1 ? s[nt][2 + e] : (mk == 0 ? masked_logit_bf16() : -INFINITY);
        s[nt][e] = lo;
        s[nt][2 + e] = hi;
        tmax0 = fmaxf(tmax0, lo);
        tmax1 = fmaxf(tmax1, hi);
      }
    }
    tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 1));
    tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 2));
    tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 1));
    tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 2));
    const float mn0 = fmaxf(m0, tmax0), mn1 = fmaxf(m1, tmax1);
    const float c0 = exp2f((m0 - mn0) * LOG2E), c1 = exp2f((m1 - mn1) * LOG2E);
    m0 = mn0;
    m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pa[KT / 16][4];
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      float p00 = exp2f((s[nt][0] - mn0) * LOG2E), p01 = exp2f((s[nt][1] - mn0) * LOG2E);
      float p10 = exp2f((s[nt][2] - mn1) * LOG2E), p11 = exp2f((s[nt][3] - mn1) * LOG2E);
      ps0 += p00 + p01;
      ps1 += p10 + p11;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p00, p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p10, p11);
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      oacc[i][0] *= c0; oacc[i][1] *= c0;
      oacc[i][2] *= c1; oacc[i][3] *= c1;
    }
#pragma unroll
    for (int k16 = 0; k16 < KT / 16; ++k16) {
#pragma unroll
      for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
        uint32_t bb[4];
        ldsm_x4_t(bb, Vs + (kt + k16 * 16 + vb_row) * LDS + nd2 * 16 + vb_col);
        mma_bf16(oacc[nd2 * 2], pa[k16], bb[0], bb[1]);
        mma_bf16(oacc[nd2 * 2 + 1], pa[k16], bb[2], bb[3]);
      }
    }
  }
  }  // key chunks
  if (!active) return;  // no block-wide sync below this point
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;

  // stage this warp's 16 output rows through its (now dead) Q rows, then 16-byte coalesced stores
  __syncwarp();
  bf16* Os = Qs + r0 * LDS;
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    *reinterpret_cast<uint32_t*>(Os + g * LDS + i * 8 + tq * 2) = pack_bf16(oacc[i][0] * i0, oacc[i][1] * i0);
    *reinterpret_cast<uint32_t*>(Os + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(oacc[i][2] * i1, oacc[i][3] * i1);
  }
  __syncwarp();
  bf16* og = o + (b * Lq + q0) * ldo + (int64_t)h * DH;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    int r = idx / CH, c = idx % CH;
    if (q0 + r0 + r < Lq)
      *reinterpret_cast<uint4*>(og + (int64_t)(r0 + r) * ldo + c * 8) = *reinterpret_cast<const uint4*>(Os + r * LDS + c * 8);
  }
  if (stats != nullptr && tq == 0) {
    int64_t base = (b * heads + h) * Lq + q0;
    if (q0 + r0 + g < Lq) { stats[(base + r0 + g) * 2] = m0; stats[(base + r0 + g) * 2 + 1] = i0; }
    if (q0 + r0 + g + 8 < Lq) { stats[(base + r0 + g + 8) * 2] = m1; stats[(base + r0 + g + 8) * 2 + 1] = i1; }
  }
}

// ------------------------------------------------------------------------------------------
// backward: dq, dk, dv for one (sequence, head) per CTA, whole sequence in shared memory.
//   P = exp(S - m) / l (recomputed from the saved row max m and 1/l),  dP = dO V^T,
//   dS = P o (dP - delta),  dQ = dS K,  dK = dS^T Q,  dV = P^T dO,   delta_i = sum_d O_id dO_id.
// Phase 1: warp w owns query rows 16w.. and accumulates dQ.  Phase 2: warp w owns key rows 16w..,
// recomputes S^T = K Q^T and dP^T = V dO^T and accumulates dK and dV - no cross-warp reduction,
// no atomics.  Masked keys (finfo.min logits, attention.py:175) are constants: they receive no
// logit gradient but keep their (all-masked-row) probability in the value path.
// ------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(320)
attn_bwd_mma_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
                    const bf16* __restrict__ v, int64_t ldv, const bf16* __restrict__ d_o, int64_t lddo,
                    bf16* __restrict__ dq, int64_t lddq, bf16* __restrict__ dk, int64_t lddk,
                    bf16* __restrict__ dv, int64_t lddv, const uint8_t* __restrict__ mask,
                    const float* __restrict__ stats, const float* __restrict__ delta, int heads,
                    int Lq, int Lk, int LqPad, int LkPad) {
  constexpr int LDS = DH + 8;
  constexpr int CH = DH / 8;
  constexpr float LOG2E = 1.4426950408889634f;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int nwarps = blockDim.x >> 5;
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Gs = Qs + (size_t)LqPad * LDS;          // dO
  bf16* Ks = Gs + (size_t)LqPad * LDS;
  bf16* Vs = Ks + (size_t)LkPad * LDS;
  bf16* Os = Vs + (size_t)LkPad * LDS;          // per-warp output staging [nwarps][16][LDS]
  float* mS = reinterpret_cast<float*>(Os + (size_t)nwarps * 16 * LDS);   // row max
  float* iS = mS + LqPad;                                                 // 1 / row sum
  float* dS = iS + LqPad;                                                 // delta
  uint8_t* Ms = reinterpret_cast<uint8_t*>(dS + LqPad);                   // 1 keep, 0 masked, 2 absent

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int64_t b = blockIdx.x / heads;
  const int h = blockIdx.x % heads;
  const bf16* qg = q + (b * Lq) * ldq + (int64_t)h * DH;
  const bf16* gg = d_o + (b * Lq) * lddo + (int64_t)h * DH;
  const bf16* kg = k + (b * Lk) * ldk + (int64_t)h * DH;
  const bf16* vg = v + (b * Lk) * ldv + (int64_t)h * DH;

  for (int idx = tid; idx < LqPad * CH; idx += nthr) {
    int r = idx / CH, c = idx % CH;
    bf16* d0 = Qs + r * LDS + c * 8;
    bf16* d1 = Gs + r * LDS + c * 8;
    if (r < Lq) {
      __pipeline_memcpy_async(d0, qg + (int64_t)r * ldq + c * 8, 16);
      __pipeline_memcpy_async(d1, gg + (int64_t)r * lddo + c * 8, 16);
    } else {
      *reinterpret_cast<uint4*>(d0) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(d1) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int idx = tid; idx < LkPad * CH; idx += nthr) {
    int r = idx / CH, c = idx % CH;
    bf16* d0 = Ks + r * LDS + c * 8;
    bf16* d1 = Vs + r * LDS + c * 8;
    if (r < Lk) {
      __pipeline_memcpy_async(d0, kg + (int64_t)r * ldk + c * 8, 16);
      __pipeline_memcpy_async(d1, vg + (int64_t)r * ldv + c * 8, 16);
    } else {
      *reinterpret_cast<uint4*>(d0) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(d1) = make_uint4(0, 0, 0, 0);
    }
  }
  {
    const int64_t base = (b * heads + h) * Lq;
    for (int r = tid; r < LqPad; r += nthr) {
      const bool ok = r < Lq;
      mS[r] = ok ? stats[(base + r) * 2] : 0.f;
      iS[r] = ok ? stats[(base + r) * 2 + 1] : 0.f;   // 0 => every probability of a padded row is 0
      dS[r] = ok ? delta[base + r] : 0.f;
    }
    for (int j = tid; j < LkPad; j += nthr)
      Ms[j] = j < Lk ? (mask == nullptr ? 1 : (mask[b * Lk + j] != 0 ? 1 : 0)) : 2;
  }
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;       // A operand, x4
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;       // B from [n][k] rows
  const int t_row = (lane & 7) + ((lane >> 3) & 1) * 8, t_col = (lane >> 4) * 8;       // B from [k][n] rows (.trans)
  bf16* Ow = Os + (size_t)warp * 16 * LDS;

  // ---------------- phase 1: dQ (warp = 16 query rows) ----------------
  if (warp * 16 < LqPad) {
    const int r0 = warp * 16;
    const float m0 = mS[r0 + g], m1 = mS[r0 + g + 8], i0 = iS[r0 + g], i1 = iS[r0 + g + 8];
    const float e0 = dS[r0 + g], e1 = dS[r0 + g + 8];
    float acc[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int j0 = 0; j0 < LkPad; j0 += 16) {
      float s[2][4], dp[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t a[4], bb[4];
        ldsm_x4(a, Qs + (r0 + a_row) * LDS + kk * 16 + a_col);
        ldsm_x4(bb, Ks + (j0 + b_row) * LDS + kk * 16 + b_col);
        mma_bf16(s[0], a, bb[0], bb[1]);
        mma_bf16(s[1], a, bb[2], bb[3]);
        ldsm_x4(a, Gs + (r0 + a_row) * LDS + kk * 16 + a_col);
        ldsm_x4(bb, Vs + (j0 + b_row) * LDS + kk * 16 + b_col);
        mma_bf16(dp[0], a, bb[0], bb[1]);
        mma_bf16(dp[1], a, bb[2], bb[3]);
      }
      uint32_t pa[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float dsv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint8_t mk = Ms[j0 + nt * 8 + tq * 2 + e];
          const float lo = mk == 1 ? s[nt][e] : masked_logit_bf16(), hi = mk == 1 ? s[nt][2 + e] : masked_logit_bf16();
          const float p0 = mk != 2 ? exp2f((lo - m0) * LOG2E) * i0 : 0.f;
          const float p1 = mk != 2 ? exp2f((hi - m1) * LOG2E) * i1 : 0.f;
          dsv[e] = mk == 1 ? p0 * (dp[nt][e] - e0) : 0.f;
          dsv[2 + e] = mk == 1 ? p1 * (dp[nt][2 + e] - e1) : 0.f;
        }
        pa[nt * 2] = pack_bf16(dsv[0], dsv[1]);
        pa[nt * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
      }
#pragma unroll
      for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
        uint32_t bb[4];
        ldsm_x4_t(bb, Ks + (j0 + t_row) * LDS + nd2 * 16 + t_col);
        mma_bf16(acc[nd2 * 2], pa, bb[0], bb[1]);
        mma_bf16(acc[nd2 * 2 + 1], pa, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      *reinterpret_cast<uint32_t*>(Ow + g * LDS + i * 8 + tq * 2) = pack_bf16(acc[i][0], acc[i][1]);
      *reinterpret_cast<uint32_t*>(Ow + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(acc[i][2], acc[i][3]);
    }
    __syncwarp();
    bf16* og = dq + (b * Lq) * lddq + (int64_t)h * DH;
    for (int idx = lane; idx < 16 * CH; idx += 32) {
      int r = idx / CH, c = idx % CH;
      if (r0 + r < Lq)
        *reinterpret_cast<uint4*>(og + (int64_t)(r0 + r) * lddq + c * 8) = *reinterpret_cast<const uint4*>(Ow + r * LDS + c * 8);
    }
    __syncwarp();
  }

  // ---------------- phase 2: dK, dV (warp = 16 key rows) ----------------
  if (warp * 16 < LkPad) {
    const int j0 = warp * 16;
    const uint8_t mk0 = Ms[j0 + g], mk1 = Ms[j0 + g + 8];
    float ak[DH / 8][4], av[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      ak[i][0] = ak[i][1] = ak[i][2] = ak[i][3] = 0.f;
      av[i][0] = av[i][1] = av[i][2] = av[i][3] = 0.f;
    }
    for (int r0 = 0; r0 < LqPad; r0 += 16) {
      float s[2][4], dp[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t a[4], bb[4];
        ldsm_x4(a, Ks + (j0 + a_row) * LDS + kk * 16 + a_col);
        ldsm_x4(bb, Qs + (r0 + b_row) * LDS + kk * 16 + b_col);
        mma_bf16(s[0], a, bb[0], bb[1]);
        mma_bf16(s[1], a, bb[2], bb[3]);
        ldsm_x4(a, Vs + (j0 + a_row) * LDS + kk * 16 + a_col);
        ldsm_x4(bb, Gs + (r0 + b_row) * LDS + kk * 16 + b_col);
        mma_bf16(dp[0], a, bb[0], bb[1]);
        mma_bf16(dp[1], a, bb[2], bb[3]);
      }
      uint32_t pp[4], pd[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int qi = r0 + nt * 8 + tq * 2 + e;       // query index = column of S^T
          const float m = mS[qi], il = iS[qi], dl = dS[qi];
          const float lo = mk0 == 1 ? s[nt][e] : masked_logit_bf16(), hi = mk1 == 1 ? s[nt][2 + e] : masked_logit_bf16();
          pv[e] = mk0 != 2 ? exp2f((lo - m) * LOG2E) * il : 0.f;
          pv[2 + e] = mk1 != 2 ? exp2f((hi - m) * LOG2E) * il : 0.f;
          dsv[e] = mk0 == 1 ? pv[e] * (dp[nt][e] - dl) : 0.f;
          dsv[2 + e] = mk1 == 1 ? pv[2 + e] * (dp[nt][2 + e] - dl) : 0.f;
        }
        pp[nt * 2] = pack_bf16(pv[0], pv[1]);
        pp[nt * 2 + 1] = pack_bf16(pv[2], pv[3]);
        pd[nt * 2] = pack_bf16(dsv[0], dsv[1]);
        pd[nt * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
      }
#pragma unroll
      for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
        uint32_t bb[4];
        ldsm_x4_t(bb, Gs + (r0 + t_row) * LDS + nd2 * 16 + t_col);
        mma_bf16(av[nd2 * 2], pp, bb[0], bb[1]);
        mma_bf16(av[nd2 * 2 + 1], pp, bb[2], bb[3]);
        ldsm_x4_t(bb, Qs + (r0 + t_row) * LDS + nd2 * 16 + t_col);
        mma_bf16(ak[nd2 * 2], pd, bb[0], bb[1]);
        mma_bf16(ak[nd2 * 2 + 1], pd, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      bf16* outp = pass == 0 ? dk : dv;
      const int64_t ldo_ = pass == 0 ? lddk : lddv;
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        const float (&a)[4] = pass == 0 ? ak[i] : av[i];
        *reinterpret_cast<uint32_t*>(Ow + g * LDS + i * 8 + tq * 2) = pack_bf16(a[0], a[1]);
        *reinterpret_cast<uint32_t*>(Ow + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(a[2], a[3]);
      }
      __syncwarp();
      bf16* og = outp + (b * Lk) * ldo_ + (int64_t)h * DH;
      for (int idx = lane; idx < 16 * CH; idx += 32) {
        int r = idx / CH, c = idx % CH;
        if (j0 + r < Lk)
          *reinterpret_cast<uint4*>(og + (int64_t)(j0 + r) * ldo_ + c * 8) = *reinterpret_cast<const uint4*>(Ow + r * LDS + c * 8);
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward for long key sequences (the 128 x N latents<-tracks cross-attention): same math as above,
// split in two launches so that no partial sums cross CTAs:
//   MODE 0 (dQ)    : CTA = (sequence, head, 32 query rows); the keys are streamed in chunks of KW.
//   MODE 1 (dK, dV): CTA = (sequence, head, KW keys); all Lq <= 160 query rows sit in shared memory.
// ------------------------------------------------------------------------------------------
template <int DH, int MODE>
__global__ void __launch_bounds__(256)
attn_bwd_win_mma_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
                        const bf16* __restrict__ v, int64_t ldv, const bf16* __restrict__ d_o, int64_t lddo,
                        bf16* __restrict__ dq, int64_t lddq, bf16* __restrict__ dk, int64_t lddk,
                        bf16* __restrict__ dv, int64_t lddv, const uint8_t* __restrict__ mask,
                        const float* __restrict__ stats, const float* __restrict__ delta, int heads, int Lq,
                        int Lk, int QW, int KW, int parts) {
  constexpr int LDS = DH + 8;
  constexpr int CH = DH / 8;
  constexpr float LOG2E = 1.4426950408889634f;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int nwarps = blockDim.x >> 5;
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Gs = Qs + (size_t)QW * LDS;
  bf16* Ks = Gs + (size_t)QW * LDS;
  bf16* Vs = Ks + (size_t)KW * LDS;
  bf16* Os = Vs + (size_t)KW * LDS;          // per-warp output staging [nwarps][16][LDS]
  float* mS = reinterpret_cast<float*>(Os + (size_t)nwarps * 16 * LDS);
  float* iS = mS + QW;
  float* dS = iS + QW;
  uint8_t* Ms = reinterpret_cast<uint8_t*>(dS + QW);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int part = blockIdx.x % parts;
  const int64_t bh = blockIdx.x / parts;
  const int64_t b = bh / heads;
  const int h = (int)(bh % heads);
  const int q0 = MODE == 0 ? part * QW : 0;
  const bf16* qg = q + (b * Lq) * ldq + (int64_t)h * DH;
  const bf16* gg = d_o + (b * Lq) * lddo + (int64_t)h * DH;
  const bf16* kg = k + (b * Lk) * ldk + (int64_t)h * DH;
  const bf16* vg = v + (b * Lk) * ldv + (int64_t)h * DH;

  for (int idx = tid; idx < QW * CH; idx += nthr) {
    int r = idx / CH, c = idx % CH;
    bf16* d0 = Qs + r * LDS + c * 8;
    bf16* d1 = Gs + r * LDS + c * 8;
    if (q0 + r < Lq) {
      __pipeline_memcpy_async(d0, qg + (int64_t)(q0 + r) * ldq + c * 8, 16);
      __pipeline_memcpy_async(d1, gg + (int64_t)(q0 + r) * lddo + c * 8, 16);
    } else {
      *reinterpret_cast<uint4*>(d0) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(d1) = make_uint4(0, 0, 0, 0);
    }
  }
  {
    const int64_t base = (b * heads + h) * Lq;
    for (int r = tid; r < QW; r += nthr) {
      const bool ok = q0 + r < Lq;
      mS[r] = ok ? stats[(base + q0 + r) * 2] : 0.f;
      iS[r] = ok ? stats[(base + q0 + r) * 2 + 1] : 0.f;
      dS[r] = ok ? delta[base + q0 + r] : 0.f;
    }
  }
  const int g = lane >> 2, tq = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;
  const int t_row = (lane & 7) + ((lane >> 3) & 1) * 8, t_col = (lane >> 4) * 8;
  bf16* Ow = Os + (size_t)warp * 16 * LDS;

  float acc[DH / 8][4], acc2[MODE == 1 ? DH / 8 : 1][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
  for (int i = 0; i < (MODE == 1 ? DH / 8 : 1); ++i) acc2[i][0] = acc2[i][1] = acc2[i][2] = acc2[i][3] = 0.f;

  const int k_begin = MODE == 0 ? 0 : part * KW;
  const int k_end = MODE == 0 ? Lk : min(Lk, k_begin + KW);
  for (int kc0 = k_begin; kc0 < k_end; kc0 += KW) {
    if (kc0 > k_begin) __syncthreads();
    for (int idx = tid; idx < KW * CH; idx += nthr) {
      int r = idx / CH, c = idx % CH;
      bf16* d0 = Ks + r * LDS + c * 8;
      bf16* d1 = Vs + r * LDS + c * 8;
      if (kc0 + r < Lk) {
        __pipeline_memcpy_async(d0, kg + (int64_t)(kc0 + r) * ldk + c * 8, 16);
        __pipeline_memcpy_async(d1, vg + (int64_t)(kc0 + r) * ldv + c * 8, 16);
      } else {
        *reinterpret_cast<uint4*>(d0) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(d1) = make_uint4(0, 0, 0, 0);
      }
    }
    for (int j = tid; j < KW; j += nthr)
      Ms[j] = kc0 + j < Lk ? (mask == nullptr ? 1 : (mask[b * Lk + kc0 + j] != 0 ? 1 : 0)) : 2;
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();

    if (MODE == 0) {
      // warp = 16 query rows of the window, keys of this chunk
      if (warp * 16 < QW) {
        const int r0 = warp * 16;
        const float m0 = mS[r0 + g], m1 = mS[r0 + g + 8], i0 = iS[r0 + g], i1 = iS[r0 + g + 8];
        const float e0 = dS[r0 + g], e1 = dS[r0 + g + 8];
        for (int j0 = 0; j0 < KW; j0 += 16) {
          if (kc0 + j0 >= Lk) break;
          float s[2][4], dp[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            uint32_t a[4], bb[4];
            ldsm_x4(a, Qs + (r0 + a_row) * LDS + kk * 16 + a_col);
            ldsm_x4(bb, Ks + (j0 + b_row) * LDS + kk * 16 + b_col);
            mma_bf16(s[0], a, bb[0], bb[1]);
            mma_bf16(s[1], a, bb[2], bb[3]);
            ldsm_x4(a, Gs + (r0 + a_row) * LDS + kk * 16 + a_col);
            ldsm_x4(bb, Vs + (j0 + b_row) * LDS + kk * 16 + b_col);
            mma_bf16(dp[0], a, bb[0], bb[1]);
            mma_bf16(dp[1], a, bb[2], bb[3]);
          }
          uint32_t pa[4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            float dsv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const uint8_t mk = Ms[j0 + nt * 8 + tq * 2 + e];
              const float lo = mk == 1 ? s[nt][e] : masked_logit_bf16(), hi = mk == 1 ? s[nt][2 + e] : masked_logit_bf16();
              const float p0 = mk != 2 ? exp2f((lo - m0) * LOG2E) * i0 : 0.f;
              const float p1 = mk != 2 ? exp2f((hi - m1) * LOG2E) * i1 : 0.f;
              dsv[e] = mk == 1 ? p0 * (dp[nt][e] - e0) : 0.f;
              dsv[2 + e] = mk == 1 ? p1 * (dp[nt][2 + e] - e1) : 0.f;
            }
            pa[nt * 2] = pack_bf16(dsv[0], dsv[1]);
            pa[nt * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
          }
#pragma unroll
          for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
            uint32_t bb[4];
            ldsm_x4_t(bb, Ks + (j0 + t_row) * LDS + nd2 * 16 + t_col);
            mma_bf16(acc[nd2 * 2], pa, bb[0], bb[1]);
            mma_bf16(acc[nd2 * 2 + 1], pa, bb[2], bb[3]);
          }
        }
      }
    } else {
      // warp = 16 keys of the window, all query rows
      if (warp * 16 < KW && kc0 + warp * 16 < Lk) {
        const int j0 = warp * 16;
        const uint8_t mk0 = Ms[j0 + g], mk1 = Ms[j0 + g + 8];
        for (int r0 = 0; r0 < QW; r0 += 16) {
          float s[2][4], dp[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            uint32_t a[4], bb[4];
            ldsm_x4(a, Ks + (j0 + a_row) * LDS + kk * 16 + a_col);
            ldsm_x4(bb, Qs + (r0 + b_row) * LDS + kk * 16 + b_col);
            mma_bf16(s[0], a, bb[0], bb[1]);
            mma_bf16(s[1], a, bb[2], bb[3]);
            ldsm_x4(a, Vs + (j0 + a_row) * LDS + kk * 16 + a_col);
            ldsm_x4(bb, Gs + (r0 + b_row) * LDS + kk * 16 + b_col);
            mma_bf16(dp[0], a, bb[0], bb[1]);
            mma_bf16(dp[1], a, bb[2], bb[3]);
          }
          uint32_t pp[4], pd[4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int qi = r0 + nt * 8 + tq * 2 + e;
              const float m = mS[qi], il = iS[qi], dl = dS[qi];
              const float lo = mk0 == 1 ? s[nt][e] : masked_logit_bf16(), hi = mk1 == 1 ? s[nt][2 + e] : masked_logit_bf16();
              pv[e] = mk0 != 2 ? exp2f((lo - m) * LOG2E) * il : 0.f;
              pv[2 + e] = mk1 != 2 ? exp2f((hi - m) * LOG2E) * il : 0.f;
              dsv[e] = mk0 == 1 ? pv[e] * (dp[nt][e] - dl) : 0.f;
              dsv[2 + e] = mk1 == 1 ? pv[2 + e] * (dp[nt][2 + e] - dl) : 0.f;
            }
            pp[nt * 2] = pack_bf16(pv[0], pv[1]);
            pp[nt * 2 + 1] = pack_bf16(pv[2], pv[3]);
            pd[nt * 2] = pack_bf16(dsv[0], dsv[1]);
            pd[nt * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
          }
#pragma unroll
          for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
            uint32_t bb[4];
            ldsm_x4_t(bb, Gs + (r0 + t_row) * LDS + nd2 * 16 + t_col);
            mma_bf16(acc2[(MODE == 1 ? nd2 * 2 : 0)], pp, bb[0], bb[1]);
            mma_bf16(acc2[(MODE == 1 ? nd2 * 2 + 1 : 0)], pp, bb[2], bb[3]);
            ldsm_x4_t(bb, Qs + (r0 + t_row) * LDS + nd2 * 16 + t_col);
            mma_bf16(acc[nd2 * 2], pd, bb[0], bb[1]);
            mma_bf16(acc[nd2 * 2 + 1], pd, bb[2], bb[3]);
          }
        }
      }
    }
  }
  // outputs: MODE 0 -> dq rows of the window (acc); MODE 1 -> dk (acc), dv (acc2) rows of the key window
  const int npass = MODE == 0 ? 1 : 2;
  const int row0 = (MODE == 0 ? q0 : k_begin) + warp * 16;
  const int row_lim = MODE == 0 ? Lq : Lk;
  const bool has_rows = warp * 16 < (MODE == 0 ? QW : KW);
  for (int pass = 0; pass < npass; ++pass) {
    bf16* outp = MODE == 0 ? dq : (pass == 0 ? dk : dv);
    const int64_t ldo_ = MODE == 0 ? lddq : (pass == 0 ? lddk : lddv);
    if (has_rows) {
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        const float(&a)[4] = (MODE == 1 && pass == 1) ? acc2[MODE == 1 ? i : 0] : acc[i];
        *reinterpret_cast<uint32_t*>(Ow + g * LDS + i * 8 + tq * 2) = pack_bf16(a[0], a[1]);
        *reinterpret_cast<uint32_t*>(Ow + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(a[2], a[3]);
      }
      __syncwarp();
      bf16* og = outp + (b * row_lim) * ldo_ + (int64_t)h * DH;
      for (int idx = lane; idx < 16 * CH; idx += 32) {
        int r = idx / CH, c = idx % CH;
        if (row0 + r < row_lim)
          *reinterpret_cast<uint4*>(og + (int64_t)(row0 + r) * ldo_ + c * 8) = *reinterpret_cast<const uint4*>(Ow + r * LDS + c * 8);
      }
      __syncwarp();
    }
  }
}

template <int DH>
static int launch_bwd_win(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                          const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk,
                          void* dv, int64_t lddv, const uint8_t* mask, const float* stats, const float* delta,
                          int64_t batch, int heads, int Lq, int Lk, cudaStream_t st) {
  const int LqPad = (Lq + 15) / 16 * 16;
  const int KW = 128;
  auto smem_for = [&](int QW, int nwarps) { return (size_t)(2 * QW + 2 * KW + nwarps * 16) * (DH + 8) * 2 + (size_t)QW * 12 + KW; };
  {   // dQ: 32 query rows per CTA (2 compute warps, 4 loading warps), keys streamed
    const int QW = 32, threads = 128, parts = (Lq + QW - 1) / QW;
    size_t smem = smem_for(QW, threads / 32);
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_win_mma_kernel<DH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_bwd_win: smem attribute: %s", cudaGetErrorString(e));
    attn_bwd_win_mma_kernel<DH, 0><<<(unsigned)(batch * heads * parts), threads, smem, st>>>(
        (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)d_o, lddo, (bf16*)dq, lddq, (bf16*)dk, lddk,
        (bf16*)dv, lddv, mask, stats, delta, heads, Lq, Lk, QW, KW, parts);
  }
  {   // dK, dV: 128 keys per CTA (8 warps), every query row in shared memory
    const int QW = LqPad, threads = 256, parts = (Lk + KW - 1) / KW;
    size_t smem = smem_for(QW, threads / 32);
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_win_mma_kernel<DH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_bwd_win: smem attribute: %s", cudaGetErrorString(e));
    attn_bwd_win_mma_kernel<DH, 1><<<(unsigned)(batch * heads * parts), threads, smem, st>>>(
        (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)d_o, lddo, (bf16*)dq, lddq, (bf16*)dk, lddk,
        (bf16*)dv, lddv, mask, stats, delta, heads, Lq, Lk, QW, KW, parts);
  }
  return check_launch("attention_bwd_win_mma");
}

template <int DH>
static int launch_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk,
                      void* dv, int64_t lddv, const uint8_t* mask, const float* stats, const float* delta,
                      int64_t batch, int heads, int Lq, int Lk, cudaStream_t st) {
  const int LqPad = (Lq + 15) / 16 * 16, LkPad = (Lk + 15) / 16 * 16;
  const int nwarps = (LqPad > LkPad ? LqPad : LkPad) / 16;
  size_t smem = (size_t)(2 * LqPad + 2 * LkPad + nwarps * 16) * (DH + 8) * 2 + (size_t)LqPad * 12 + LkPad;
  static size_t max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_bwd_mma: smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
    max_set = smem;
  }
  int64_t blocks = batch * heads;
  SPA3D_REQUIRE(blocks < (1ll << 31), "attention_bwd_mma: grid too large");
  attn_bwd_mma_kernel<DH><<<(unsigned)blocks, nwarps * 32, smem, st>>>(
      (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)d_o, lddo, (bf16*)dq, lddq, (bf16*)dk,
      lddk, (bf16*)dv, lddv, mask, stats, delta, heads, Lq, Lk, LqPad, LkPad);
  return check_launch("attention_bwd_mma");
}

template <int DH>
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  void* o, int64_t ldo, const uint8_t* mask, float* stats, int64_t batch, int heads,
                  int Lq, int Lk, cudaStream_t st) {
  // whole-sequence tile when the keys fit (self-attention); otherwise 32 query rows per CTA and
  // the keys streamed in chunks of 128 (more CTAs, two or more resident per SM)
  const bool one_tile = Lk <= 256;
  const int LqPad = (Lq + 15) / 16 * 16;
  const int QR = one_tile ? LqPad : 32;
  const int KC = one_tile ? (Lk + KT - 1) / KT * KT : 128;
  const int q_tiles = (Lq + QR - 1) / QR;
  size_t smem = (size_t)(QR + 2 * KC) * (DH + 8) * 2 + KC;
  static size_t max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_mma: smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
    max_set = smem;
  }
  int64_t blocks = batch * heads * q_tiles;
  SPA3D_REQUIRE(blocks < (1ll << 31), "attention_mma: grid too large");
  // at least 4 warps per CTA: with few query rows (the pruned last layer has ONE) the extra warps
  // only help staging K and V, but a 1-warp CTA leaves the SM at 3 resident warps
  const int threads = QR * 2 < 128 ? 128 : QR * 2;
  attn_fwd_mma_kernel<DH><<<(unsigned)blocks, threads, smem, st>>>(
      (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, mask, stats, heads, Lq, Lk, QR, KC, q_tiles);
  return check_launch("attention_fwd_mma");
}

}  // namespace am

bool attention_fwd_mma_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk,
                                  int64_t ldv, int64_t ldo) {
  if (dtype != SPA3D_BF16) return false;
  if (Dh != 64 && Dh != 96) return false;
  if (Lq > 256 && Lk <= 256) return false;  // one-tile mode keeps all queries of a sequence in one CTA
  return (ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (ldo % 8 == 0);
}

int attention_fwd_mma(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse_out,
                      int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al(q) && al(k) && al(v) && al(o), "attention_mma: q/k/v/o must be 16-byte aligned");
  if (Dh == 96) return am::launch<96>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Lk, st);
  return am::launch<64>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Lk, st);
}

bool attention_bwd_mma_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv,
                                  int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv) {
  if (dtype != SPA3D_BF16) return false;
  if (Dh != 64 && Dh != 96) return false;
  if (Lq > 160) return false;   // every query row (q, dO) of a sequence sits in one CTA's shared memory
  return (ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (lddo % 8 == 0) && (lddq % 8 == 0) &&
         (lddk % 8 == 0) && (lddv % 8 == 0);
}

int attention_bwd_mma(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk,
                      void* dv, int64_t lddv, const uint8_t* key_mask, const float* stats,
                      const float* delta, int64_t batch, int heads, int Lq, int Lk, int Dh,
                      cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al(q) && al(k) && al(v) && al(d_o) && al(dq) && al(dk) && al(dv), "attention_bwd_mma: operands must be 16-byte aligned");
  if (Lk > 160) {   // long key sequences: dQ and dK/dV as two launches over query / key windows
    if (Dh == 96)
      return am::launch_bwd_win<96>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, Lq, Lk, st);
    return am::launch_bwd_win<64>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, Lq, Lk, st);
  }
  if (Dh == 96)
    return am::launch_bwd<96>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, Lq, Lk, st);
  return am::launch_bwd<64>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, Lq, Lk, st);
}

}  // namespace spa3d
