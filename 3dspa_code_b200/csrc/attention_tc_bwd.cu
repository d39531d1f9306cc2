// Backward of the short-sequence self-attention core on tcgen05 + TMEM (bf16, L <= 160):
//   P = exp(S - m) / l,  dP = dO V^T,  dS = P o (dP - delta),  dQ = dS K,  dK = dS^T Q,  dV = P^T dO
// per (sequence, head), whole sequence in one CTA pass (attention.py:175 under jax.value_and_grad,
// train.py:161-162).  m, 1/l come from the forward (row statistics); delta_i = sum_d O_id dO_id is
// evaluated in place as sum_j P_ij dP_ij, so the forward output O is not read at all.
//
// No operand is ever transposed in memory: P and dS are written once to shared memory as
// [query][key] tiles (64B-swizzled atoms of 32 keys) and are read as the K-major A operand of
// dQ = dS K and as the MN-major B operand of the TRANSPOSED products dV^T = dO^T P and
// dK^T = Q^T dS, whose A operands are dO and Q read MN-major in place.  TMEM: columns [0,160) are
// time-shared by S, dP and dQ of the current 128-row query tile, [160,320) hold dV^T, [320,480) dK^T
// (lane = head channel, column = key), accumulated over the query tiles.
//
//   warp 0      TMA loads (Q, K, V, dO as 32-channel x L-row boxes, 3-D maps)
//   warp 1      MMA issuer
//   warps 2..9  element-wise phases and epilogues: two warps per TMEM lane quarter, each owning
//               half of the 32-key chunks (no row reductions are needed in the backward)
//   warp 10     builds the key-bias operand of the next item (mask as a rank-1 MMA K-step)
#include "tc_ptx.cuh"

namespace spa3d {
namespace tb {

using namespace tc;

constexpr int THREADS = 352;
constexpr int CW = 8;   // compute warps

__device__ __forceinline__ uint64_t desc_k64(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_k32(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn64(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// kind::f16, fp32 accumulate, bf16 operands, M = 128
__host__ __device__ constexpr uint32_t idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Maps {
  CUtensorMap q, k, v, g, dq, dk, dv;
};

// CROSS = true: one 128-key chunk of a (sequence, head) of the latents<-tracks cross-attention backward (Lq <= 128 queries, Lk keys
// as ceil(Lk / 128) chunks, one work item each).  dK / dV of the chunk's own keys are complete and leave as in the self case;
// dQ is a partial sum over this chunk's keys: it goes to a workspace as fp32 and attn_cross_dq_merge_kernel adds the chunks.
// delta_i = sum_d O_id dO_id spans ALL keys, so it is read from `delta` (attention_delta) instead of being formed in place.
struct CrossBwdArgs {
  int Lk, nchunks;
  float* part_dq;   // [items][128][DH]
};

template <int DH, int LPAD, int MT, bool CROSS = false>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ Maps tm, const uint8_t* __restrict__ mask, const float* __restrict__ stats,
                   const float* __restrict__ delta, int64_t items, int heads, int L, CrossBwdArgs cx) {
  static_assert(!CROSS || (MT == 1 && LPAD == 128), "cross-attention items are 128 queries x 128 keys");
  auto item_b = [&](int64_t it) -> int { return CROSS ? (int)(it / ((int64_t)heads * cx.nchunks)) : (int)(it / heads); };
  auto item_h = [&](int64_t it) -> int { return CROSS ? (int)((it / cx.nchunks) % heads) : (int)(it % heads); };
  auto item_k0 = [&](int64_t it) -> int { return CROSS ? (int)(it % cx.nchunks) * 128 : 0; };
  const int Lkeys = CROSS ? cx.Lk : L;
  constexpr int DA = DH / 32;            // 32-channel atoms per row
  constexpr int KA = LPAD / 32;          // 32-key atoms per P / dS row
  constexpr int OPA = LPAD * 64;         // bytes of one 32-channel atom region of Q / K / V / dO
  constexpr int OP_BYTES = DA * OPA;
  constexpr int PT_BYTES = KA * 128 * 64;   // one [128 query][LPAD key] bf16 tile
  constexpr int C_S = 0, C_DV = LPAD, C_DK = 2 * LPAD;
  static_assert(3 * LPAD <= 512 && DH <= LPAD, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + OP_BYTES;
  uint8_t* sV = sK + OP_BYTES;
  uint8_t* sG = sV + OP_BYTES;             // dO
  uint8_t* sP = sG + OP_BYTES;             // P of the current query tile (also epilogue staging)
  uint8_t* sD = sP + PT_BYTES;             // dS of the current query tile
  uint8_t* sE = sD + PT_BYTES;             // ones operand [MT*128][32 B]
  uint8_t* sB = sE + MT * 128 * 32;        // [2][LPAD][32 B] key bias operand
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * LPAD * 32);
  uint64_t* ld_full = bars, *ld_empty = bars + 1, *s_full = bars + 2, *p_done = bars + 3, *dp_full = bars + 4;
  uint64_t* ds_done = bars + 5, *dq_full = bars + 6, *dq_read = bars + 7, *m_full = bars + 8, *m_empty = bars + 10;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
  float* dpart = reinterpret_cast<float*>(bars + 16);   // [2 halves][128 rows] partial row sums of P o dP

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.v)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.g)) : "memory");
    mbar_init(ld_full, 1); mbar_init(ld_empty, 1); mbar_init(s_full, 1); mbar_init(p_done, CW * 32);
    mbar_init(dp_full, 1); mbar_init(ds_done, CW * 32); mbar_init(dq_full, 1); mbar_init(dq_read, CW * 32);
    mbar_init(&m_full[0], 1); mbar_init(&m_full[1], 1); mbar_init(&m_empty[0], 1); mbar_init(&m_empty[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < MT * 128; r += THREADS) {   // ones operand: element 0 of every row = 1.0
    const int pc = (r >> 2) & 1;
    *reinterpret_cast<uint4*>(sE + r * 32 + pc * 16) = make_uint4(0x3F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sE + r * 32 + (pc ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA loads =====================
    if (lane == 0) {
      uint32_t ph = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1) {
        const int b = item_b(it), h = item_h(it), k0 = item_k0(it);
        mbar_wait(ld_empty, ph ^ 1);
        mbar_arrive_expect_tx(ld_full, (uint32_t)(4 * OP_BYTES));
#pragma unroll
        for (int a = 0; a < DA; ++a) {
          tma_load_3d(sQ + a * OPA, &tm.q, h * DH + a * 32, 0, b, ld_full);
          tma_load_3d(sK + a * OPA, &tm.k, h * DH + a * 32, k0, b, ld_full);
          tma_load_3d(sG + a * OPA, &tm.g, h * DH + a * 32, 0, b, ld_full);
          tma_load_3d(sV + a * OPA, &tm.v, h * DH + a * 32, k0, b, ld_full);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idS = idesc(LPAD, false, false), idQ = idesc(DH, false, true), idT = idesc(LPAD, true, true);
      uint32_t ph = 0;
      int n = 0, tcount = 0;   // tcount = query tiles processed so far (phase of the per-tile barriers)
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1, ++n) {
        const int mb = n & 1;
        mbar_wait(ld_full, ph);
        mbar_wait(&m_full[mb], (n >> 1) & 1);
#pragma unroll 1
        for (int t = 0; t < MT; ++t, ++tcount) {
          const uint32_t tp = tcount & 1;
          const int qk = (LPAD - t * 128 < 128 ? LPAD - t * 128 : 128) / 16;   // 16-row K-steps over this tile's query rows
          mbar_wait(dq_read, tp ^ 1);   // columns [0,160) drained, P / dS tiles free
          tcgen05_fence_after();
          // S_t = Q_t K^T + ones x bias
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sQ + (kk >> 1) * OPA + t * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sK + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
          umma_bf16(tmem_base + C_S, desc_k32(smem_u32(sE + t * (128 * 32))), desc_k32(smem_u32(sB + mb * (LPAD * 32))), idS, 1u);
          umma_commit(s_full);
          if (t == MT - 1) umma_commit(&m_empty[mb]);
          // dP_t = dO_t V^T (same columns, once the threads have turned S into P)
          mbar_wait(p_done, tp);
          tcgen05_fence_after();
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sG + (kk >> 1) * OPA + t * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sV + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
          umma_commit(dp_full);
          // dQ_t = dS_t K ;  dV^T += dO_t^T P_t ;  dK^T += Q_t^T dS_t
          mbar_wait(ds_done, tp);
          tcgen05_fence_after();
#pragma unroll
          for (int ks = 0; ks < LPAD / 16; ++ks)
            umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sD + (ks >> 1) * (128 * 64) + (ks & 1) * 32)),
                      desc_mn64(smem_u32(sK + ks * 1024), OPA), idQ, ks > 0 ? 1u : 0u);
          for (int j = 0; j < qk; ++j) {
            const uint32_t acc = (t > 0 || j > 0) ? 1u : 0u;
            umma_bf16(tmem_base + C_DV, desc_mn64(smem_u32(sG + (t * 128 + j * 16) * 64), OPA),
                      desc_mn64(smem_u32(sP + j * 1024), 128 * 64), idT, acc);
            umma_bf16(tmem_base + C_DK, desc_mn64(smem_u32(sQ + (t * 128 + j * 16) * 64), OPA),
                      desc_mn64(smem_u32(sD + j * 1024), 128 * 64), idT, acc);
          }
          umma_commit(dq_full);
        }
        umma_commit(ld_empty);
      }
    }
  } else if (warp == 2 + CW) {
    // ===================== key-bias operand builder (one item ahead) =====================
    const uint32_t NEG_BIG = 0xF14Au, NEG_INF = 0xFF80u;   // bf16(-1e30), bf16(-inf)
    int n = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int b = item_b(it), k0 = item_k0(it);
      const int mb = n & 1;
      if (lane == 0) mbar_wait(&m_empty[mb], ((n >> 1) & 1) ^ 1);
      __syncwarp();
      uint8_t* dstb = sB + mb * (LPAD * 32);
      for (int j = lane; j < LPAD; j += 32) {
        uint32_t val = NEG_INF;
        if (k0 + j < Lkeys) val = (mask == nullptr || mask[(int64_t)b * Lkeys + k0 + j] != 0) ? 0u : NEG_BIG;
        const int pc = (j >> 2) & 1;
        *reinterpret_cast<uint4*>(dstb + j * 32 + pc * 16) = make_uint4(val, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dstb + j * 32 + (pc ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&m_full[mb]);
    }
  } else {
    // ===================== element-wise phases + epilogues =====================
    const int cw = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter
    const int half = cw >> 2;              // which half of the key chunks / output chunks
    const int c_lo = half == 0 ? 0 : (KA + 1) / 2, c_hi = half == 0 ? (KA + 1) / 2 : KA;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rloc = quarter * 32 + lane;  // row inside the 128-row tile
    const int sw64 = (lane >> 1) & 3;
    uint8_t* pRow = sP + rloc * 64;
    uint8_t* dRow = sD + rloc * 64;
    uint8_t* slab = sP + cw * 4096;        // 2 x 2 KB staging per warp (P tile is free when the epilogues run)
    constexpr float LOG2E = 1.4426950408889634f;
    int tcount = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int b = item_b(it), h = item_h(it), k0 = item_k0(it);
      const int64_t sbase = ((int64_t)b * heads + h) * L;
#pragma unroll 1
      for (int t = 0; t < MT; ++t, ++tcount) {
        const uint32_t tp = tcount & 1;
        const int row = t * 128 + rloc;
        const bool tile_live = t * 128 + quarter * 32 < L;     // this warp has valid query rows in tile t
        float m = 0.f, il = 0.f;
        if (row < L) {
          m = stats[(sbase + row) * 2];
          il = stats[(sbase + row) * 2 + 1];
        }
        const bool row_grad = m > -1e29f;    // a fully masked row: its logits are constants, dS = 0
        // ---- phase A: P = exp(S - m) / l -> bf16, [query][key] tile ----
        mbar_wait(s_full, tp);
        tcgen05_fence_after();
        if (tile_live) {
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + c * 32), r);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_fast((__uint_as_float(r[2 * i]) - m) * LOG2E) * il;
              const float p1 = ex2_fast((__uint_as_float(r[2 * i + 1]) - m) * LOG2E) * il;
              __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
              w[i] = *reinterpret_cast<uint32_t*>(&hb);
            }
            uint8_t* dst = pRow + c * (128 * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
        }
        else {
          // rows past the sequence end: dO / Q rows are zero there, but 0 x (stale shared memory) must not be NaN
          for (int c = c_lo; c < c_hi; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              *reinterpret_cast<uint4*>(pRow + c * (128 * 64) + (j << 4)) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(dRow + c * (128 * 64) + (j << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(p_done);
        // ---- phase B: dS = P o (dP - delta) -> bf16 tile ----
        mbar_wait(dp_full, tp);
        tcgen05_fence_after();
        if (tile_live) {
          // delta_i = sum_d O_id dO_id = sum_j P_ij dP_ij: a row sum over the keys, half of them in the
          // partner warp of this lane quarter -> exchange through shared memory
          float dl;
          if constexpr (CROSS) {
            dl = row < L ? delta[sbase + row] : 0.f;   // over all keys of the row, not only this chunk's
          } else {
            float part = 0.f;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
              uint32_t r[32];
              tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + c * 32), r);
              const uint8_t* src = pRow + c * (128 * 64);
              uint4 pw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) pw[j] = *reinterpret_cast<const uint4*>(src + ((j ^ sw64) << 4));
              tmem_ld_wait();
              const uint32_t* pwu = reinterpret_cast<const uint32_t*>(pw);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&pwu[i]);
                part = fmaf(__low2float(pb), __uint_as_float(r[2 * i]), part);
                part = fmaf(__high2float(pb), __uint_as_float(r[2 * i + 1]), part);
              }
            }
            dpart[half * 128 + rloc] = part;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            dl = part + dpart[(half ^ 1) * 128 + rloc];
          }
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + c * 32), r);
            const uint8_t* src = pRow + c * (128 * 64);
            uint4 pw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pw[j] = *reinterpret_cast<const uint4*>(src + ((j ^ sw64) << 4));
            tmem_ld_wait();
            const uint32_t* pwu = reinterpret_cast<const uint32_t*>(pw);
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&pwu[i]);
              float d0 = __low2float(pb) * (__uint_as_float(r[2 * i]) - dl);
              float d1 = __high2float(pb) * (__uint_as_float(r[2 * i + 1]) - dl);
              if (!row_grad) d0 = d1 = 0.f;
              __nv_bfloat162 hb = __floats2bfloat162_rn(d0, d1);
              w[i] = *reinterpret_cast<uint32_t*>(&hb);
            }
            uint8_t* dst = dRow + c * (128 * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          if constexpr (!CROSS) asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // dpart is rewritten for the next tile
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(ds_done);
        // ---- dQ_t epilogue (and, after the last tile, dV^T / dK^T) ----
        mbar_wait(dq_full, tp);
        tcgen05_fence_after();
        int sb = 0;
        if (tile_live) {
          const int o_lo = half == 0 ? 0 : (DA + 1) / 2, o_hi = half == 0 ? (DA + 1) / 2 : DA;
          for (int c = o_lo; c < o_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + c * 32), r);
            tmem_ld_wait();
            if constexpr (CROSS) {   // partial dQ of this key chunk, fp32, 128 contiguous bytes per thread and channel atom
              float* drow = cx.part_dq + ((int64_t)it * 128 + rloc) * DH + c * 32;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(drow + j * 4) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
              continue;
            }
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            uint8_t* dst = slab + sb * 2048 + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
                w[e] = *reinterpret_cast<uint32_t*>(&hb);
              }
              *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tm.dq, slab + sb * 2048, h * DH + c * 32, t * 128 + quarter * 32, b);
              bulk_commit();
            }
            sb ^= 1;
          }
        }
        if (t == MT - 1 && quarter * 32 < DH) {
          // transposed accumulators: lane = head channel quarter*32 + lane, column = key.
          // half 0 stores dV, half 1 stores dK; each 32-key chunk goes out as a [32 keys][32 channels] box.
          const uint32_t cbase = half == 0 ? C_DV : C_DK;
          const CUtensorMap* omap = half == 0 ? &tm.dv : &tm.dk;
          for (int c = 0; c < KA; ++c) {
            if (k0 + c * 32 >= Lkeys) break;
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_off + cbase + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            uint8_t* dst = slab + sb * 2048;
#pragma unroll
            for (int i = 0; i < 32; ++i) {   // key i of the chunk = row i of the box; this lane's channel = column lane
              const __nv_bfloat16 hv = __float2bfloat16_rn(__uint_as_float(r[i]));
              *reinterpret_cast<__nv_bfloat16*>(dst + i * 64 + ((((lane >> 3) ^ ((i >> 1) & 3))) << 4) + (lane & 7) * 2) = hv;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(omap, dst, h * DH + quarter * 32, k0 + c * 32, b);
              bulk_commit();
            }
            sb ^= 1;
          }
        }
        if (lane == 0) bulk_wait_read<0>();   // staging lives in the P tile of the next query tile / item
        __syncwarp();
        tcgen05_fence_before();
        mbar_arrive(dq_read);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int make_map3(CUtensorMap* map, const void* ptr, int cols, int L, int64_t batch, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SPA3D_REQUIRE(fn != nullptr, "attention_tc_bwd: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPA3D_REQUIRE(r == CUDA_SUCCESS, "attention_tc_bwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int DH, int LPAD, int MT>
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                  int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                  const uint8_t* mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                  cudaStream_t st) {
  constexpr int DA = DH / 32, KA = LPAD / 32;
  constexpr int SMEM = 4 * DA * LPAD * 64 + 2 * KA * 128 * 64 + MT * 128 * 32 + 2 * LPAD * 32 + 128 + 1024 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  Maps tm;
  const int cols = heads * DH;
  if (make_map3(&tm.q, q, cols, L, batch, ldq, LPAD)) return 1;
  if (make_map3(&tm.k, k, cols, L, batch, ldk, LPAD)) return 1;
  if (make_map3(&tm.v, v, cols, L, batch, ldv, LPAD)) return 1;
  if (make_map3(&tm.g, d_o, cols, L, batch, lddo, LPAD)) return 1;
  if (make_map3(&tm.dq, dq, cols, L, batch, lddq, 32)) return 1;
  if (make_map3(&tm.dk, dk, cols, L, batch, lddk, 32)) return 1;
  if (make_map3(&tm.dv, dv, cols, L, batch, lddv, 32)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<DH, LPAD, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_tc_bwd: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t items = batch * heads;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_bwd_tc_kernel<DH, LPAD, MT><<<grid, THREADS, SMEM, st>>>(tm, mask, stats, delta, items, heads, L, CrossBwdArgs{0, 1, nullptr});
  return check_launch("attention_bwd_tc");
}

// dq[b, row, h*DH + d] = sum over key chunks of the partial dQ tiles (fp32) -> bf16
template <int DH>
__global__ void attn_cross_dq_merge_kernel(const float* __restrict__ part_dq, bf16* __restrict__ dq, int64_t lddq, int64_t rows_total,
                                           int heads, int Lq, int nchunks) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b * heads + h) * Lq + row
  const int lane = threadIdx.x & 31;
  if (w >= rows_total) return;
  const int row = (int)(w % Lq);
  const int64_t bh = w / Lq;
  const int h = (int)(bh % heads);
  const int64_t b = bh / heads;
  float acc[DH / 32];
#pragma unroll
  for (int i = 0; i < DH / 32; ++i) acc[i] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float* src = part_dq + ((bh * nchunks + c) * 128 + row) * DH;
#pragma unroll
    for (int i = 0; i < DH / 32; ++i) acc[i] += src[i * 32 + lane];
  }
  bf16* dst = dq + (b * Lq + row) * lddq + h * DH;
#pragma unroll
  for (int i = 0; i < DH / 32; ++i) dst[i * 32 + lane] = __float2bfloat16_rn(acc[i]);
}

template <int DH>
static int launch_cross(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                        int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const uint8_t* mask,
                        const float* stats, const float* delta, float* workspace, int64_t batch, int heads, int Lq, int Lk, cudaStream_t st) {
  constexpr int DA = DH / 32, KA = 4;
  constexpr int SMEM = 4 * DA * 128 * 64 + 2 * KA * 128 * 64 + 128 * 32 + 2 * 128 * 32 + 128 + 1024 + 1024;
  Maps tm;
  const int cols = heads * DH;
  if (make_map3(&tm.q, q, cols, Lq, batch, ldq, 128)) return 1;
  if (make_map3(&tm.k, k, cols, Lk, batch, ldk, 128)) return 1;
  if (make_map3(&tm.v, v, cols, Lk, batch, ldv, 128)) return 1;
  if (make_map3(&tm.g, d_o, cols, Lq, batch, lddo, 128)) return 1;
  tm.dq = tm.q;   // unused: dQ leaves through the workspace
  if (make_map3(&tm.dk, dk, cols, Lk, batch, lddk, 32)) return 1;
  if (make_map3(&tm.dv, dv, cols, Lk, batch, lddv, 32)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<DH, 128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_cross_bwd: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int nchunks = (Lk + 127) / 128;
  const int64_t items = batch * heads * nchunks;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_bwd_tc_kernel<DH, 128, 1, true><<<grid, THREADS, SMEM, st>>>(tm, mask, stats, delta, items, heads, Lq, CrossBwdArgs{Lk, nchunks, workspace});
  if (int rc = check_launch("attention_cross_bwd")) return rc;
  const int64_t rows_total = batch * heads * Lq;
  attn_cross_dq_merge_kernel<DH><<<(unsigned)((rows_total + 7) / 8), 256, 0, st>>>(workspace, (bf16*)dq, lddq, rows_total, heads, Lq, nchunks);
  return check_launch("attention_cross_dq_merge");
}

}  // namespace tb

bool attention_bwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t lddo,
                                 int64_t lddq, int64_t lddk, int64_t lddv) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SPA3D_ATTN_TC_BWD");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || dtype != SPA3D_BF16 || Lq != Lk || Lq < 2 || Lq > 160) return false;
  if (Dh != 96 && Dh != 64) return false;
  return ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0;
}

int attention_cross_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o, int64_t lddo,
                           void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const uint8_t* key_mask, const float* stats,
                           const float* delta, float* workspace, int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st) {
  using namespace tb;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al16(q) && al16(k) && al16(v) && al16(d_o) && al16(dk) && al16(dv) && al16(workspace) && workspace != nullptr,
                "attention_cross_bwd: operands must be 16-byte aligned");
  if (Dh == 96)
    return launch_cross<96>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, workspace, batch, heads, Lq, Lk, st);
  return launch_cross<64>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, workspace, batch, heads, Lq, Lk, st);
}

int attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                     int Dh, cudaStream_t st) {
  using namespace tb;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al(q) && al(k) && al(v) && al(d_o) && al(dq) && al(dk) && al(dv), "attention_tc_bwd: operands must be 16-byte aligned");
  if (Dh == 96) {
    if (L <= 128) return launch<96, 128, 1>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, L, st);
    return launch<96, 160, 2>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, L, st);
  }
  if (L <= 128) return launch<64, 128, 1>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, L, st);
  return launch<64, 160, 2>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, L, st);
}

}  // namespace spa3d
