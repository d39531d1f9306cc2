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
//   warps 2..   element-wise phases and epilogues: HW warps per TMEM lane quarter (4 for the two-tile self-attention shapes,
//               2 otherwise), each owning a share of the 32-key chunks; the only cross-warp value is the row sum delta.
//               These phases are latency-bound chains (tcgen05.ld -> ex2 / fma -> st.shared) that sit between dependent MMAs,
//               so their length is set by the chunks ONE warp walks: 4 warps per quarter cut it from 3 chunks to 2.
//   last warp   builds the key-bias operand of the next item (mask as a rank-1 MMA K-step)
// Loads are released in two groups: V and dO are last read by the dP / dV products of the last query tile, so the next
// item's V / dO stream in under the rest of the item (phase B, dQ / dK products, all epilogues); Q and K follow at the end.
#include "tc_ptx.cuh"

namespace spa3d {
namespace tb {

using namespace tc;

__host__ __device__ constexpr int threads_for(int hw) { return (3 + 4 * hw) * 32; }

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

template <int DH, int LPAD, int MT, bool CROSS = false, int HW = 2>
__global__ void __launch_bounds__(threads_for(HW), 1)
attn_bwd_tc_kernel(const __grid_constant__ Maps tm, const uint8_t* __restrict__ mask, const float* __restrict__ stats,
                   const float* __restrict__ delta, int64_t items, int heads, int L, CrossBwdArgs cx) {
  static_assert(!CROSS || (MT == 1 && LPAD == 128), "cross-attention items are 128 queries x 128 keys");
  auto item_b = [&](int64_t it) -> int { return CROSS ? (int)(it / ((int64_t)heads * cx.nchunks)) : (int)(it / heads); };
  auto item_h = [&](int64_t it) -> int { return CROSS ? (int)((it / cx.nchunks) % heads) : (int)(it % heads); };
  auto item_k0 = [&](int64_t it) -> int { return CROSS ? (int)(it % cx.nchunks) * 128 : 0; };
  const int Lkeys = CROSS ? cx.Lk : L;
  constexpr int THREADS = threads_for(HW);
  constexpr int CW = 4 * HW;             // compute warps
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
  uint64_t* qk_full = bars, *qk_empty = bars + 1, *s_full = bars + 2, *p_done = bars + 3, *dp_full = bars + 4;
  uint64_t* ds_done = bars + 5, *dq_full = bars + 6, *dq_read = bars + 7, *m_full = bars + 8, *m_empty = bars + 10;
  uint64_t* vg_full = bars + 12, *vg_empty = bars + 13;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);
  float* dpart = reinterpret_cast<float*>(bars + 16);   // [HW parts][128 rows] partial row sums of P o dP

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.v)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.g)) : "memory");
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(vg_full, 1); mbar_init(vg_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_done, CW * 32);
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
      // V / dO of an item are released before its Q / K, so per item the wait order is (V, dO) then (Q, K)
      uint32_t ph = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1) {
        const int b = item_b(it), h = item_h(it), k0 = item_k0(it);
        if (it == (int64_t)blockIdx.x) {   // first item: Q / K first, the S product needs them first
          mbar_arrive_expect_tx(qk_full, (uint32_t)(2 * OP_BYTES));
#pragma unroll
          for (int a = 0; a < DA; ++a) {
            tma_load_3d(sQ + a * OPA, &tm.q, h * DH + a * 32, 0, b, qk_full);
            tma_load_3d(sK + a * OPA, &tm.k, h * DH + a * 32, k0, b, qk_full);
          }
        }
        mbar_wait(vg_empty, ph ^ 1);
        mbar_arrive_expect_tx(vg_full, (uint32_t)(2 * OP_BYTES));
#pragma unroll
        for (int a = 0; a < DA; ++a) {
          tma_load_3d(sG + a * OPA, &tm.g, h * DH + a * 32, 0, b, vg_full);
          tma_load_3d(sV + a * OPA, &tm.v, h * DH + a * 32, k0, b, vg_full);
        }
        if (it != (int64_t)blockIdx.x) {
          mbar_wait(qk_empty, ph ^ 1);
          mbar_arrive_expect_tx(qk_full, (uint32_t)(2 * OP_BYTES));
#pragma unroll
          for (int a = 0; a < DA; ++a) {
            tma_load_3d(sQ + a * OPA, &tm.q, h * DH + a * 32, 0, b, qk_full);
            tma_load_3d(sK + a * OPA, &tm.k, h * DH + a * 32, k0, b, qk_full);
          }
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
        mbar_wait(qk_full, ph);
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
          if (t == 0) mbar_wait(vg_full, ph);
          mbar_wait(p_done, tp);
          tcgen05_fence_after();
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sG + (kk >> 1) * OPA + t * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sV + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
          umma_commit(dp_full);
          // dV^T += dO_t^T P_t needs only P: it runs under phase B, and after the last tile V and dO are free
          for (int j = 0; j < qk; ++j)
            umma_bf16(tmem_base + C_DV, desc_mn64(smem_u32(sG + (t * 128 + j * 16) * 64), OPA),
                      desc_mn64(smem_u32(sP + j * 1024), 128 * 64), idT, (t > 0 || j > 0) ? 1u : 0u);
          if (t == MT - 1) umma_commit(vg_empty);
          // dQ_t = dS_t K ;  dK^T += Q_t^T dS_t
          mbar_wait(ds_done, tp);
          tcgen05_fence_after();
#pragma unroll
          for (int ks = 0; ks < LPAD / 16; ++ks)
            umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sD + (ks >> 1) * (128 * 64) + (ks & 1) * 32)),
                      desc_mn64(smem_u32(sK + ks * 1024), OPA), idQ, ks > 0 ? 1u : 0u);
          for (int j = 0; j < qk; ++j)
            umma_bf16(tmem_base + C_DK, desc_mn64(smem_u32(sQ + (t * 128 + j * 16) * 64), OPA),
                      desc_mn64(smem_u32(sD + j * 1024), 128 * 64), idT, (t > 0 || j > 0) ? 1u : 0u);
          umma_commit(dq_full);
        }
        umma_commit(qk_empty);
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
    const int part = cw >> 2;              // which share of the key chunks / output chunks
    const int c_lo = part * KA / HW, c_hi = (part + 1) * KA / HW;
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
            float psum = 0.f;
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
                psum = fmaf(__low2float(pb), __uint_as_float(r[2 * i]), psum);
                psum = fmaf(__high2float(pb), __uint_as_float(r[2 * i + 1]), psum);
              }
            }
            dpart[part * 128 + rloc] = psum;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(HW * 32) : "memory");
            dl = 0.f;
#pragma unroll
            for (int hp = 0; hp < HW; ++hp) dl += dpart[hp * 128 + rloc];   // same order in every warp of the quarter
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
          if constexpr (!CROSS) asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(HW * 32) : "memory");   // dpart is rewritten for the next tile
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(ds_done);
        // ---- dQ_t epilogue (and, after the last tile, dV^T / dK^T) ----
        mbar_wait(dq_full, tp);
        tcgen05_fence_after();
        int sb = 0;
        if (tile_live) {
          const int o_lo = part * DA / HW, o_hi = (part + 1) * DA / HW;
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
          // transposed accumulators: lane = head channel quarter*32 + lane, column = key.  The 2 * KA units (dV chunks, then
          // dK chunks) are shared out over the HW warps of the quarter; each 32-key chunk goes out as a [32 keys][32 channels] box.
          for (int u = part * 2 * KA / HW; u < (part + 1) * 2 * KA / HW; ++u) {
            const int c = u < KA ? u : u - KA;
            const uint32_t cbase = u < KA ? C_DV : C_DK;
            const CUtensorMap* omap = u < KA ? &tm.dv : &tm.dk;
            if (k0 + c * 32 >= Lkeys) continue;
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

// ---------------------------------------------------------------------------------------------------------------------
// 128 < L <= 152 (the track transformers: L = 151 and L = 129): "transposed tail".
//
// tcgen05.ld moves 16 B per clock per TMEM lane quarter whatever the number of useful lanes, so in the two-tile kernel above
// the 23-row (or 1-row) second query tile costs as many TMEM read clocks as the full first tile (all of its 160 key columns are
// read through lane quarter 0), and its S -> P -> dP -> dS -> dQ chain runs after the first tile's chain.  Here the tail queries
// sit on the N (column) side instead: S^T = K Q_tail^T is [keys x 24 queries], two M tiles (keys 0..127, keys 128..159) of 24
// columns each, as are dP^T = V dO_tail^T and dQ_tail^T = K^T dS_tail^T [channels x 24].  The tail then costs 24-column reads,
// has its own TMEM columns (so its chain runs inside the waits of the main tile's chain), and its P^T / dS^T tile
// ([key][query], 10 KB) feeds the transposed dV^T / dK^T accumulators as a K-major B operand.
//   TMEM columns: [0,160) S / dP / dQ of the main tile, [160,312) dV^T, [312,464) dK^T (152 key columns), [464,512) the tail.
//   The softmax denominators and delta of the tail rows are per COLUMN here; the key-bias warp stages them in shared memory one
//   item ahead (delta_j = sum_d O_jd dO_jd from global memory: 24 rows x DH channels).
//   dS overwrites P in place (after the dV^T product has read P); the freed 40 KB hold the tail tile and per-warp store staging.
struct TailArgs {
  const bf16* o;
  int64_t ldo;
  const bf16* d_o;
  int64_t lddo;
};

constexpr int TT_THREADS = (2 + 16) * 32;   // TMA + side data, MMA, 16 element-wise warps

// NV values of this lane (its head channel) become column `lane` of rows ROW0 .. ROW0 + NV - 1 of a 64B-swizzled
// [32 rows][32 channels] bf16 box
template <int ROW0, int NV_>
__device__ __forceinline__ void stage_transposed(uint8_t* dst, const uint32_t* r, int lane) {
#pragma unroll
  for (int j = 0; j < NV_; ++j) {
    const int i = ROW0 + j;
    *reinterpret_cast<__nv_bfloat16*>(dst + i * 64 + ((((lane >> 3) ^ ((i >> 1) & 3))) << 4) + (lane & 7) * 2) =
        __float2bfloat16_rn(__uint_as_float(r[j]));
  }
}

constexpr int NT = 24;    // tail query columns (L - 128 <= 24)
constexpr int NV = 152;   // key columns of the dV^T / dK^T accumulators

template <int DH>
constexpr int tt_smem_bytes() {
  return 4 * (DH / 32) * 160 * 64 + 5 * 128 * 64 + 16 * 2048 + 160 * 64 + 128 * 32 + 2 * 160 * 32 + 1920 + 2048 + 256 + 1024;
}

// Development aid (-DSPA3D_ATTN_TRACE, e.g. SPA3D_NVCC_EXTRA=-DSPA3D_ATTN_TRACE python -m 3dspa_code_b200.build): clock64 stamps of the
// fifth item of CTA 0 at every barrier of the MMA thread (slots 0..15) and of three element-wise warps (32.., 64.., 96..);
// the launcher prints them after its third call.  This is how the phase costs quoted above were measured.
#ifdef SPA3D_ATTN_TRACE
__device__ long long g_attn_trace[128];
#define TRM(slot) if (blockIdx.x == 0 && n == 4) g_attn_trace[slot] = clock64();
#define TRC(slot) if (blockIdx.x == 0 && n == 4 && lane == 0 && (cw == 0 || cw == 14 || cw == 6)) g_attn_trace[(cw == 0 ? 32 : cw == 14 ? 64 : 96) + slot] = clock64();
#else
#define TRM(slot)
#define TRC(slot)
#endif
#ifdef SPA3D_ATTN_TRACE_SERIAL   // additionally wait for each MMA group to finish (serialises the tensor pipe: group durations, not a timeline)
#define PROBE(slot) { umma_commit(probe); mbar_wait(probe, pph); pph ^= 1; TRM(slot) }
#else
#define PROBE(slot)
#endif
template <int DH>
__global__ void __launch_bounds__(TT_THREADS, 1)   // 96 registers per thread at this CTA size
attn_bwd_tt_kernel(const __grid_constant__ Maps tm, const uint8_t* __restrict__ mask, const float* __restrict__ stats, TailArgs ta,
                   int64_t items, int heads, int L) {
  constexpr int LPAD = 160, DA = DH / 32, KA = 5, HW = 4, CW = 16;
  constexpr int OPA = LPAD * 64, OP_BYTES = DA * OPA, PT_BYTES = KA * 128 * 64;
  constexpr int C_S = 0, C_DV = 160, C_DK = 312, C_T = 464;
  static_assert(C_T + 2 * NT == 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + OP_BYTES;
  uint8_t* sV = sK + OP_BYTES;
  uint8_t* sG = sV + OP_BYTES;             // dO
  uint8_t* sP = sG + OP_BYTES;             // P, then dS, of the main query tile [query][key]
  uint8_t* sX = sP + PT_BYTES;             // 16 x 2 KB store staging, one per element-wise warp
  uint8_t* sT = sX + 16 * 2048;            // P^T, then dS^T, of the tail [key][query], 64 B rows
  uint8_t* sE = sT + 160 * 64;             // ones operand [128][32 B]
  uint8_t* sB = sE + 128 * 32;             // [2][160][32 B] key bias operand
  float* side = reinterpret_cast<float*>(sB + 2 * LPAD * 32);   // [2][3][NT]: row maximum, 1 / denominator, delta of the tail rows
  float* kbias = side + 160;               // [2][160]: 0, -1e30 (masked key) or -inf (past the sequence), for the tail's per-lane keys
  float* dpart = kbias + 320;              // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dpart + 4 * 128);
  uint64_t* qk_full = bars, *qk_empty = bars + 1, *vg_full = bars + 2, *vg_empty = bars + 3, *m_full = bars + 4;
  uint64_t* s_full = bars + 8, *st_full = bars + 9, *p_done = bars + 10, *pt_done = bars + 11, *dp_full = bars + 12, *dv_done = bars + 13;
  uint64_t* dpt_full = bars + 14, *dvt_done = bars + 15, *ds_done = bars + 16, *dst_done = bars + 17, *dq_full = bars + 18;
  uint64_t* fin_full = bars + 19, *c_free = bars + 20, *t_free = bars + 21, *acc_free = bars + 22;
  uint64_t* probe = bars + 23;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.v)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm.g)) : "memory");
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(vg_full, 1); mbar_init(vg_empty, 1);
    mbar_init(&m_full[0], 1); mbar_init(&m_full[1], 1);
    mbar_init(s_full, 1); mbar_init(st_full, 1); mbar_init(p_done, CW); mbar_init(pt_done, 12); mbar_init(dp_full, 1);
    mbar_init(dv_done, 1); mbar_init(dpt_full, 1); mbar_init(dvt_done, 1); mbar_init(ds_done, CW); mbar_init(dst_done, 12);
    mbar_init(dq_full, 1); mbar_init(fin_full, 1); mbar_init(c_free, CW); mbar_init(t_free, DA); mbar_init(acc_free, 4 * DA); mbar_init(probe, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < 128; r += TT_THREADS) {   // ones operand: element 0 of every row = 1.0
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
    // ===================== TMA loads + per-item side data =====================
    // (Q, K) and (V, dO) are released separately: V and dO are last read by the tail's dV^T product in the middle of an item.
    // The same warp stages, one item ahead, the key-bias operand, the tail rows' softmax statistics and their delta
    // (delta_j = sum_d O_jd dO_jd straight from global memory: 24 rows x DH channels); buffer (n + 1) & 1 is free once item n - 1
    // has released Q / K, which is the wait this warp sits in anyway.
    const uint32_t NEG_BIG = 0xF14Au, NEG_INF = 0xFF80u;   // bf16(-1e30), bf16(-inf)
    auto build_side = [&](int64_t it, int mb) {
      const int b = (int)(it / heads), h = (int)(it % heads);
      uint8_t* dstb = sB + mb * (LPAD * 32);
#pragma unroll
      for (int jj = 0; jj < KA; ++jj) {
        const int j = jj * 32 + lane;
        uint32_t val = NEG_INF;
        if (j < L) val = (mask == nullptr || mask[(int64_t)b * L + j] != 0) ? 0u : NEG_BIG;
        const int pc = (j >> 2) & 1;
        *reinterpret_cast<uint4*>(dstb + j * 32 + pc * 16) = make_uint4(val, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dstb + j * 32 + (pc ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
        kbias[mb * 160 + j] = __uint_as_float(val << 16);   // the same bf16 value the rank-1 product adds in the main tile
      }
      float tm_ = 0.f, til = 0.f, tdl = 0.f;
      const int trow = 128 + lane;
      if (lane < NT && trow < L) {
        const float2 st2 = *reinterpret_cast<const float2*>(stats + (((int64_t)b * heads + h) * L + trow) * 2);
        tm_ = st2.x;
        til = st2.y;
        const uint4* po = reinterpret_cast<const uint4*>(ta.o + ((int64_t)b * L + trow) * ta.ldo + h * DH);
        const uint4* pg = reinterpret_cast<const uint4*>(ta.d_o + ((int64_t)b * L + trow) * ta.lddo + h * DH);
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          const uint4 a = po[c], g = pg[c];
          const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 ab = *reinterpret_cast<const __nv_bfloat162*>(&au[e]);
            const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&gu[e]);
            tdl = fmaf(__low2float(ab), __low2float(gb), tdl);
            tdl = fmaf(__high2float(ab), __high2float(gb), tdl);
          }
        }
      }
      if (lane < NT) {
        float* sdw = side + mb * (3 * NT);
        sdw[lane] = tm_;
        sdw[NT + lane] = til;
        sdw[2 * NT + lane] = tdl;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&m_full[mb]);
    };
    auto load_qk = [&](int64_t it) {
      const int b = (int)(it / heads), h = (int)(it % heads);
      mbar_arrive_expect_tx(qk_full, (uint32_t)(2 * OP_BYTES));
#pragma unroll
      for (int a = 0; a < DA; ++a) {
        tma_load_3d(sQ + a * OPA, &tm.q, h * DH + a * 32, 0, b, qk_full);
        tma_load_3d(sK + a * OPA, &tm.k, h * DH + a * 32, 0, b, qk_full);
      }
    };
    if (lane == 0) load_qk(blockIdx.x);
    build_side(blockIdx.x, 0);
    if ((int64_t)blockIdx.x + gridDim.x < items) build_side((int64_t)blockIdx.x + gridDim.x, 1);
    int n = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const uint32_t ph = n & 1;
      if (lane == 0) {
        const int b = (int)(it / heads), h = (int)(it % heads);
        mbar_wait(vg_empty, ph ^ 1);
        mbar_arrive_expect_tx(vg_full, (uint32_t)(2 * OP_BYTES));
#pragma unroll
        for (int a = 0; a < DA; ++a) {
          tma_load_3d(sG + a * OPA, &tm.g, h * DH + a * 32, 0, b, vg_full);
          tma_load_3d(sV + a * OPA, &tm.v, h * DH + a * 32, 0, b, vg_full);
        }
        if (n > 0) {
          mbar_wait(qk_empty, ph ^ 1);
          load_qk(it);
        }
      }
      __syncwarp();
      if (n > 0 && it + gridDim.x < items) build_side(it + gridDim.x, (n + 1) & 1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idS = idesc(LPAD, false, false), idQ = idesc(DH, false, true), idV = idesc(NV, true, true);
      constexpr uint32_t idTs = idesc(NT, false, false), idTq = idesc(NT, true, true), idTv = idesc(NV, true, false);
      int n = 0;
      uint32_t pph = 0;
      (void)pph;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        const uint32_t ph = n & 1;
        const int mb = n & 1;
        const uint8_t* bias = sB + mb * (LPAD * 32);
        mbar_wait(qk_full, ph);
        mbar_wait(&m_full[mb], (n >> 1) & 1);
        TRM(0)
        // S = Q K^T + ones x bias (main tile: queries 0..127)
        mbar_wait(c_free, ph ^ 1);
        tcgen05_fence_after();
        TRM(1)
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sQ + (kk >> 1) * OPA + (kk & 1) * 32)),
                    desc_k64(smem_u32(sK + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
        umma_bf16(tmem_base + C_S, desc_k32(smem_u32(sE)), desc_k32(smem_u32(bias)), idS, 1u);
        umma_commit(s_full);
        PROBE(18)
        TRM(2)
        // tail: S^T = K Q_tail^T, key tiles 0..127 and 128..159 (a key is a TMEM lane there: its bias is added by the thread)
        mbar_wait(t_free, ph ^ 1);
        tcgen05_fence_after();
        TRM(3)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_T + kt * NT, desc_k64(smem_u32(sK + (kk >> 1) * OPA + kt * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sQ + (kk >> 1) * OPA + 128 * 64 + (kk & 1) * 32)), idTs, kk > 0 ? 1u : 0u);
        umma_commit(st_full);
        PROBE(20)
        TRM(4)
        // dP = dO V^T ; dV^T = dO^T P (main tile)
        mbar_wait(vg_full, ph);
        mbar_wait(p_done, ph);
        tcgen05_fence_after();
        TRM(5)
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sG + (kk >> 1) * OPA + (kk & 1) * 32)),
                    desc_k64(smem_u32(sV + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
        umma_commit(dp_full);
        PROBE(22)
        TRM(6)
        mbar_wait(acc_free, ph ^ 1);
        tcgen05_fence_after();
        TRM(7)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem_base + C_DV, desc_mn64(smem_u32(sG + (j * 16) * 64), OPA), desc_mn64(smem_u32(sP + j * 1024), 128 * 64), idV,
                    j > 0 ? 1u : 0u);
        umma_commit(dv_done);   // P may now be overwritten by dS
        PROBE(24)
        TRM(8)
        // tail: dP^T = V dO_tail^T ; dV^T += dO_tail^T P_tail
        mbar_wait(pt_done, ph);
        tcgen05_fence_after();
        TRM(9)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_T + kt * NT, desc_k64(smem_u32(sV + (kk >> 1) * OPA + kt * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sG + (kk >> 1) * OPA + 128 * 64 + (kk & 1) * 32)), idTs, kk > 0 ? 1u : 0u);
        umma_commit(dpt_full);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16(tmem_base + C_DV, desc_mn64(smem_u32(sG + (128 + ks * 16) * 64), OPA), desc_k64(smem_u32(sT + ks * 32)), idTv, 1u);
        umma_commit(dvt_done);  // P^T may now be overwritten by dS^T
        umma_commit(vg_empty);  // V and dO are free
        PROBE(26)
        TRM(10)
        // dQ = dS K ; dK^T = Q^T dS (main tile)
        mbar_wait(ds_done, ph);
        tcgen05_fence_after();
        TRM(11)
#pragma unroll
        for (int ks = 0; ks < LPAD / 16; ++ks)
          umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sP + (ks >> 1) * (128 * 64) + (ks & 1) * 32)),
                    desc_mn64(smem_u32(sK + ks * 1024), OPA), idQ, ks > 0 ? 1u : 0u);
        umma_commit(dq_full);
        PROBE(28)
        TRM(12)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tmem_base + C_DK, desc_mn64(smem_u32(sQ + (j * 16) * 64), OPA), desc_mn64(smem_u32(sP + j * 1024), 128 * 64), idV,
                    j > 0 ? 1u : 0u);
        // tail: dQ_tail^T = K^T dS_tail^T ; dK^T += Q_tail^T dS_tail
        mbar_wait(dst_done, ph);
        tcgen05_fence_after();
        TRM(14)
#pragma unroll
        for (int ks = 0; ks < LPAD / 16; ++ks)
          umma_bf16(tmem_base + C_T, desc_mn64(smem_u32(sK + ks * 1024), OPA), desc_mn64(smem_u32(sT + ks * 1024), 1024), idTq,
                    ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_bf16(tmem_base + C_DK, desc_mn64(smem_u32(sQ + (128 + ks * 16) * 64), OPA), desc_k64(smem_u32(sT + ks * 32)), idTv, 1u);
        umma_commit(fin_full);
        umma_commit(qk_empty);   // also frees this item's side-data buffer
        TRM(15)
        PROBE(31)
      }
    }
  } else {
    // ===================== element-wise phases + epilogues =====================
    // TMEM reads are the resource here (16 B per clock and lane quarter), so the 160 key columns of the main tile are shared
    // out evenly: each of the four warps of a lane quarter owns 40 columns = five 8-key groups (16 B of a P / dS row each).
    const int cw = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter
    const int part = cw >> 2;              // 0..3: key groups 5 * part .. 5 * part + 4
    const int g0 = 5 * part;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rloc = quarter * 32 + lane;  // query row of the main tile
    const int sw64 = (lane >> 1) & 3;
    uint8_t* pRow = sP + rloc * 64;
    auto grp = [&](int g) -> uint8_t* { return pRow + (g >> 2) * (128 * 64) + (((g & 3) ^ sw64) << 4); };
    // the tail's [key x query] tiles: parts 0..2 own query columns 8 * part .. + 7, of key quarter*32 + lane and, in lane quarter 0,
    // also of key 128 + lane
    const bool tail_warp = part < 3;
    const bool tail_hi = tail_warp && quarter == 0;
    uint8_t* tRow0 = sT + rloc * 64;
    uint8_t* tRow1 = sT + (128 + lane) * 64;
    // store staging: this warp's own 2 KB slab, and a 2 KB block of the P tile (rows of this lane quarter in key atom `part`; other
    // warps of the quarter write there in phase A, hence the quarter-wide barrier that ends an item)
    uint8_t* const stage0 = sX + cw * 2048;
    uint8_t* const stage1 = sP + part * (128 * 64) + quarter * 2048;
    constexpr float LOG2E = 1.4426950408889634f;
    int n = 0;
    const int nitems = (int)items;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
      const uint32_t ph = n & 1;
      const float* sd = side + (n & 1) * (3 * NT) + 8 * part;   // this warp's 8 tail columns: [0] maximum, [NT] 1 / denominator, [2 NT] delta
      const float* kb = kbias + (n & 1) * 160;
      const float2 st2 = *reinterpret_cast<const float2*>(stats + ((int64_t)it * L + rloc) * 2);   // it = b * heads + h
      const float m = st2.x, il = st2.y;
      const bool row_grad = m > -1e29f;    // a fully masked row: its logits are constants, dS = 0
      // ---- phase A: P = exp(S - m) / l -> bf16, [query][key] tile ----
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      TRC(0)
      {
        uint32_t r[32], r4[8];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + g0 * 8), r);
        tmem_ld_wait();
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_S + (g0 + 4) * 8), r4);   // in flight under the first four groups
        auto put_p = [&](int g, const uint32_t* rr) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = ex2_fast((__uint_as_float(rr[2 * e]) - m) * LOG2E) * il;
            const float p1 = ex2_fast((__uint_as_float(rr[2 * e + 1]) - m) * LOG2E) * il;
            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            w[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
          *reinterpret_cast<uint4*>(grp(g)) = make_uint4(w[0], w[1], w[2], w[3]);
        };
#pragma unroll
        for (int i = 0; i < 4; ++i) put_p(g0 + i, r + 8 * i);
        tmem_ld_wait();
        put_p(g0 + 4, r4);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_done);
      TRC(1)
      // ---- tail phase A: P^T[key][j] = exp(S^T + bias_key - m_j) / l_j ----
      if (tail_warp) {
        uint32_t tp0[4], tp1[4];   // this key's 8 P^T values per key tile, bf16 pairs
        mbar_wait(st_full, ph);
        tcgen05_fence_after();
        TRC(2)
        uint32_t r[16];
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + 8 * part), r);
        if (tail_hi) tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + NT + 8 * part), r + 8);
        const float kb0 = kb[rloc], kb1 = kb[128 + lane];
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m0 = sd[2 * e], m1 = sd[2 * e + 1], l0 = sd[NT + 2 * e], l1 = sd[NT + 2 * e + 1];
          __nv_bfloat162 hb = __floats2bfloat162_rn(ex2_fast((__uint_as_float(r[2 * e]) + kb0 - m0) * LOG2E) * l0,
                                                    ex2_fast((__uint_as_float(r[2 * e + 1]) + kb0 - m1) * LOG2E) * l1);
          tp0[e] = *reinterpret_cast<uint32_t*>(&hb);
          if (tail_hi) {
            hb = __floats2bfloat162_rn(ex2_fast((__uint_as_float(r[8 + 2 * e]) + kb1 - m0) * LOG2E) * l0,
                                       ex2_fast((__uint_as_float(r[8 + 2 * e + 1]) + kb1 - m1) * LOG2E) * l1);
            tp1[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
        }
        *reinterpret_cast<uint4*>(tRow0 + ((part ^ sw64) << 4)) = make_uint4(tp0[0], tp0[1], tp0[2], tp0[3]);
        if (part == 2) *reinterpret_cast<uint4*>(tRow0 + ((3 ^ sw64) << 4)) = make_uint4(0u, 0u, 0u, 0u);   // query columns 24..31 of the K = 32 steps
        if (tail_hi) {
          *reinterpret_cast<uint4*>(tRow1 + ((part ^ sw64) << 4)) = make_uint4(tp1[0], tp1[1], tp1[2], tp1[3]);
          if (part == 2) *reinterpret_cast<uint4*>(tRow1 + ((3 ^ sw64) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pt_done);
        TRC(3)
      }
      // ---- phase B: dS = P o (dP - delta), written over P ----
      mbar_wait(dp_full, ph);
      tcgen05_fence_after();
      TRC(4)
      {
        // delta_i = sum_j P_ij dP_ij over all keys: the four warps of the lane quarter add their shares through shared memory.
        // dP stays in registers between the two passes (each TMEM column costs read bandwidth once)
        // (32 of the 40 columns: the register file allows 96 per thread at this CTA size, and a spill costs an L2 round trip)
        uint32_t r0[32];
        float psum = 0.f;
        auto dot_p = [&](int g, const uint32_t* rr) {
          const uint4 pq = *reinterpret_cast<const uint4*>(grp(g));
          const uint32_t pu[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&pu[e]);
            psum = fmaf(__low2float(pb), __uint_as_float(rr[2 * e]), psum);
            psum = fmaf(__high2float(pb), __uint_as_float(rr[2 * e + 1]), psum);
          }
        };
        {
          uint32_t r4[8];
          tmem_ld8(tmem_base + lane_off + (uint32_t)(C_S + (g0 + 4) * 8), r4);
          tmem_ld_wait();
          tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + g0 * 8), r0);
          dot_p(g0 + 4, r4);
          tmem_ld_wait();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) dot_p(g0 + i, r0 + 8 * i);
        dpart[part * 128 + rloc] = psum;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        float dl = 0.f;
#pragma unroll
        for (int hp = 0; hp < 4; ++hp) dl += dpart[hp * 128 + rloc];   // same order in every warp of the quarter
        TRC(5)
        mbar_wait(dv_done, ph);   // dV^T = dO^T P has read the P tile
        TRC(6)
        auto put_ds = [&](int g, const uint32_t* rr) {
          uint8_t* addr = grp(g);
          const uint4 pq = *reinterpret_cast<const uint4*>(addr);
          const uint32_t pu[4] = {pq.x, pq.y, pq.z, pq.w};
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&pu[e]);
            float d0 = __low2float(pb) * (__uint_as_float(rr[2 * e]) - dl);
            float d1 = __high2float(pb) * (__uint_as_float(rr[2 * e + 1]) - dl);
            if (!row_grad) d0 = d1 = 0.f;
            __nv_bfloat162 hb = __floats2bfloat162_rn(d0, d1);
            w[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
          *reinterpret_cast<uint4*>(addr) = make_uint4(w[0], w[1], w[2], w[3]);
        };
        {
          uint32_t r4[8];
          tmem_ld8(tmem_base + lane_off + (uint32_t)(C_S + (g0 + 4) * 8), r4);   // the one group that is read twice
#pragma unroll
          for (int i = 0; i < 4; ++i) put_ds(g0 + i, r0 + 8 * i);
          tmem_ld_wait();
          put_ds(g0 + 4, r4);
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_done);
      TRC(7)
      // ---- tail phase B: dS^T[key][j] = P^T o (dP^T - delta_j), written over P^T ----
      if (tail_warp) {
        mbar_wait(dpt_full, ph);
        tcgen05_fence_after();
        TRC(8)
        uint32_t r[16];
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + 8 * part), r);
        if (tail_hi) tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + NT + 8 * part), r + 8);
        // P^T comes back from shared memory (holding it in registers across phase B costs spills, and a spill costs an L2 round trip)
        const uint4 q0 = *reinterpret_cast<const uint4*>(tRow0 + ((part ^ sw64) << 4));
        uint4 q1 = make_uint4(0u, 0u, 0u, 0u);
        if (tail_hi) q1 = *reinterpret_cast<const uint4*>(tRow1 + ((part ^ sw64) << 4));
        const uint32_t tp0[4] = {q0.x, q0.y, q0.z, q0.w}, tp1[4] = {q1.x, q1.y, q1.z, q1.w};
        tmem_ld_wait();
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dl0 = sd[2 * NT + 2 * e], dl1 = sd[2 * NT + 2 * e + 1];
          const bool g0_ = sd[2 * e] > -1e29f, g1_ = sd[2 * e + 1] > -1e29f;
          {
            const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&tp0[e]);
            const float d0 = g0_ ? __low2float(pb) * (__uint_as_float(r[2 * e]) - dl0) : 0.f;
            const float d1 = g1_ ? __high2float(pb) * (__uint_as_float(r[2 * e + 1]) - dl1) : 0.f;
            __nv_bfloat162 hb = __floats2bfloat162_rn(d0, d1);
            w0[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
          if (tail_hi) {
            const __nv_bfloat162 pb = *reinterpret_cast<const __nv_bfloat162*>(&tp1[e]);
            const float d0 = g0_ ? __low2float(pb) * (__uint_as_float(r[8 + 2 * e]) - dl0) : 0.f;
            const float d1 = g1_ ? __high2float(pb) * (__uint_as_float(r[8 + 2 * e + 1]) - dl1) : 0.f;
            __nv_bfloat162 hb = __floats2bfloat162_rn(d0, d1);
            w1[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
        }
        TRC(9)
        mbar_wait(dvt_done, ph);   // dV^T += dO_tail^T P_tail has read the P^T tile
        *reinterpret_cast<uint4*>(tRow0 + ((part ^ sw64) << 4)) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        if (tail_hi) *reinterpret_cast<uint4*>(tRow1 + ((part ^ sw64) << 4)) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        tcgen05_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dst_done);
        TRC(10)
      }
      // ---- dQ epilogue of the main tile: parts 1.. take one 32-channel atom each ----
      int sb = 0;
      const int b = it / heads, h = it - b * heads;
      mbar_wait(dq_full, ph);
      tcgen05_fence_after();
      TRC(11)
      if (part >= 1 && part - 1 < DA) {
        const int c = part - 1;
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(C_S + c * 32), r);
        tmem_ld_wait();
        uint8_t* dst = stage0 + lane * 64;
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
          tma_store_3d(&tm.dq, stage0, h * DH + c * 32, quarter * 32, b);
          bulk_commit();
        }
        sb = 1;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c_free);
      TRC(12)
      // ---- after the item's last product: dQ_tail^T, dV^T, dK^T (lane = head channel, column = query / key) ----
      mbar_wait(fin_full, ph);
      tcgen05_fence_after();
      TRC(13)
      if (quarter * 32 < DH) {
        const int u_lo = part * 2 * KA / HW, u_hi = (part + 1) * 2 * KA / HW;   // 2 or 3 of the 10 units: dV chunks, then dK chunks
        auto unit_addr = [&](int u) -> uint32_t { return tmem_base + lane_off + (uint32_t)((u < KA ? C_DV : C_DK) + (u < KA ? u : u - KA) * 32); };
        auto send_box = [&](uint8_t* dst, const CUtensorMap* omap, int row0) {   // rows >= L are clipped by the store
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(omap, dst, h * DH + quarter * 32, row0, b);
            bulk_commit();
          }
          sb ^= 1;
        };
        uint32_t ra[16], rb[16];
        if (part == 0) {
          uint32_t rt[NT];
          tmem_ld8(tmem_base + lane_off + (uint32_t)C_T, rt);
          tmem_ld8(tmem_base + lane_off + (uint32_t)C_T + 8, rt + 8);
          tmem_ld8(tmem_base + lane_off + (uint32_t)C_T + 16, rt + 16);
          tmem_ld_wait();
          tmem_ld16(unit_addr(u_lo), ra);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(t_free);
            bulk_wait_read<1>();
          }
          __syncwarp();
          uint8_t* dst = sb ? stage1 : stage0;
          stage_transposed<0, NT>(dst, rt, lane);
          send_box(dst, &tm.dq, 128);
        } else {
          tmem_ld16(unit_addr(u_lo), ra);
        }
        TRC(14)
        // each 32-key unit goes through two 16-column reads; one half is staged while the next read is in flight
#pragma unroll 1
        for (int u = u_lo; u < u_hi; ++u) {
          const int c = u < KA ? u : u - KA;
          const bool live = c * 32 < L;
          uint8_t* dst = sb ? stage1 : stage0;
          tmem_ld_wait();
          tmem_ld16(unit_addr(u) + 16, rb);
          if (live) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            stage_transposed<0, 16>(dst, ra, lane);
          }
          tmem_ld_wait();
          if (u + 1 < u_hi) tmem_ld16(unit_addr(u + 1), ra);
          if (live) {
            stage_transposed<16, 16>(dst, rb, lane);
            send_box(dst, u < KA ? &tm.dv : &tm.dk, c * 32);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
      }
      TRC(15)
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      // staging blocks inside the P tile are rewritten by the next item's phase A of OTHER warps of this lane quarter
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      TRC(16)
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#undef TRM
#undef TRC
#undef PROBE

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

template <int DH, int LPAD, int MT, int HW>
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                  int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                  const uint8_t* mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                  cudaStream_t st) {
  constexpr int DA = DH / 32, KA = LPAD / 32;
  constexpr int SMEM = 4 * DA * LPAD * 64 + 2 * KA * 128 * 64 + MT * 128 * 32 + 2 * LPAD * 32 + 128 + HW * 512 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(4 * HW * 4096 <= 2 * KA * 128 * 64, "epilogue staging lives in the P / dS tiles");
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
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<DH, LPAD, MT, false, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_tc_bwd: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t items = batch * heads;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_bwd_tc_kernel<DH, LPAD, MT, false, HW><<<grid, threads_for(HW), SMEM, st>>>(tm, mask, stats, delta, items, heads, L, CrossBwdArgs{0, 1, nullptr});
  return check_launch("attention_bwd_tc");
}

template <int DH>
static int launch_tt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                     const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* mask, const float* stats, int64_t batch, int heads, int L, cudaStream_t st) {
  constexpr int SMEM = tt_smem_bytes<DH>();
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  Maps tm;
  const int cols = heads * DH;
  if (make_map3(&tm.q, q, cols, L, batch, ldq, 160)) return 1;
  if (make_map3(&tm.k, k, cols, L, batch, ldk, 160)) return 1;
  if (make_map3(&tm.v, v, cols, L, batch, ldv, 160)) return 1;
  if (make_map3(&tm.g, d_o, cols, L, batch, lddo, 160)) return 1;
  if (make_map3(&tm.dq, dq, cols, L, batch, lddq, 32)) return 1;
  if (make_map3(&tm.dk, dk, cols, L, batch, lddk, 32)) return 1;
  if (make_map3(&tm.dv, dv, cols, L, batch, lddv, 32)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tt_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_tc_bwd (tail): smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t items = batch * heads;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_bwd_tt_kernel<DH><<<grid, TT_THREADS, SMEM, st>>>(tm, mask, stats, TailArgs{(const bf16*)o, ldo, (const bf16*)d_o, lddo}, items, heads, L);
#ifdef SPA3D_ATTN_TRACE
  static int calls = 0;
  if (++calls == 3) {
    cudaDeviceSynchronize();
    long long hbuf[128];
    cudaMemcpyFromSymbol(hbuf, g_attn_trace, sizeof(hbuf));
    for (int i = 0; i < 128; ++i) if (hbuf[i]) printf("TRACE %d %lld\n", i, hbuf[i] - hbuf[0]);
    fflush(stdout);
  }
#endif
  return check_launch("attention_bwd_tc_tail");
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
  attn_bwd_tc_kernel<DH, 128, 1, true><<<grid, threads_for(2), SMEM, st>>>(tm, mask, stats, delta, items, heads, Lq, CrossBwdArgs{Lk, nchunks, workspace});
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

int attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                     const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                     int Dh, cudaStream_t st) {
  using namespace tb;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al(q) && al(k) && al(v) && al(d_o) && al(dq) && al(dk) && al(dv), "attention_tc_bwd: operands must be 16-byte aligned");
  // 128 < L <= 152: the queries past 128 go through the transposed-tail kernel (SPA3D_ATTN_BWD_TT=0: two-tile kernel, for A/B runs)
  static int tt = -1;
  if (tt < 0) {
    const char* e = getenv("SPA3D_ATTN_BWD_TT");
    tt = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (tt && L > 128 && L <= 128 + NT && o != nullptr && al(o) && ldo % 8 == 0) {
    if (Dh == 96) return launch_tt<96>(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, batch, heads, L, st);
    return launch_tt<64>(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, batch, heads, L, st);
  }
  // SPA3D_ATTN_BWD_HW=2 keeps two element-wise warps per lane quarter everywhere (A/B switch for the 4-warp variant)
  static int hw4 = -1;
  if (hw4 < 0) {
    const char* e = getenv("SPA3D_ATTN_BWD_HW");
    hw4 = (e && atoi(e) == 2) ? 0 : 1;
  }
#define SPA3D_BWD_LAUNCH(DH_, LPAD_, MT_, HW_) \
  launch<DH_, LPAD_, MT_, HW_>(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, stats, delta, batch, heads, L, st)
  if (Dh == 96) {
    if (L <= 128) return hw4 ? SPA3D_BWD_LAUNCH(96, 128, 1, 4) : SPA3D_BWD_LAUNCH(96, 128, 1, 2);
    return hw4 ? SPA3D_BWD_LAUNCH(96, 160, 2, 4) : SPA3D_BWD_LAUNCH(96, 160, 2, 2);
  }
  if (L <= 128) return hw4 ? SPA3D_BWD_LAUNCH(64, 128, 1, 4) : SPA3D_BWD_LAUNCH(64, 128, 1, 2);
  return hw4 ? SPA3D_BWD_LAUNCH(64, 160, 2, 4) : SPA3D_BWD_LAUNCH(64, 160, 2, 2);
#undef SPA3D_BWD_LAUNCH
}

}  // namespace spa3d
