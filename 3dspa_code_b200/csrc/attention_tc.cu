// Short-sequence self-attention core on the 5th-gen tensor cores (tcgen05 + TMEM), bf16:
//   o = softmax(q k^T [+ key mask]) v   per (sequence, head), whole sequence in one CTA pass.
//
// The per-track temporal self-attention (L = T+1 = 151, track_autoencoder_3d.py:182-184), the
// decoder read-out attention (L = 129, :285) and the latent self-attentions (L = 128) all fit one
// tile: S = Q K^T is one UMMA per 16 head channels with N = L (padded to 160) accumulating in TMEM,
// the softmax reads S straight from TMEM with one thread per query row (no shuffles, no online
// rescaling: the whole row is there), P goes back through shared memory as the K-major A operand
// of O = P V, and V is consumed in place as an MN-major B operand (no transpose).
//
// Persistent CTA, one item = (sequence, head):
//   warp 0      TMA: Q,K (3 + 3 boxes of 32 channels x L rows, 64B swizzle) then V, 3-D tensor maps
//               [channels, L, sequences] so rows past the sequence end are zero-filled / clipped.
//   warp 1      MMA issuer: S MMAs (both 128-row query tiles), later the P V MMAs.
//   warps 2..9  softmax + epilogue, thread = query row (tile = warp/4, TMEM lane quarter = warp%4).
// Q/K/V are single-buffered but released early (tcgen05.commit -> mbarrier), so the loads of item
// i+1 run under the softmax / PV / epilogue of item i.
//
// Mask semantics (attention.py:175, flax dot_product_attention): masked logits become a huge negative
// constant, so an all-masked row is uniform; keys beyond L do not exist (-inf).  The mask costs the
// softmax nothing: it is one extra K-step of the S MMA, a rank-1 update ones[q] x bias[k] with
// bias = 0 / -1e30 / -inf per key (a 32-byte-per-row operand pair built by an otherwise idle warp).
// -1e30 instead of finfo.min: any finite logit is absorbed by rounding (ulp(1e30) = 7.6e22), so
// masked keys are the same constant as in the reference, and (s - max) stays exact.
#include "tc_ptx.cuh"

namespace spa3d {
namespace ta {

using namespace tc;

constexpr int THREADS = 320;
constexpr int SM_WARPS = 7;   // softmax / epilogue warps 2..8; warp 9 builds the key-bias operand

__device__ __forceinline__ uint64_t desc_k64(uint32_t addr) {   // K-major, 64B swizzle: 8 rows x 64 B atoms
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ uint64_t desc_k32(uint32_t addr) {   // K-major, 32B swizzle: 8 rows x 32 B atoms (one K-step wide)
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;   // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ uint64_t desc_mn64(uint32_t addr, uint32_t lbo_bytes) {   // MN-major, 64B swizzle
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// kind::f16, fp32 accumulate, bf16 operands, M = 128; b_mn: B operand MN-major
__host__ __device__ constexpr uint32_t idesc_attn(int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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

// CROSS = true: one 128-key chunk of a (sequence, head) of the latents<-tracks cross-attention (track_autoencoder_3d.py:200-201;
// Lq <= 128 queries, Lk keys streamed as ceil(Lk / 128) chunks, one work item each, so every SM takes part): the item's
// unnormalised O (fp32) and its row statistics (max, sum) go to a workspace and attn_cross_merge_kernel combines the chunks
// (flash-style split-K: O = sum_c e^(m_c - M) O_c / sum_c e^(m_c - M) l_c).
struct CrossArgs {
  int Lk, nchunks;
  float* part_o;    // [items][128][DH]
  float* part_ml;   // [items][128][2]
};

// NP = 2: TWO softmax warps per (query tile, TMEM lane quarter), each owning half of the key chunks of the same 32 rows (a TMEM
// lane quarter is readable by every warp with the same warp % 4); row maxima and sums are exchanged through shared memory with
// one 64-thread named barrier each.  The softmax is a chain of fixed-latency instructions per warp (ncu: one issue every ~6
// cycles per warp), so halving the chain and doubling the warps per scheduler shortens the serial S -> softmax -> PV item.
constexpr int threads_for(int MT, int NP) { return NP == 1 ? THREADS : (MT == 2 ? 544 : 384); }

template <int DH, int LPAD, int MT, bool CROSS = false, int NP = 1>   // head width, padded sequence length (multiple of 32), 128-row query tiles
__global__ void __launch_bounds__(threads_for(MT, NP), 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const uint8_t* __restrict__ mask, float* __restrict__ stats, int64_t items, int heads, int L, CrossArgs cx) {
  static_assert(!CROSS || (MT == 1 && LPAD == 128 && NP == 1), "cross-attention items are 128 queries x 128 keys");
  constexpr int NTHREADS = threads_for(MT, NP);
  // warp roles.  NP = 1: warps 2..8 softmax (tile = (warp-2)/4, quarter = warp%4), warp 9 key-bias builder.
  //             NP = 2: warp 2 builder; warps 4..11 tile 0 (part = (warp-4)/4); warps 12 and 16 tile 1, lane quarter 0 (the only
  //             quarter of a second tile with live rows for L <= 160), parts 0 and 1; the other warps only take part in the set-up.
  constexpr int BUILDER = NP == 1 ? 2 + SM_WARPS : 2;
  constexpr int NCOMP = NP == 1 ? SM_WARPS * 32 : (8 + (MT == 2 ? 2 : 0)) * 32;   // threads arriving on s_empty / p_full
  // item -> (sequence b, head h, key chunk ck); self-attention has one chunk (ck = 0) that starts at key 0
  auto item_b = [&](int64_t it) -> int { return CROSS ? (int)(it / ((int64_t)heads * cx.nchunks)) : (int)(it / heads); };
  auto item_h = [&](int64_t it) -> int { return CROSS ? (int)((it / cx.nchunks) % heads) : (int)(it % heads); };
  auto item_k0 = [&](int64_t it) -> int { return CROSS ? (int)(it % cx.nchunks) * 128 : 0; };
  const int Lkeys = CROSS ? cx.Lk : L;
  constexpr int DA = DH / 32;          // 32-channel atoms per operand row
  constexpr int KA = LPAD / 32;        // 32-key atoms per P row
  constexpr int QROWS = MT * 128;
  // DB (two query tiles, NP = 2): K and V are DOUBLE buffered so the loads of item i+1 run under the whole of item i (the Q load
  // already runs under the softmax).  The shared memory for it comes from the padding of the second query tile, which has at most
  // LPAD - 128 = 32 live rows: the Q regions hold LPAD rows instead of 256 and the second tile's P atoms 32 rows instead of 128.  The
  // 128-row MMAs of that tile then read rows that belong to the neighbouring region - harmless, every output row depends on its own
  // operand row only, and the rows past the live ones are never stored - and every such over-read stays inside the allocation.
  constexpr bool DB = NP == 2 && MT == 2;
  constexpr int QR = DB ? LPAD : QROWS;             // rows of one 32-channel Q atom region
  constexpr int NKV = DB ? 2 : 1;
  constexpr int Q_BYTES = DA * QR * 64, K_BYTES = DA * LPAD * 64, P_BYTES = KA * 128 * 64;
  constexpr int P1_ATOM = DB ? 32 * 64 : 128 * 64;  // bytes of one 32-key atom of the SECOND tile's P
  constexpr int P_TOTAL = MT == 1 ? P_BYTES : P_BYTES + KA * P1_ATOM;
  constexpr int S_COL = 0, O_COL = MT * LPAD;
  static_assert(MT * (LPAD + DH) <= 512, "TMEM: S and O accumulators of every query tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                      // [NKV] buffers
  uint8_t* sV = sK + NKV * K_BYTES;                // [NKV] buffers
  uint8_t* sP = sV + NKV * K_BYTES;                // tile 0: [KA][128 rows][64 B]; tile 1: [KA][P1_ATOM]
  uint8_t* sE = sP + P_TOTAL;                      // ones operand [QROWS][32 B], 32B-swizzled, constant
  uint8_t* sB = sE + QROWS * 32;                   // [2][LPAD][32 B] per-item key bias operand
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * LPAD * 32);
  uint64_t* q_full = bars, *q_empty = bars + 1, *k_full = bars + 2, *k_empty = bars + 4, *v_full = bars + 6, *v_empty = bars + 8;   // k / v: [2]
  uint64_t* s_full = bars + 10, *s_empty = bars + 11, *p_full = bars + 12, *o_full = bars + 13;
  uint64_t* m_full = bars + 14, *m_empty = bars + 16;   // [2] each
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);
  float* xch = reinterpret_cast<float*>(bars + 20);    // NP = 2: [max | sum][part][QROWS] exchanged between the two warps of a row
  auto p_tile = [&](int mt) -> uint8_t* { return sP + (mt == 0 ? 0 : P_BYTES); };
  auto p_atom = [&](int mt) -> int { return mt == 0 ? 128 * 64 : P1_ATOM; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, NCOMP); mbar_init(p_full, NCOMP); mbar_init(o_full, 1);
    mbar_init(&m_full[0], 1); mbar_init(&m_full[1], 1); mbar_init(&m_empty[0], 1); mbar_init(&m_empty[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ones operand: element 0 of every row = 1.0 (logical chunk 0 sits at physical chunk (row >> 2) & 1)
  for (int r = threadIdx.x; r < QROWS; r += NTHREADS) {
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
      int n = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1, ++n) {
        const int b = item_b(it), h = item_h(it), k0 = item_k0(it);
        const int kb = DB ? (n & 1) : 0;                               // K / V buffer of this item
        const uint32_t kph = DB ? (uint32_t)((n >> 1) & 1) : ph;      // ... and the phase of its barriers
        mbar_wait(&k_empty[kb], kph ^ 1);
        mbar_arrive_expect_tx(&k_full[kb], (uint32_t)(DA * LPAD * 64));
#pragma unroll
        for (int a = 0; a < DA; ++a) tma_load_3d(sK + kb * K_BYTES + a * (LPAD * 64), &tmK, h * DH + a * 32, k0, b, &k_full[kb]);
        auto load_v = [&]() {
          mbar_wait(&v_empty[kb], kph ^ 1);
          mbar_arrive_expect_tx(&v_full[kb], (uint32_t)(DA * LPAD * 64));
#pragma unroll
          for (int a = 0; a < DA; ++a) tma_load_3d(sV + kb * K_BYTES + a * (LPAD * 64), &tmV, h * DH + a * 32, k0, b, &v_full[kb]);
        };
        // double-buffered: V's buffer was released two items ago, issue it before waiting for the Q region; single-buffered: the V
        // region is released only by the previous item's P V, so Q (released earlier, by its S) must not queue behind it
        if constexpr (DB) load_v();
        mbar_wait(q_empty, ph ^ 1);
        mbar_arrive_expect_tx(q_full, (uint32_t)(DA * LPAD * 64));
#pragma unroll
        for (int a = 0; a < DA; ++a) tma_load_3d(sQ + a * (QR * 64), &tmQ, h * DH + a * 32, 0, b, q_full);
        if constexpr (!DB) load_v();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idS = idesc_attn(LPAD, false), idO = idesc_attn(DH, true);
      uint32_t ph = 0;
      int n = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1, ++n) {
        const int mb = n & 1;
        const int kb = DB ? (n & 1) : 0;
        const uint32_t kph = DB ? (uint32_t)((n >> 1) & 1) : ph;
        mbar_wait(&k_full[kb], kph);
        mbar_wait(q_full, ph);
        mbar_wait(&m_full[mb], (n >> 1) & 1);
        mbar_wait(s_empty, ph ^ 1);   // the softmax of the previous item has finished reading S
        tcgen05_fence_after();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t da = desc_k64(smem_u32(sQ + (kk >> 1) * (QR * 64) + mt * (128 * 64) + (kk & 1) * 32));
            const uint64_t db = desc_k64(smem_u32(sK + kb * K_BYTES + (kk >> 1) * (LPAD * 64) + (kk & 1) * 32));
            umma_bf16(tmem_base + (uint32_t)(S_COL + mt * LPAD), da, db, idS, kk > 0 ? 1u : 0u);
          }
          // S += ones[q] x bias[k]: the key mask as one more K-step
          umma_bf16(tmem_base + (uint32_t)(S_COL + mt * LPAD), desc_k32(smem_u32(sE + mt * (128 * 32))),
                    desc_k32(smem_u32(sB + mb * (LPAD * 32))), idS, 1u);
        }
        umma_commit(s_full);
        umma_commit(q_empty);
        umma_commit(&k_empty[kb]);
        umma_commit(&m_empty[mb]);
        mbar_wait(&v_full[kb], kph);
        mbar_wait(p_full, ph);        // P is in shared memory (and O of the previous item has been drained)
        tcgen05_fence_after();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int ks = 0; ks < LPAD / 16; ++ks) {
            const uint64_t da = desc_k64(smem_u32(p_tile(mt) + (ks >> 1) * p_atom(mt) + (ks & 1) * 32));
            const uint64_t db = desc_mn64(smem_u32(sV + kb * K_BYTES + ks * (16 * 64)), (uint32_t)(LPAD * 64));
            umma_bf16(tmem_base + (uint32_t)(O_COL + mt * DH), da, db, idO, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(o_full);
        umma_commit(&v_empty[kb]);
      }
    }
  } else if (warp == BUILDER) {
    // ===================== key-bias operand builder (one warp, runs one item ahead) =====================
    const uint32_t NEG_BIG = 0xF14Au;   // bf16(-1e30)
    const uint32_t NEG_INF = 0xFF80u;
    int n = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int b = item_b(it), k0 = item_k0(it);
      const int mb = n & 1;
      if (lane == 0) mbar_wait(&m_empty[mb], ((n >> 1) & 1) ^ 1);
      __syncwarp();
      uint8_t* dstb = sB + mb * (LPAD * 32);
      for (int j = lane; j < LPAD; j += 32) {
        uint32_t val = NEG_INF;                                              // key does not exist
        if (k0 + j < Lkeys) val = (mask == nullptr || mask[(int64_t)b * Lkeys + k0 + j] != 0) ? 0u : NEG_BIG;
        const int pc = (j >> 2) & 1;
        *reinterpret_cast<uint4*>(dstb + j * 32 + pc * 16) = make_uint4(val, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dstb + j * 32 + (pc ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&m_full[mb]);
    }
  } else if (NP == 2) {
    // ===================== softmax + epilogue, two warps per (tile, lane quarter) =====================
    const bool tile0 = warp >= 4 && warp < 12;
    const bool is_compute = tile0 || (MT == 2 && (warp == 12 || warp == 16));
    if (is_compute) {
      const int mt = tile0 ? 0 : 1;
      const int part = tile0 ? (warp - 4) >> 2 : (warp == 12 ? 0 : 1);
      const int quarter = warp & 3;          // 0 for the tile-1 warps
      const int rloc = quarter * 32 + lane;
      const int row = mt * 128 + rloc;
      const bool warp_live = mt * 128 + quarter * 32 < L;   // uniform over the pair of warps of this (tile, quarter)
      constexpr int C_SPLIT = (KA + 1) / 2, O_SPLIT = (DA + 1) / 2;
      const int c_lo = part == 0 ? 0 : C_SPLIT, c_hi = part == 0 ? C_SPLIT : KA;
      const int o_lo = part == 0 ? 0 : O_SPLIT, o_hi = part == 0 ? O_SPLIT : DA;
      float* x_max = xch + part * QROWS + row;
      float* x_max_other = xch + (part ^ 1) * QROWS + row;
      float* x_sum = xch + (2 + part) * QROWS + row;
      float* x_sum_other = xch + (2 + (part ^ 1)) * QROWS + row;
      const int bar_id = 1 + mt * 4 + quarter;
      const uint32_t tS = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(S_COL + mt * LPAD);
      const uint32_t tO = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(O_COL + mt * DH);
      uint8_t* pRow = p_tile(mt) + rloc * 64;     // this thread's 64 B in every 32-key atom (tile 1: rloc < 32)
      const int pstride = p_atom(mt);
      const int sw64 = (lane >> 1) & 3;
      constexpr float LOG2E = 1.4426950408889634f;
      uint32_t ph = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1) {
        const int b = item_b(it), h = item_h(it);
        mbar_wait(s_full, ph);
        tcgen05_fence_after();
        float mx = -INFINITY;
        if (warp_live) {
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tS + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(r[i]));
              m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
            }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          *x_max = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          mx = fmaxf(mx, *x_max_other);
          float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tS + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_fast((__uint_as_float(r[2 * i]) - mx) * LOG2E);
              const float p1 = ex2_fast((__uint_as_float(r[2 * i + 1]) - mx) * LOG2E);
              l0 += p0;
              l1 += p1;
              __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
              w[i] = *reinterpret_cast<uint32_t*>(&hb);
            }
            uint8_t* dst = pRow + c * pstride;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          *x_sum = l0 + l1;
        }
        tcgen05_fence_before();
        mbar_arrive(s_empty);
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        mbar_wait(o_full, ph);
        tcgen05_fence_after();
        if (warp_live) {
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");   // the partner's partial sum is in shared memory
          const float inv = 1.f / (*x_sum + *x_sum_other);
          for (int c = o_lo; c < o_hi; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            uint8_t* dst = pRow + c * pstride;    // staging: this warp's rows of P atom c (free once P V has completed)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]) * inv, __uint_as_float(r[j * 8 + 2 * e + 1]) * inv);
                w[e] = *reinterpret_cast<uint32_t*>(&hb);
              }
              *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            for (int c = o_lo; c < o_hi; ++c)
              tma_store_3d(&tmO, p_tile(mt) + c * pstride + quarter * 2048, h * DH + c * 32, mt * 128 + quarter * 32, b);
            bulk_commit();
          }
          if (part == 0 && stats != nullptr && row < L) {
            const int64_t base = (((int64_t)b * heads + h) * L + row) * 2;
            stats[base] = mx;
            stats[base + 1] = inv;
          }
          if (lane == 0) bulk_wait_read<0>();   // the staging rows are P rows of the next item
          __syncwarp();
        }
        tcgen05_fence_before();
      }
    }
  } else {
    // ===================== softmax + epilogue: thread = query row =====================
    const int sw = warp - 2;
    const int mt = sw >> 2;            // query tile
    const int quarter = warp & 3;      // TMEM lane quarter of this warp
    const int row = mt * 128 + quarter * 32 + lane;
    const bool warp_live = mt < MT && mt * 128 + quarter * 32 < L;   // any valid row in this warp
    const uint32_t tS = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(S_COL + mt * LPAD);
    const uint32_t tO = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(O_COL + mt * DH);
    uint8_t* pRow = sP + mt * P_BYTES + (quarter * 32 + lane) * 64;   // this thread's 64 B in every 32-key atom
    const int sw64 = (lane >> 1) & 3;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t ph = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ph ^= 1) {
      const int b = item_b(it), h = item_h(it);
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      float mx = -INFINITY, lsum = 0.f;
      if (warp_live) {
        // pass 1: row maximum (the mask is already in the logits)
#pragma unroll 1
        for (int c = 0; c < KA; ++c) {
          uint32_t r[32];
          tmem_ld32(tS + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[i]));
            m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        }
        // pass 2: probabilities -> bf16 P rows (K-major, 64B-swizzled atoms of 32 keys)
        float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < KA; ++c) {
          uint32_t r[32];
          tmem_ld32(tS + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2_fast((__uint_as_float(r[2 * i]) - mx) * LOG2E);
            const float p1 = ex2_fast((__uint_as_float(r[2 * i + 1]) - mx) * LOG2E);
            l0 += p0;
            l1 += p1;
            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            w[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          uint8_t* dst = pRow + c * (128 * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        lsum = l0 + l1;
      }
      tcgen05_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      // epilogue: O / l -> bf16 -> staging -> TMA store (rows past L are clipped by the tensor map).
      // Staging reuses this warp's own 32 P rows of the first DA atoms (free once P V has completed).
      mbar_wait(o_full, ph);
      tcgen05_fence_after();
      if constexpr (CROSS) {
        // this chunk's unnormalised output row and (max, sum): 384 contiguous bytes per thread
        if (warp_live) {
          float* orow = cx.part_o + ((int64_t)it * 128 + quarter * 32 + lane) * DH;
#pragma unroll
          for (int c = 0; c < DA; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + (uint32_t)(c * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(orow + c * 32 + j * 4) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          *reinterpret_cast<float2*>(cx.part_ml + ((int64_t)it * 128 + quarter * 32 + lane) * 2) = make_float2(mx, lsum);
        }
      } else if (warp_live) {
        const float inv = 1.f / lsum;
#pragma unroll
        for (int c = 0; c < DA; ++c) {
          uint32_t r[32];
          tmem_ld32(tO + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          uint8_t* dst = pRow + c * (128 * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]) * inv, __uint_as_float(r[j * 8 + 2 * e + 1]) * inv);
              w[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < DA; ++c)
            tma_store_3d(&tmO, sP + mt * P_BYTES + c * (128 * 64) + quarter * 2048, h * DH + c * 32, mt * 128 + quarter * 32, b);
          bulk_commit();
        }
        if (stats != nullptr && row < L) {
          const int64_t base = (((int64_t)b * heads + h) * L + row) * 2;
          stats[base] = mx;
          stats[base + 1] = inv;
        }
        if (lane == 0) bulk_wait_read<0>();   // the staging rows are P rows of the next item
        __syncwarp();
      }
      tcgen05_fence_before();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// [cols, L rows, batch] view of a token-major matrix; box = {32 channels, box_rows, 1}, 64B swizzle
static int make_map3(CUtensorMap* map, const void* ptr, int cols, int L, int64_t batch, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SPA3D_REQUIRE(fn != nullptr, "attention_tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPA3D_REQUIRE(r == CUDA_SUCCESS, "attention_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int DH, int LPAD, int MT, int NP = 2>
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                  const uint8_t* mask, float* stats, int64_t batch, int heads, int L, cudaStream_t st) {
  constexpr int DA = DH / 32, KA = LPAD / 32;
  constexpr bool DB = NP == 2 && MT == 2;
  constexpr int SMEM = DA * (DB ? LPAD : MT * 128) * 64 + (DB ? 4 : 2) * DA * LPAD * 64 + KA * 128 * 64 + (MT == 2 ? KA * (DB ? 32 : 128) * 64 : 0) +
                       MT * 128 * 32 + 2 * LPAD * 32 + 256 + 1024 + (NP == 2 ? 4 * MT * 128 * 4 : 0);
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  CUtensorMap tmQ, tmK, tmV, tmO;
  const int cols = heads * DH;
  if (make_map3(&tmQ, q, cols, L, batch, ldq, LPAD)) return 1;
  if (make_map3(&tmK, k, cols, L, batch, ldk, LPAD)) return 1;
  if (make_map3(&tmV, v, cols, L, batch, ldv, LPAD)) return 1;
  if (make_map3(&tmO, o, cols, L, batch, ldo, 32)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, LPAD, MT, false, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_tc: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t items = batch * heads;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_fwd_tc_kernel<DH, LPAD, MT, false, NP><<<grid, threads_for(MT, NP), SMEM, st>>>(tmQ, tmK, tmV, tmO, mask, stats, items, heads, L, CrossArgs{0, 1, nullptr, nullptr});
  return check_launch("attention_fwd_tc");
}

// ---------------------------------------------------------------------------------------------------------------------
// 128 < L <= 152 (the track transformers: L = 151 and L = 129): "transposed tail", one TMEM read of S.
//
// tcgen05.ld moves 16 B per clock per TMEM lane quarter whatever the number of useful lanes (measured with the clock64 trace of
// attention_tc_bwd.cu), so in the two-tile kernel above the 23-row (or 1-row) second query tile costs lane quarter 0 as many read
// clocks as a full tile, on top of its share of the first tile, and every S column is read twice (row maximum, then exp).  Here
//   * the queries past 128 sit on the N side of their own products: S^T = K Q_tail^T is [keys x 24], two M tiles (keys 0..127 and
//     128..159), and O_tail^T = V^T P_tail^T is [channels x 24].  The softmax of those 24 columns runs down the TMEM lanes: column
//     maxima and sums are butterfly shuffles inside a warp plus one exchange through shared memory;
//   * four warps per lane quarter own 40 key columns of the main tile each and keep them in registers between the maximum and the
//     exponentials, so S is read from TMEM once.
// TMEM columns: [0,160) S, [160,256) O, [256,304) S^T (two key tiles), [304,328) O_tail^T.
constexpr int FT_THREADS = 18 * 32;   // TMA + side data, MMA, 16 softmax / epilogue warps
constexpr int FNT = 24;               // tail query columns (L - 128 <= 24)

// Reduce eight per-lane values over the 32 lanes of a warp with 9 shuffles instead of 40 (the shuffle unit issues one warp
// instruction per clock per SM): each exchange step halves the values a lane is responsible for.  On return every lane holds the
// reduction of value (lane >> 2).
template <bool IS_MAX>
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
  auto op = [](float a, float b) { return IS_MAX ? fmaxf(a, b) : a + b; };
  float a4[4], a2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h16 ? v[j] : v[j + 4], keep = h16 ? v[j + 4] : v[j];
    a4[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h8 ? a4[j] : a4[j + 2], keep = h8 ? a4[j + 2] : a4[j];
    a2[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
  const float send = h4 ? a2[0] : a2[1], keep = h4 ? a2[1] : a2[0];
  float c = op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  c = op(c, __shfl_xor_sync(0xffffffffu, c, 2));
  c = op(c, __shfl_xor_sync(0xffffffffu, c, 1));
  return c;
}

template <int DH>
constexpr int ft_smem_bytes() {
  return 4 * (DH / 32) * 160 * 64 + 5 * 128 * 64 + 15 * 2048 + 160 * 64 + 128 * 32 + 2 * 160 * 32 + 10752 + 1024;
}
__host__ __device__ constexpr uint32_t idesc_tt(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

// Development aid (-DSPA3D_ATTN_TRACE): clock64 stamps of the fifth item of CTA 0, MMA thread (slots 0..) and two softmax warps (32.., 64..)
#ifdef SPA3D_ATTN_TRACE
__device__ long long g_fwd_trace[128];
#define FTM(slot) if (blockIdx.x == 0 && n == 4) g_fwd_trace[slot] = clock64();
#define FTC(slot) if (blockIdx.x == 0 && n == 4 && lane == 0 && (cw == 0 || cw == 14)) g_fwd_trace[(cw == 0 ? 32 : 64) + slot] = clock64();
#else
#define FTM(slot)
#define FTC(slot)
#endif
template <int DH>
__global__ void __launch_bounds__(FT_THREADS, 1)
attn_fwd_tt_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, const uint8_t* __restrict__ mask, float* __restrict__ stats, int64_t items,
                   int heads, int L) {
  constexpr int LPAD = 160, DA = DH / 32, KA = 5;
  constexpr int OPA = LPAD * 64, OP_BYTES = DA * OPA, PT_BYTES = KA * 128 * 64;
  constexpr int C_S = 0, C_O = 160, C_T = 352, C_OT = 400;   // O and O_tail^T are double-buffered: [160,352) and [400,448)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + OP_BYTES;
  uint8_t* sV = sK + OP_BYTES;             // [2] buffers
  uint8_t* sP = sV + 2 * OP_BYTES;         // P of the main tile [query][key], 64B-swizzled atoms of 32 keys
  uint8_t* sX = sP + PT_BYTES;             // 15 x 2 KB store staging: one per warp that stores
  uint8_t* sT = sX + 15 * 2048;            // P^T of the tail [key][query], 64 B rows
  uint8_t* sE = sT + 160 * 64;             // ones operand [128][32 B]
  uint8_t* sB = sE + 128 * 32;             // [2][160][32 B] key bias operand
  float* kbias = reinterpret_cast<float*>(sB + 2 * LPAD * 32);   // [3][160] key bias per key, for the tail's per-lane keys
  float* xmax = kbias + 480;               // [4][128] row maxima of the four column shares
  float* xsum = xmax + 512;                // [2][4][128] row sums (by item parity: the epilogue of item n overlaps the softmax of item n+1)
  float* tmax = xsum + 1024;               // [3][4][FNT] column maxima per key quarter (written before the tail's exchange barrier: three items deep)
  float* tsum = tmax + 288;                // [2][4][FNT] column sums per key quarter
  float* tinv = tsum + 192;                // [4][32] 1 / column sum, private to each tail-epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(tinv + 128);
  uint64_t* qk_full = bars, *qk_empty = bars + 1, *v_full = bars + 2, *v_empty = bars + 4, *m_full = bars + 6;
  uint64_t* s_full = bars + 8, *st_full = bars + 9, *p_done = bars + 10, *pt_done = bars + 11, *o_full = bars + 12, *ot_full = bars + 13;
  uint64_t* s_free = bars + 14, *t_free = bars + 15, *o_free = bars + 16 /* [2], one per O buffer */, *ot_free = bars + 18;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); mbar_init(&m_full[i], 1); }
    mbar_init(s_full, 1); mbar_init(st_full, 1); mbar_init(p_done, 16); mbar_init(pt_done, 12); mbar_init(o_full, 1); mbar_init(ot_full, 1);
    mbar_init(s_free, 16); mbar_init(t_free, 12); mbar_init(&o_free[0], 16); mbar_init(&o_free[1], 16); mbar_init(ot_free, DA);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < 128; r += FT_THREADS) {   // ones operand: element 0 of every row = 1.0
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
    // ===================== TMA loads + per-item key bias (operand for the main tile, per-key floats for the tail) =====================
    const uint32_t NEG_BIG = 0xF14Au, NEG_INF = 0xFF80u;   // bf16(-1e30), bf16(-inf)
    auto build_side = [&](int64_t it, int n_of_it) {
      const int b = (int)(it / heads), mb = n_of_it & 1;
      uint8_t* dstb = sB + mb * (LPAD * 32);
      float* kbd = kbias + (n_of_it % 3) * 160;
#pragma unroll
      for (int jj = 0; jj < KA; ++jj) {
        const int j = jj * 32 + lane;
        uint32_t val = NEG_INF;
        if (j < L) val = (mask == nullptr || mask[(int64_t)b * L + j] != 0) ? 0u : NEG_BIG;
        const int pc = (j >> 2) & 1;
        *reinterpret_cast<uint4*>(dstb + j * 32 + pc * 16) = make_uint4(val, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dstb + j * 32 + (pc ^ 1) * 16) = make_uint4(0u, 0u, 0u, 0u);
        kbd[j] = __uint_as_float(val << 16);   // the same bf16 value the rank-1 product adds in the main tile
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
        tma_load_3d(sK + a * OPA, &tmK, h * DH + a * 32, 0, b, qk_full);
        tma_load_3d(sQ + a * OPA, &tmQ, h * DH + a * 32, 0, b, qk_full);
      }
    };
    if (lane == 0) load_qk(blockIdx.x);
    build_side(blockIdx.x, 0);
    if ((int64_t)blockIdx.x + gridDim.x < items) build_side((int64_t)blockIdx.x + gridDim.x, 1);
    int n = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int vb = n & 1;
      if (lane == 0) {
        const int b = (int)(it / heads), h = (int)(it % heads);
        mbar_wait(&v_empty[vb], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vb], (uint32_t)OP_BYTES);
#pragma unroll
        for (int a = 0; a < DA; ++a) tma_load_3d(sV + vb * OP_BYTES + a * OPA, &tmV, h * DH + a * 32, 0, b, &v_full[vb]);
        if (it + gridDim.x < items) {
          mbar_wait(qk_empty, n & 1);   // S and S^T of this item have read Q / K (and the key bias operand)
          load_qk(it + gridDim.x);
        }
      }
      __syncwarp();
      if (it + 2 * (int64_t)gridDim.x < items) build_side(it + 2 * (int64_t)gridDim.x, n + 2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idS = idesc_tt(LPAD, false, false), idO = idesc_tt(DH, false, true);
      constexpr uint32_t idTs = idesc_tt(FNT, false, false), idTo = idesc_tt(FNT, true, true);
      // issue order: S(0) S^T(0) | P V(n), S(n+1) S^T(n+1), V^T P^T(n) | ...: the logits of the next item are ready before the softmax
      // warps have finished with this one, and neither product of an item sits behind a wait that belongs to the other
      auto issue_s = [&](int m) {   // S and S^T of the CTA's m-th item
        const uint32_t ph = m & 1;
        const int mb = m & 1;
        mbar_wait(qk_full, ph);
        mbar_wait(&m_full[mb], (m >> 1) & 1);
        mbar_wait(s_free, ph ^ 1);   // the softmax of the previous item has S in registers
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16(tmem_base + C_S, desc_k64(smem_u32(sQ + (kk >> 1) * OPA + (kk & 1) * 32)),
                    desc_k64(smem_u32(sK + (kk >> 1) * OPA + (kk & 1) * 32)), idS, kk > 0 ? 1u : 0u);
        umma_bf16(tmem_base + C_S, desc_k32(smem_u32(sE)), desc_k32(smem_u32(sB + mb * (LPAD * 32))), idS, 1u);
        umma_commit(s_full);
        // tail: S^T = K Q_tail^T, key tiles 0..127 and 128..159 (a key is a TMEM lane there: its bias is added by the thread)
        mbar_wait(t_free, ph ^ 1);
        tcgen05_fence_after();
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + C_T + kt * FNT, desc_k64(smem_u32(sK + (kk >> 1) * OPA + kt * (128 * 64) + (kk & 1) * 32)),
                      desc_k64(smem_u32(sQ + (kk >> 1) * OPA + 128 * 64 + (kk & 1) * 32)), idTs, kk > 0 ? 1u : 0u);
        umma_commit(st_full);
        umma_commit(qk_empty);
      };
      issue_s(0);
      int n = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        const uint32_t ph = n & 1;
        const int vb = n & 1;
        const uint8_t* vbuf = sV + vb * OP_BYTES;
        FTM(0)
        // O = P V (main tile)
        mbar_wait(&v_full[vb], (n >> 1) & 1);
        mbar_wait(p_done, ph);
        FTM(6)
        // the epilogue of item n - 2 has read this O buffer.  One barrier per buffer: with a single one this wait (completion n - 2)
        // could find the epilogue of item n - 1 complete as well and mistake the phase
        mbar_wait(&o_free[n & 1], ((n >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        FTM(7)
#pragma unroll
        for (int ks = 0; ks < LPAD / 16; ++ks)
          umma_bf16(tmem_base + C_O + (n & 1) * DH, desc_k64(smem_u32(sP + (ks >> 1) * (128 * 64) + (ks & 1) * 32)),
                    desc_mn64(smem_u32(vbuf + ks * 1024), OPA), idO, ks > 0 ? 1u : 0u);
        umma_commit(o_full);
        FTM(8)
        if (it + gridDim.x < items) issue_s(n + 1);
        FTM(4)
        // tail: O_tail^T = V^T P_tail^T
        mbar_wait(pt_done, ph);
        FTM(9)
        // the tail epilogue of the PREVIOUS item has read its O_tail^T: not for the buffer (there are two) but so that the three warps
        // that wait on ot_full only there have seen its previous phase before this product can complete the next one
        if (n >= 1) mbar_wait(ot_free, ph ^ 1);
        tcgen05_fence_after();
        FTM(10)
#pragma unroll
        for (int ks = 0; ks < LPAD / 16; ++ks)
          umma_bf16(tmem_base + C_OT + (n & 1) * FNT, desc_mn64(smem_u32(vbuf + ks * 1024), OPA), desc_mn64(smem_u32(sT + ks * 1024), 1024), idTo,
                    ks > 0 ? 1u : 0u);
        umma_commit(ot_full);
        umma_commit(&v_empty[vb]);
        FTM(11)
      }
    }
  } else {
    // ===================== softmax + epilogues: four warps per TMEM lane quarter, 40 key columns of the main tile each =====================
    const int cw = warp - 2;
    const int quarter = warp & 3;
    const int part = cw >> 2;              // 0..3: key groups 5 * part .. 5 * part + 4 (8 keys = 16 B of a P row each)
    const int g0 = 5 * part;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rloc = quarter * 32 + lane;  // query row of the main tile; key of the tail's first key tile
    const int sw64 = (lane >> 1) & 3;
    uint8_t* pRow = sP + rloc * 64;
    auto grp = [&](int g) -> uint8_t* { return pRow + (g >> 2) * (128 * 64) + (((g & 3) ^ sw64) << 4); };
    // the tail: parts 0..2 own query columns 8 * part .. + 7 of key rloc and, in lane quarter 0, of key 128 + lane as well
    const bool tail_warp = part < 3;
    const bool tail_hi = tail_warp && quarter == 0;
    uint8_t* tRow0 = sT + rloc * 64;
    uint8_t* tRow1 = sT + (128 + lane) * 64;
    // staging slabs: parts 1..3 store one 32-channel atom of the main tile each; part 3 (which has no share of the tail's softmax)
    // of quarters 0..2 also stores the tail
    uint8_t* const slab = sX + ((part >= 1 ? part - 1 : 0) * 4 + quarter) * 2048;
    uint8_t* const slab_t = sX + (12 + quarter) * 2048;
    constexpr float LOG2E = 1.4426950408889634f;
    const int nitems = (int)items;
    int n = 0;
    // The epilogues lag one item behind the softmax: O (and O_tail^T) of item n are drained after the softmax of item n + 1, so the
    // P V and V^T P^T products, and the TMEM reads they have to wait for, never sit between two phases of the same warp.
    float mx_prev = 0.f;
    for (int it = blockIdx.x;; it += gridDim.x, ++n) {
      const bool have = it < nitems;
      const uint32_t ph = n & 1;
      const int par = n & 1;
      float mx = 0.f;
      if (have) {
      // ---- main tile: row maximum and exponentials from ONE read of this warp's 40 S columns ----
      mbar_wait(s_full, ph);
      tcgen05_fence_after();
      FTC(0)
      float lsum;
      {
        uint32_t r[40];
        tmem_ld32p(tmem_base + lane_off + (uint32_t)(C_S + g0 * 8), r);
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_S + (g0 + 4) * 8), r + 32);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        FTC(1)
        float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
        for (int i = 4; i < 40; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(r[i]));
          m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        xmax[part * 128 + rloc] = mx;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        FTC(2)
        mx = fmaxf(fmaxf(xmax[rloc], xmax[128 + rloc]), fmaxf(xmax[256 + rloc], xmax[384 + rloc]));
        if (n > 0) mbar_wait(o_full, ph ^ 1);   // P V of the previous item has read the P tile
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = ex2_fast((__uint_as_float(r[8 * i + 2 * e]) - mx) * LOG2E);
            const float p1 = ex2_fast((__uint_as_float(r[8 * i + 2 * e + 1]) - mx) * LOG2E);
            l0 += p0;
            l1 += p1;
            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            w[e] = *reinterpret_cast<uint32_t*>(&hb);
          }
          *reinterpret_cast<uint4*>(grp(g0 + i)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        lsum = l0 + l1;
        xsum[(par * 4 + part) * 128 + rloc] = lsum;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_done);
      FTC(3)
      // ---- tail: softmax down the lanes of S^T[key][query] ----
      if (tail_warp) {
        mbar_wait(st_full, ph);
        tcgen05_fence_after();
        FTC(4)
        uint32_t r[16];
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + 8 * part), r);
        if (tail_hi) tmem_ld8(tmem_base + lane_off + (uint32_t)(C_T + FNT + 8 * part), r + 8);
        const float* kb = kbias + (n % 3) * 160;
        const float kb0 = kb[rloc], kb1 = kb[128 + lane];
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_free);
        FTC(5)
        float s0[8], s1[8], cm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] = __uint_as_float(r[i]) + kb0;
          s1[i] = tail_hi ? __uint_as_float(r[8 + i]) + kb1 : -INFINITY;
          cm[i] = fmaxf(s0[i], s1[i]);
        }
        float* tmx = tmax + (n % 3) * (4 * FNT);
        {
          const float cmax = warp_reduce8<true>(cm, lane);   // of column 8 * part + (lane >> 2)
          if ((lane & 3) == 0) tmx[quarter * FNT + 8 * part + (lane >> 2)] = cmax;
        }
        asm volatile("bar.sync 5, 384;" ::: "memory");
        FTC(6)
        float cs[8];
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pv[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = 2 * e + u, col = 8 * part + i;
            const float mj = fmaxf(fmaxf(tmx[col], tmx[FNT + col]), fmaxf(tmx[2 * FNT + col], tmx[3 * FNT + col]));
            pv[u] = ex2_fast((s0[i] - mj) * LOG2E);
            pv[2 + u] = tail_hi ? ex2_fast((s1[i] - mj) * LOG2E) : 0.f;
            cs[i] = pv[u] + pv[2 + u];
          }
          __nv_bfloat162 hb = __floats2bfloat162_rn(pv[0], pv[1]);
          w0[e] = *reinterpret_cast<uint32_t*>(&hb);
          hb = __floats2bfloat162_rn(pv[2], pv[3]);
          w1[e] = *reinterpret_cast<uint32_t*>(&hb);
        }
        if (n > 0) mbar_wait(ot_full, ph ^ 1);   // V^T P^T of the previous item has read the P^T tile
        float* tsm = tsum + par * (4 * FNT);
        {
          const float csum = warp_reduce8<false>(cs, lane);
          if ((lane & 3) == 0) tsm[quarter * FNT + 8 * part + (lane >> 2)] = csum;
        }
        *reinterpret_cast<uint4*>(tRow0 + ((part ^ sw64) << 4)) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        if (part == 2) *reinterpret_cast<uint4*>(tRow0 + ((3 ^ sw64) << 4)) = make_uint4(0u, 0u, 0u, 0u);   // query columns 24..31
        if (tail_hi) {
          *reinterpret_cast<uint4*>(tRow1 + ((part ^ sw64) << 4)) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          if (part == 2) *reinterpret_cast<uint4*>(tRow1 + ((3 ^ sw64) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pt_done);
        FTC(7)
      }
      }   // have
      if (n > 0) {
      const int itp = it - (int)gridDim.x;   // the item whose outputs are drained now
      const uint32_t pph = (n - 1) & 1;
      const int ppar = (n - 1) & 1;
      const int b = itp / heads, h = itp - b * heads;
      // ---- tail epilogue: O_tail^T[channel][query] / l_query, transposed into rows 128.. of O ----
      if (part == 3 && quarter * 32 < DH) {
        mbar_wait(ot_full, pph);
        tcgen05_fence_after();
        FTC(10)
        uint32_t r[FNT];
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_OT + ppar * FNT), r);
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_OT + ppar * FNT) + 8, r + 8);
        tmem_ld8(tmem_base + lane_off + (uint32_t)(C_OT + ppar * FNT) + 16, r + 16);
        const float* tsm = tsum + ppar * (4 * FNT);
        const float* tmx = tmax + ((n - 1) % 3) * (4 * FNT);
        // a lone warp issues about one instruction every five clocks, so every lane computes ONE reciprocal column sum and the
        // warp shares them through shared memory
        float* tiv = tinv + quarter * 32;
        float my_inv = 0.f, my_max = 0.f;
        if (lane < FNT) {
          my_inv = 1.f / ((tsm[lane] + tsm[FNT + lane]) + (tsm[2 * FNT + lane] + tsm[3 * FNT + lane]));
          my_max = fmaxf(fmaxf(tmx[lane], tmx[FNT + lane]), fmaxf(tmx[2 * FNT + lane], tmx[3 * FNT + lane]));
          tiv[lane] = my_inv;
        }
        __syncwarp();
        float invj[FNT];
#pragma unroll
        for (int i = 0; i < FNT; ++i) invj[i] = tiv[i];
        if (lane == 0) bulk_wait_read<1>();   // the tail store of the previous item has read slab_t (its main-tile store may be pending)
        __syncwarp();
        FTC(12)
        tmem_ld_wait();
        FTC(13)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ot_free);
#pragma unroll
        for (int i = 0; i < FNT; ++i)   // query 128 + i = row i of the box; this lane's channel = column lane
          *reinterpret_cast<__nv_bfloat16*>(slab_t + i * 64 + ((((lane >> 3) ^ ((i >> 1) & 3))) << 4) + (lane & 7) * 2) =
              __float2bfloat16_rn(__uint_as_float(r[i]) * invj[i]);
        FTC(14)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, slab_t, h * DH + quarter * 32, 128, b);   // rows >= L are clipped
          bulk_commit();
        }
        FTC(11)
        if (quarter == 0 && stats != nullptr && lane < FNT && 128 + lane < L)
          *reinterpret_cast<float2*>(stats + ((int64_t)itp * L + 128 + lane) * 2) = make_float2(my_max, my_inv);
      }
          // ---- main tile epilogue: O / l -> bf16 -> TMA store; parts 1.. take one 32-channel atom each, part 0 the row statistics ----
      // o_full of that item was already waited for in this iteration's softmax (before P was overwritten); waiting again here could
      // see the barrier two phases on (P V of the current item may have completed) and block for good
      if (!have) mbar_wait(o_full, pph);
      tcgen05_fence_after();
      FTC(8)
      {
        const float* xs = xsum + ppar * 512 + rloc;
        const float inv = 1.f / ((xs[0] + xs[128]) + (xs[256] + xs[384]));
        if (part >= 1 && part - 1 < DA) {
          const int c = part - 1;
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_off + (uint32_t)(C_O + ppar * DH + c * 32), r);
          if (lane == 0) {   // the previous item's store has read this slab (part 3 may have its tail store of this item pending)
            if (part == 3 && quarter * 32 < DH) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          __syncwarp();
          tmem_ld_wait();
          uint8_t* dst = slab + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]) * inv, __uint_as_float(r[j * 8 + 2 * e + 1]) * inv);
              w[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(dst + ((j ^ sw64) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmO, slab, h * DH + c * 32, quarter * 32, b);
            bulk_commit();
          }
        } else if (part == 0 && stats != nullptr) {
          *reinterpret_cast<float2*>(stats + ((int64_t)itp * L + rloc) * 2) = make_float2(mx_prev, inv);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ppar]);
      FTC(9)
      }   // n > 0
      mx_prev = mx;
      if (!have) break;
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#undef FTM
#undef FTC

template <int DH>
static int launch_tt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                     const uint8_t* mask, float* stats, int64_t batch, int heads, int L, cudaStream_t st) {
  constexpr int SMEM = ft_smem_bytes<DH>();
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  CUtensorMap tmQ, tmK, tmV, tmO;
  const int cols = heads * DH;
  if (make_map3(&tmQ, q, cols, L, batch, ldq, 160)) return 1;
  if (make_map3(&tmK, k, cols, L, batch, ldk, 160)) return 1;
  if (make_map3(&tmV, v, cols, L, batch, ldv, 160)) return 1;
  if (make_map3(&tmO, o, cols, L, batch, ldo, 32)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tt_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_tc (tail): smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t items = batch * heads;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_fwd_tt_kernel<DH><<<grid, FT_THREADS, SMEM, st>>>(tmQ, tmK, tmV, tmO, mask, stats, items, heads, L);
#ifdef SPA3D_ATTN_TRACE
  static int calls = 0;
  if (++calls == 3) {
    cudaDeviceSynchronize();
    long long hbuf[128];
    cudaMemcpyFromSymbol(hbuf, g_fwd_trace, sizeof(hbuf));
    for (int i = 0; i < 128; ++i) if (hbuf[i]) printf("FTRACE %d %lld\n", i, hbuf[i] - hbuf[0]);
    fflush(stdout);
  }
#endif
  return check_launch("attention_fwd_tc_tail");
}

// combine the key chunks of the cross-attention: one warp per (sequence, head, query row)
template <int DH>
__global__ void attn_cross_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, bf16* __restrict__ o, int64_t ldo,
                                        float* __restrict__ stats, int64_t rows_total, int heads, int Lq, int nchunks) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b * heads + h) * Lq + row
  const int lane = threadIdx.x & 31;
  if (w >= rows_total) return;
  const int row = (int)(w % Lq);
  const int64_t bh = w / Lq;
  const int h = (int)(bh % heads);
  const int64_t b = bh / heads;
  const int64_t item0 = bh * nchunks;
  constexpr float LOG2E = 1.4426950408889634f;
  float M = -INFINITY;
  for (int c = 0; c < nchunks; ++c) M = fmaxf(M, part_ml[((item0 + c) * 128 + row) * 2]);
  float lsum = 0.f, acc[DH / 32];
#pragma unroll
  for (int i = 0; i < DH / 32; ++i) acc[i] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float2 ml = *reinterpret_cast<const float2*>(part_ml + ((item0 + c) * 128 + row) * 2);
    const float wgt = exp2f((ml.x - M) * LOG2E);
    lsum = fmaf(wgt, ml.y, lsum);
    const float* src = part_o + ((item0 + c) * 128 + row) * DH;
#pragma unroll
    for (int i = 0; i < DH / 32; ++i) acc[i] = fmaf(wgt, src[i * 32 + lane], acc[i]);
  }
  const float inv = 1.f / lsum;
  bf16* dst = o + (b * Lq + row) * ldo + h * DH;
#pragma unroll
  for (int i = 0; i < DH / 32; ++i) dst[i * 32 + lane] = __float2bfloat16_rn(acc[i] * inv);
  if (stats != nullptr && lane == 0) {
    stats[w * 2] = M;
    stats[w * 2 + 1] = inv;
  }
}

template <int DH>
static int launch_cross(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                        const uint8_t* mask, float* stats, float* workspace, int64_t batch, int heads, int Lq, int Lk, cudaStream_t st) {
  constexpr int DA = DH / 32, KA = 4;
  constexpr int SMEM = DA * 128 * 64 + 2 * DA * 128 * 64 + KA * 128 * 64 + 128 * 32 + 2 * 128 * 32 + 256 + 1024;
  CUtensorMap tmQ, tmK, tmV;
  const int cols = heads * DH;
  if (make_map3(&tmQ, q, cols, Lq, batch, ldq, 128)) return 1;
  if (make_map3(&tmK, k, cols, Lk, batch, ldk, 128)) return 1;
  if (make_map3(&tmV, v, cols, Lk, batch, ldv, 128)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, 128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_cross_fwd: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const int nchunks = (Lk + 127) / 128;
  const int64_t items = batch * heads * nchunks;
  CrossArgs cx{Lk, nchunks, workspace, workspace + items * 128 * DH};
  const int grid = (int)(items < num_sms() ? items : num_sms());
  attn_fwd_tc_kernel<DH, 128, 1, true><<<grid, THREADS, SMEM, st>>>(tmQ, tmK, tmV, tmQ, mask, nullptr, items, heads, Lq, cx);
  if (int rc = check_launch("attention_cross_fwd")) return rc;
  const int64_t rows_total = batch * heads * Lq;
  attn_cross_merge_kernel<DH><<<(unsigned)((rows_total + 7) / 8), 256, 0, st>>>(cx.part_o, cx.part_ml, (bf16*)o, ldo, stats, rows_total, heads, Lq, nchunks);
  return check_launch("attention_cross_merge");
}

}  // namespace ta

bool attention_fwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                 const void* q, const void* k, const void* v, const void* o) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SPA3D_ATTN_TC");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || dtype != SPA3D_BF16 || Lq != Lk || Lq < 2 || Lq > 160) return false;
  if (Dh != 96 && Dh != 64) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(q) && al16(k) && al16(v) && al16(o) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0;
}

// Cross-attention (Lq <= 128 queries over Lk keys): key chunks split across CTAs + merge.  workspace: attention_cross_workspace_bytes.
bool attention_cross_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo_or_grads) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SPA3D_ATTN_CROSS_TC");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled || dtype != SPA3D_BF16 || Lq < 2 || Lq > 128 || Lk < 1 || Lq == Lk) return false;
  if (Dh != 96 && Dh != 64) return false;
  return ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo_or_grads % 8 == 0;
}

int64_t attention_cross_workspace_floats(int64_t batch, int heads, int Lk, int Dh) {
  const int64_t items = batch * heads * ((Lk + 127) / 128);
  return items * 128 * (Dh + 2);
}

int attention_cross_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                           const uint8_t* key_mask, float* stats, float* workspace, int64_t batch, int heads, int Lq, int Lk, int Dh,
                           cudaStream_t st) {
  using namespace ta;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al16(q) && al16(k) && al16(v) && al16(workspace) && workspace != nullptr, "attention_cross_fwd: operands must be 16-byte aligned");
  if (Dh == 96) return launch_cross<96>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, workspace, batch, heads, Lq, Lk, st);
  return launch_cross<64>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, workspace, batch, heads, Lq, Lk, st);
}

int attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int L, int Dh,
                     cudaStream_t st) {
  using namespace ta;
  // 128 < L <= 152: transposed-tail kernel (SPA3D_ATTN_FWD_TT=0: two-tile kernel, for A/B runs)
  static int tt = -1;
  if (tt < 0) {
    const char* e = getenv("SPA3D_ATTN_FWD_TT");
    tt = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (tt && L > 128 && L <= 128 + FNT) {
    if (Dh == 96) return launch_tt<96>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
    return launch_tt<64>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
  }
  static int np1 = -1;   // A/B switch for measurements: SPA3D_ATTN_NP=1 keeps one softmax warp per (tile, lane quarter)
  if (np1 < 0) {
    const char* e = getenv("SPA3D_ATTN_NP");
    np1 = (e && atoi(e) == 1) ? 1 : 0;
  }
  if (np1) {
    if (Dh == 96) {
      if (L <= 128) return launch<96, 128, 1, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
      return launch<96, 160, 2, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
    }
    if (L <= 128) return launch<64, 128, 1, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
    return launch<64, 160, 2, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
  }
  if (Dh == 96) {
    if (L <= 128) return launch<96, 128, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
    return launch<96, 160, 2>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
  }
  if (L <= 128) return launch<64, 128, 1>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
  return launch<64, 160, 2>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, stats, batch, heads, L, st);
}

}  // namespace spa3d
