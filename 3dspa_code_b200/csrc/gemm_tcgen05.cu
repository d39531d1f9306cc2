// bf16 GEMM on the 5th-gen tensor cores: C = epilogue(A[M,K] . Wt[N,K]^T + bias) (+ residual).
//
// This is the throughput path for every Dense / DenseGeneral of the 3DSPA hot path
// (attention.py:106-107,154-183; track_autoencoder_3d.py:73-115), 90+ % of its FLOPs.
//
// Structure (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor 2D loads of the A (128 x 64) and Wt
//                               (BN x 64) K-blocks into a STAGES-deep 128B-swizzled smem ring,
//                               completion signalled on mbarriers (expect_tx).
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.cta_group::1.kind::f16
//                               (M=128, N=BN, K=16) x4 per K-block, accumulating fp32 in TMEM;
//                               tcgen05.commit releases smem slots and publishes accumulators.
//   warps 2..9  epilogue      : two warps per TMEM lane quarter, each owning half of the tile's
//                               columns.  tcgen05.ld 32 lanes x 32 columns at a time (thread = output
//                               row), software pipelined (the load of chunk c+1 is in flight while
//                               chunk c is processed); bias / tanh-GELU / per-head RMSNorm run on
//                               packed fp32x2 instructions.  Stores never go row-per-thread to
//                               global memory (32 cache lines per instruction): every chunk is
//                               transposed through a swizzled smem slab and leaves either as a TMA
//                               store (bf16 outputs) or as 128-byte-per-row coalesced stores with the
//                               residual read the same coalesced way (fp32 / residual outputs).
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA and masked in
// the epilogue, so M, N need no padding (K % 8 == 0 for the 16-byte TMA stride rule).
//
// Fused per-head RMSNorm (attention.py:166-167): BN/2 is a multiple of the head width, so every
// epilogue thread holds whole heads of its row in registers: q/k normalisation (+ q/sqrt(Dh)) is a
// sum of squares and a scale on values already loaded from TMEM.
#include <type_traits>

#include "tc_ptx.cuh"

namespace spa3d {

namespace tc {

constexpr int tmem_cols_for(int bn) {
  int need = 2 * bn;
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

// epilogue flavours (template parameter EPI)
constexpr int EPI_TMA = 0;     // bf16 output, no residual: swizzled slab -> TMA store
constexpr int EPI_DIRECT = 1;  // fp32 output and/or residual: slab transpose -> coalesced ld/st.global
constexpr int EPI_RMS = 2;     // bf16 output with fused per-head RMSNorm of the q / k columns

// timing experiment only (accumulators are drained, nothing is computed or stored): a compile-time switch so that no
// environment variable can make a product GEMM store nothing
#ifdef SPA3D_GEMM_SKIP_EPI
constexpr bool kSkipEpilogue = true;
#else
constexpr bool kSkipEpilogue = false;
#endif

// Cache policy of the streams a GEMM moves (A/B of four builds in one run, profiles/r02_gemm_cache_policy_ab.txt): the fp32 residual
// read and the fp32 output of the residual flavour are touched once per kernel (0.5 - 1 GB, several times the 126 MB L2) and use
// streaming loads / stores, so they do not push the weight and activation tiles other CTAs are about to re-read out of the L2:
// N=1280 out-projection 0.183 -> 0.168 ms, N=384 MLP_out 0.422 -> 0.412 ms.  NOT for the bf16 side inputs of the backward epilogues
// (the gelu'(z) the forward just wrote is still L2-resident: streaming it measured 0.557 -> 0.613 ms) and NOT for a bf16 residual
// stream either (measured: N=384 out-projection 0.201 -> 0.212 ms, N=1280 0.127 -> 0.142 ms).  An L2 evict_last policy on the
// weight-tile TMA loads changed nothing (the weights never leave the L2 anyway) and is compiled out.
#ifndef SPA3D_TMA_HINTS
#define SPA3D_TMA_HINTS 0
#endif
#ifndef SPA3D_EPI_STREAM
#define SPA3D_EPI_STREAM 1
#endif

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = NUM_EPI_WARPS * 4096;  // one 4 KB slab (or 2 x 2 KB) per epilogue warp
  static constexpr int SCALE_BYTES = 2 * 128 * 4;         // RMSNorm scales (q, k)
  static constexpr int BAR_BYTES = 256;
  static constexpr int BUDGET = 227 * 1024 - EPI_BYTES - SCALE_BYTES - BAR_BYTES - 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + SCALE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

struct EpiParams {
  const float* bias;
  const void* residual;
  int64_t ldr;
  int r_dtype;
  void* C;
  int64_t ldc;
  int c_dtype;
  int act;
  // fused per-head RMSNorm: columns [0,q_cols) use scale_q * q_mul, [q_cols, q_cols+k_cols) use
  // scale_k, the rest is stored as is.
  int q_cols;
  int k_cols;
  const float* scale_q;
  const float* scale_k;
  float q_mul;
  float* rstd_out;   // [M, (q_cols+k_cols)/DH] or null
  int aux_pre;       // EPI_TMA + GELU: second bf16 output through the aux tensor map: 1 = the pre-activation z, 2 = gelu'(z)
  int res_op;        // EPI_DIRECT: 0 = C = f(acc) + residual; 1 = C = acc * gelu'(residual) (residual = saved pre-activation);
                     //             2 = C = acc * residual (residual = gelu'(z) saved by the forward GEMM)
  int atomic_add;    // EPI_DIRECT, fp32 C: C += tile with red.global.add.v4.f32 (split-K weight gradients)
  int x3;            // "bf16 x 3" accurate mode: A and Wt are [rows, 3K] bf16 split operands (hi | mid | lo, x = hi + mid + lo to 24 bits);
                     // the K loop runs the six products hi.hi, hi.mid, mid.hi, hi.lo, mid.mid, lo.hi into ONE fp32 accumulator
                     // (every bf16 x bf16 product is exact in fp32), smallest terms first
  float* colsum;     // EPI_DIRECT: [N] f32, += column sums of the stored values (a bias gradient), or null
};

// bias on one 32-column chunk of an output row held in registers (v = packed pairs).  Lane i holds
// the bias of column i of the chunk (loaded while the MMAs of the tile were still running: a global
// load here would sit on the critical path, L1 is nearly all carved out as shared memory); every
// thread owns a whole row, so the 32 values are broadcast with shuffles.
__device__ __forceinline__ void epi_bias(uint64_t (&v)[16], float breg) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = add2(v[i], pk(__shfl_sync(0xffffffffu, breg, 2 * i), __shfl_sync(0xffffffffu, breg, 2 * i + 1)));
}
// activation (bf16 outputs only: api.cu routes fp32 + GELU to the SIMT path)
__device__ __forceinline__ void epi_act(const EpiParams& ep, uint64_t (&v)[16]) {
  if (ep.act == SPA3D_ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu2_fast(v[i]);
  }
}
// d/dx of flax nn.gelu(approximate=True), hardware tanh
__device__ __forceinline__ float gelu_grad_fast(float x) {
  constexpr float k = 0.7978845608028654f;
  const float x2 = x * x;
  const float t = tanh_fast(k * fmaf(0.044715f * x, x2, x));
  const float du = k * fmaf(3.0f * 0.044715f, x2, 1.0f);
  return fmaf(0.5f * x * (1.0f - t * t), du, 0.5f * (1.0f + t));
}

// 32 bf16 of this lane's row -> 64-byte row of a [32 x 64 B] slab laid out for a SWIZZLE_64B TMA store
__device__ __forceinline__ void slab_store_bf16(uint8_t* slab, int lane, const uint64_t (&v)[16]) {
  const uint32_t srow = smem_u32(slab) + lane * 64;   // explicit shared-space stores (see sts128)
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(srow + ((j ^ sw) << 4), pack_bf16x2(v[4 * j]), pack_bf16x2(v[4 * j + 1]), pack_bf16x2(v[4 * j + 2]), pack_bf16x2(v[4 * j + 3]));
}

// TN = false: C[M,N] = A[M,K] . Wt[N,K]^T (both operands K-major, boxes of 64 K-elements).
// TN = true : C[M,N] = P[K,M]^T . Q[K,N]  (weight gradient dW = dY^T X: the reduction runs over
//             tokens, both operands are token-major = MN-major for the MMA; boxes of 64 tokens x
//             64 columns).  The reduction is split over `splits` CTAs per tile which add their
//             partial tiles into C with vector atomics.
// EW = epilogue warps: 8 (two per TMEM lane quarter, each owning half of the tile's columns) or 16 (four per quarter, a quarter of
// the columns each: twice the warps per scheduler to hide the fixed-latency chains of the GELU epilogue; bf16 TMA-store flavour only).
template <int BN, int EPI, int DH, bool TN, int EW = NUM_EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                    EpiParams ep, int64_t M, int N, int64_t K, int splits) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int TMEM_COLS = tmem_cols_for(BN);
  constexpr int NUM_THREADS = 64 + 32 * EW;
  constexpr int PARTS = EW / 4;      // column parts of a tile (one epilogue warp per TMEM lane quarter and part)
  constexpr int HC = BN / PARTS;     // columns per epilogue warp
  constexpr int NCH = HC / 32;       // 32-column chunks per epilogue warp
  constexpr int SLAB = EW == NUM_EPI_WARPS ? 4096 : 2048;   // staging bytes per epilogue warp (2 x 2 KB, or 1 x 2 KB with 16 warps)
  constexpr int NBUF = SLAB / 2048;
  static_assert(HC % 32 == 0, "tile parts must be whole 32-column chunks");
  static_assert(EW == NUM_EPI_WARPS || (EW == 16 && EPI == EPI_TMA && !TN), "16 epilogue warps: bf16 TMA-store flavour only");
  // (measured and not kept: 12 / 16 epilogue warps for the fp32 / residual flavour - 4 KB slabs per warp cost a smem stage at 256-wide
  //  tiles and the 112-register cap spills: N=384 out-projection 0.254 -> 0.307 ms, gelu-backward 0.556 -> 0.698 ms; and
  //  a 12-warp variant of the head-width-96 RMSNorm epilogue - partial sums of squares exchanged between the
  //  three warps of a lane quarter - ran the ITT QKV shape at 987 instead of 999 TFLOP/s; that epilogue is not latency-bound)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * L::A_BYTES;
  uint8_t* smem_epi = smem + STAGES * L::STAGE_BYTES;                    // [8 warps][4096], 1024-aligned
  float* smem_scale = reinterpret_cast<float*>(smem_epi + L::EPI_BYTES);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + L::EPI_BYTES + L::SCALE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk1 = (int)((K + BK - 1) / BK);                // K-blocks of one operand pair
  const int num_kb_all = ep.x3 ? 6 * nk1 : nk1;
  const int kb_per_split = (num_kb_all + splits - 1) / splits;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t m_tiles = (M + BM - 1) / BM;
  const int64_t tiles_mn = m_tiles * n_tiles;
  const int64_t num_tiles = tiles_mn * splits;            // work items: (output tile, K split)
  // Work item -> (tile, split).  Split-major: the CTAs that run together share one token range and
  // differ in the output tile, so every operand block they stream is reused through L2 by all the
  // tiles of that range (a tile-major order would have each CTA stream its own range: no reuse).
  auto item_tile = [&](int64_t t) -> int64_t { return t % tiles_mn; };
  auto item_split = [&](int64_t t) -> int { return (int)(t / tiles_mn); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (EPI != EPI_DIRECT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EW * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp), address written to smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (EPI == EPI_RMS && threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < DH; i += NUM_THREADS - 64) {
      smem_scale[i] = ep.scale_q ? ep.scale_q[i] * ep.q_mul : 0.f;
      smem_scale[128 + i] = ep.scale_k ? ep.scale_k[i] : 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
#if SPA3D_TMA_HINTS
      const uint64_t pol_w = l2_policy_evict_last();
#endif
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int64_t tile = item_tile(t);
        const int sp = item_split(t);
        const int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb_all, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if constexpr (!TN) {
            int ka = kb * BK, kw = kb * BK;
            if (ep.x3) {
              // pair order (A part, W part): (0,2) (1,1) (2,0) (0,1) (1,0) (0,0) - the 2^-16-sized terms first
              const int pr = kb / nk1, kk = (kb - pr * nk1) * BK;
              const int ai = (0x012010 >> (4 * (5 - pr))) & 15, wi = (0x210100 >> (4 * (5 - pr))) & 15;
              ka = ai * (int)K + kk;
              kw = wi * (int)K + kk;
            }
            tma_load_2d(smem_a + stage * L::A_BYTES, &tmA, ka, m_blk * BM, &full_bar[stage]);
#if SPA3D_TMA_HINTS
            tma_load_2d_hint(smem_b + stage * L::B_BYTES, &tmB, kw, n_blk * BN, &full_bar[stage], pol_w);
#else
            tma_load_2d(smem_b + stage * L::B_BYTES, &tmB, kw, n_blk * BN, &full_bar[stage]);
#endif
          } else {
            // 64 tokens x 64 columns per box; column chunk c lands 8 KB after chunk c-1 (= LBO)
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d(smem_a + stage * L::A_BYTES + c * 8192, &tmA, m_blk * BM + c * 64, kb * BK, &full_bar[stage]);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(smem_b + stage * L::B_BYTES + c * 8192, &tmB, n_blk * BN + c * 64, kb * BK, &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, TN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int sp = item_split(t);
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb_all, kb0 + kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if constexpr (!TN) {
            const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * L::A_BYTES));
            const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * L::B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 32 bytes (16 bf16) along K inside the 128B swizzle atom: +2 in the >>4 field
              umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
            }
          } else {
            const uint64_t da = make_smem_desc_mn(smem_u32(smem_a + stage * L::A_BYTES), 8192);
            const uint64_t db = make_smem_desc_mn(smem_u32(smem_b + stage * L::B_BYTES), 8192);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 tokens = 16 rows of 128 bytes = 2048 bytes: +128 in the >>4 field
              umma_bf16(tmem_d, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which part of the tile's columns (0 .. PARTS-1)
    uint8_t* slab = smem_epi + (size_t)ew * SLAB;
    const bool out_f32 = ep.c_dtype == SPA3D_F32;
    int sbuf = 0;
    int it = 0;
    float4 bq_next[NCH];
    auto load_bias_direct = [&](int64_t tt, float4 (&dst)[NCH]) {
      if (EPI != EPI_DIRECT) return;
      const int nb = (int)(item_tile(tt) % n_tiles);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = nb * BN + half * HC + c * 32 + (lane & 7) * 4;
        dst[c] = (ep.bias && tt < num_tiles && col < N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + col))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_bias_direct(blockIdx.x, bq_next);
    float breg_next[NCH];
    auto load_bias_tma = [&](int64_t tt, float (&dst)[NCH]) {
      if (EPI != EPI_TMA) return;
      const int nb = (int)(item_tile(tt) % n_tiles);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = nb * BN + half * HC + c * 32 + lane;
        dst[c] = (ep.bias && tt < num_tiles && col < N) ? __ldg(ep.bias + col) : 0.f;
      }
    };
    load_bias_tma(blockIdx.x, breg_next);
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int64_t tile = item_tile(t);
      const int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * BM + quarter * 32;
      const int64_t row = (int64_t)row0 + lane;
      const int colbase = n_blk * BN + half * HC;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * HC);

      if constexpr (EPI == EPI_DIRECT) {
        // ---- fp32 / residual outputs -------------------------------------------------------------
        // coalesced view of a 32 x 32 chunk: pass i covers rows 4i..4i+3, lane -> (row 4i + lane/8,
        // 4 columns at (lane%8)*4): every 8 lanes touch one 128-byte (fp32) row segment.
        const int cr = lane >> 3, cc = lane & 7;
        uint4 res[8];
        auto load_res = [&](int c) {
          const int col = colbase + c * 32 + cc * 4;
          const int esz_r = ep.r_dtype == SPA3D_F32 ? 4 : 2;
          const char* rp = reinterpret_cast<const char*>(ep.residual) + (((int64_t)row0 + cr) * ep.ldr + col) * esz_r;
          const int64_t rstep = 4 * ep.ldr * esz_r;
          const int rows_ok = (ep.residual && col < N) ? (int)min((int64_t)32, M - row0) : 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            res[i] = make_uint4(0, 0, 0, 0);
            if (i * 4 + cr < rows_ok) {
              if (ep.r_dtype == SPA3D_F32) {
#if SPA3D_EPI_STREAM
                res[i] = __ldcs(reinterpret_cast<const uint4*>(rp));
#else
                res[i] = *reinterpret_cast<const uint4*>(rp);
#endif
              } else {
                const uint2 h = *reinterpret_cast<const uint2*>(rp);
                res[i].x = h.x;
                res[i].y = h.y;
              }
            }
            rp += rstep;
          }
        };
        // bias of the 4 columns this lane stores in the coalesced pass, one float4 per chunk; the
        // loads for the NEXT tile are issued now (L1 is carved out as smem: a bias load is an L2 round
        // trip, and the epilogue is the pace-setter for these shapes)
        float4 bq[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) bq[c] = bq_next[c];
        load_bias_direct(t + gridDim.x, bq_next);
        if (!kSkipEpilogue) {
          load_res(0);   // in flight while the MMAs of this tile still run
          // pull the residual of this CTA's next tile into L2 (one prefetch per 128-byte line)
          const int64_t tn = item_tile(t + gridDim.x);
          if (ep.residual && t + gridDim.x < num_tiles) {
            const int64_t prow = (int64_t)(tn / n_tiles) * BM + quarter * 32 + lane;
            const int pcol = (int)(tn % n_tiles) * BN + half * HC;
            const int esz = ep.r_dtype == SPA3D_F32 ? 4 : 2;
            if (prow < M) {
              const char* pp = reinterpret_cast<const char*>(ep.residual) + (prow * ep.ldr + pcol) * esz;
              for (int b = 0; b < HC * esz && pcol + b / esz < N; b += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + b));
            }
          }
        }
        mbar_wait(&tfull_bar[as], aphase);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (c == NCH - 1) {
            tcgen05_fence_before();
            mbar_arrive(&tempty_bar[as]);   // every value of this accumulator is in registers
          }
          const int col0 = colbase + c * 32;
          if (col0 < N && !kSkipEpilogue) {
            const uint32_t slab_s = smem_u32(slab);
            const uint32_t srow_s = slab_s + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128(srow_s + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            __syncwarp();
            float4 bcur = bq[0];
#pragma unroll
            for (int c2 = 1; c2 < NCH; ++c2)
              if (c == c2) bcur = bq[c2];
            const int col = col0 + cc * 4;
            const int esz_c = out_f32 ? 4 : 2;
            char* cp = reinterpret_cast<char*>(ep.C) + (((int64_t)row0 + cr) * ep.ldc + col) * esz_c;
            const int64_t cstep = 4 * ep.ldc * esz_c;
            const int rows_ok = col < N ? (int)min((int64_t)32, M - row0) : 0;   // rows of this chunk inside the matrix
            // row 4i + cr sits at chunk cc ^ ((4i + cr) & 7) = cc ^ cr ^ 4(i & 1): odd passes flip 64 bytes
            const uint32_t sp = slab_s + cr * 128 + ((cc ^ cr) << 4);
            const int odd_off = (cc & 4) ? -64 : 64;
            // The eight coalesced passes of a chunk as straight-line code specialised at compile time for the epilogue
            // flavours the model uses (RM: residual 0 none / 1 fp32 / 2 bf16, OP: 0 add / 1 * gelu'(res) / 2 * res,
            // OF: fp32 output, AT: atomic accumulate) - all loads of the chunk, then all the arithmetic, then all stores.
            // With the flavour tested inside the loop every pass was a chain of branches (12 % of the samples resolving
            // branches, one LDS latency exposed per pass).
            auto passes = [&](auto rm_c, auto op_c, auto of_c, auto at_c, auto cs_c) {
              constexpr int RM = decltype(rm_c)::value, OP = decltype(op_c)::value;
              constexpr bool OF = decltype(of_c)::value, AT = decltype(at_c)::value, CSUM = decltype(cs_c)::value;
              float4 a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = lds128(sp + i * 512 + ((i & 1) ? odd_off : 0));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                a[i].x += bcur.x; a[i].y += bcur.y; a[i].z += bcur.z; a[i].w += bcur.w;
                if constexpr (RM != 0) {
                  const uint4 q = res[i];
                  float4 rv;
                  if constexpr (RM == 1) {
                    rv = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
                  } else {
                    rv = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                                     __uint_as_float(q.y & 0xffff0000u));
                  }
                  if constexpr (OP == 0) {
                    a[i].x += rv.x; a[i].y += rv.y; a[i].z += rv.z; a[i].w += rv.w;
                  } else if constexpr (OP == 2) {
                    a[i].x *= rv.x; a[i].y *= rv.y; a[i].z *= rv.z; a[i].w *= rv.w;
                  } else {
                    a[i].x *= gelu_grad_fast(rv.x); a[i].y *= gelu_grad_fast(rv.y);
                    a[i].z *= gelu_grad_fast(rv.z); a[i].w *= gelu_grad_fast(rv.w);
                  }
                }
              }
              if constexpr (CSUM) {
                // column sums of this chunk (bias gradient): 8 rows per lane in registers, the four row groups by shuffle,
                // one vector reduction per 4 columns into the [N] buffer
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (i * 4 + cr < rows_ok) { cs.x += a[i].x; cs.y += a[i].y; cs.z += a[i].z; cs.w += a[i].w; }
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                  cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                  cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
                }
                if (cr == 0 && col < N)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum + col), "f"(cs.x), "f"(cs.y), "f"(cs.z), "f"(cs.w)
                               : "memory");
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                char* cpi = cp + i * cstep;
                if (i * 4 + cr < rows_ok) {
                  if constexpr (AT) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cpi), "f"(a[i].x), "f"(a[i].y), "f"(a[i].z),
                                 "f"(a[i].w)
                                 : "memory");
                  } else if constexpr (OF) {
#if SPA3D_EPI_STREAM
                    __stcs(reinterpret_cast<float4*>(cpi), a[i]);
#else
                    *reinterpret_cast<float4*>(cpi) = a[i];
#endif
                  } else {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(a[i].x, a[i].y), h1 = __floats2bfloat162_rn(a[i].z, a[i].w);
                    *reinterpret_cast<uint2*>(cpi) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                  }
                }
              }
            };
            using std::integral_constant;
            using I0 = integral_constant<int, 0>; using I1 = integral_constant<int, 1>; using I2 = integral_constant<int, 2>;
            using BT = integral_constant<bool, true>; using BF = integral_constant<bool, false>;
            const int rmode = ep.residual == nullptr ? 0 : (ep.r_dtype == SPA3D_F32 ? 1 : 2);
            if (ep.act == SPA3D_ACT_GELU_TANH) {
              // (not produced by the model path: api.cu routes fp32 + GELU to the SIMT kernel) generic form
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 a = lds128(sp + i * 512 + ((i & 1) ? odd_off : 0));
                a.x += bcur.x; a.y += bcur.y; a.z += bcur.z; a.w += bcur.w;
                uint64_t g0 = gelu2_fast(pk(a.x, a.y)), g1 = gelu2_fast(pk(a.z, a.w));
                upk(g0, a.x, a.y);
                upk(g1, a.z, a.w);
                if (rmode) {
                  const uint4 q = res[i];
                  float4 rv = rmode == 1 ? make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w))
                                         : make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u),
                                                       __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
                  a.x += rv.x; a.y += rv.y; a.z += rv.z; a.w += rv.w;
                }
                char* cpi = cp + i * cstep;
                if (i * 4 + cr < rows_ok) {
                  if (out_f32) {
                    *reinterpret_cast<float4*>(cpi) = a;
                  } else {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
                    *reinterpret_cast<uint2*>(cpi) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                  }
                }
              }
            } else if (ep.atomic_add) {
              passes(I0{}, I0{}, BT{}, BT{}, BF{});
            } else if (ep.res_op == 2) {
              if (ep.colsum != nullptr) passes(I2{}, I2{}, BF{}, BF{}, BT{});   // host side: bf16 output only
              else if (out_f32) passes(I2{}, I2{}, BT{}, BF{}, BF{});
              else passes(I2{}, I2{}, BF{}, BF{}, BF{});
            } else if (ep.res_op == 1) {
              if (out_f32) passes(I2{}, I1{}, BT{}, BF{}, BF{}); else passes(I2{}, I1{}, BF{}, BF{}, BF{});
            } else if (rmode == 1) {
              if (out_f32) passes(I1{}, I0{}, BT{}, BF{}, BF{}); else passes(I1{}, I0{}, BF{}, BF{}, BF{});
            } else if (rmode == 2) {
              if (out_f32) passes(I2{}, I0{}, BT{}, BF{}, BF{}); else passes(I2{}, I0{}, BF{}, BF{}, BF{});
            } else {
              if (out_f32) passes(I0{}, I0{}, BT{}, BF{}, BF{}); else passes(I0{}, I0{}, BF{}, BF{}, BF{});
            }
            __syncwarp();   // slab is rewritten by the next chunk
          }
          if (c + 1 < NCH && !kSkipEpilogue) load_res(c + 1);
        }
      } else if constexpr (EPI == EPI_TMA) {
        // ---- bf16 outputs: slab -> TMA store, two 2 KB buffers per warp --------------------------
        float breg[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) breg[c] = breg_next[c];
        load_bias_tma(t + gridDim.x, breg_next);   // next tile's bias: its L2 latency hides under this tile
        mbar_wait(&tfull_bar[as], aphase);
        tcgen05_fence_after();
        // 8 warps: the TMEM load of chunk c+1 is in flight while chunk c is processed (two register buffers); 16 warps: one
        // buffer (112 registers per thread), the other three warps of the scheduler cover the load
        constexpr bool PREFETCH = EW == NUM_EPI_WARPS;
        const int aux_mode = EW == NUM_EPI_WARPS ? ep.aux_pre : 0;   // the training forward's side outputs stay on the 8-warp kernel
        uint32_t r[PREFETCH ? 2 : 1][32];
        if constexpr (PREFETCH) tmem_ld32(tbase, r[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if constexpr (!PREFETCH) tmem_ld32(tbase + (uint32_t)(c * 32), r[0]);
          tmem_ld_wait();
          if (c + 1 < NCH) {
            if constexpr (PREFETCH) tmem_ld32(tbase + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
          } else {   // every value of this accumulator is in registers: hand it back to the MMA warp
            tcgen05_fence_before();
            mbar_arrive(&tempty_bar[as]);
          }
          const int col0 = colbase + c * 32;
          if (col0 < N && !kSkipEpilogue) {
            uint64_t v[16];
            constexpr int RB = PREFETCH ? 1 : 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = pku(r[c & RB][2 * i], r[c & RB][2 * i + 1]);
            if (ep.bias) epi_bias(v, breg[c]);
            if (aux_mode) {   // what the backward pass needs leaves through its own map: z, or gelu'(z) (shares tanh(u) with the activation)
              uint64_t vg[16];
              const bool grad = aux_mode == 2;
              if (grad) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = gelu2_fast_grad(v[i], vg[i]);
              }
              if (lane == 0) bulk_wait_read<NBUF - 1>();
              __syncwarp();
              if (grad) slab_store_bf16(slab + sbuf * 2048, lane, vg);
              else slab_store_bf16(slab + sbuf * 2048, lane, v);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmAux, slab + sbuf * 2048, col0, row0);
                bulk_commit();
              }
              sbuf = (sbuf + 1) % NBUF;
            }
            if (aux_mode != 2) epi_act(ep, v);
            if (lane == 0) bulk_wait_read<NBUF - 1>();   // the store issued two chunks ago has left this buffer
            __syncwarp();
            slab_store_bf16(slab + sbuf * 2048, lane, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, slab + sbuf * 2048, col0, row0);
              bulk_commit();
            }
            sbuf = (sbuf + 1) % NBUF;
          }
        }
      } else {
        // ---- bf16 outputs with per-head RMSNorm: whole heads of this row live in registers ---------
        constexpr int HPW = HC / DH;     // heads per warp
        constexpr int NCHH = DH / 32;    // chunks per head
        static_assert(HC % DH == 0 && DH % 32 == 0, "tile half must hold whole heads");
        const int nq = ep.q_cols + ep.k_cols;
        mbar_wait(&tfull_bar[as], aphase);
        tcgen05_fence_after();
#pragma unroll
        for (int hh = 0; hh < HPW; ++hh) {
          const int gcol = colbase + hh * DH;
          uint32_t r[NCHH][32];
#pragma unroll
          for (int c = 0; c < NCHH; ++c) tmem_ld32(tbase + (uint32_t)(hh * DH + c * 32), r[c]);
          tmem_ld_wait();
          if (hh == HPW - 1) {
            tcgen05_fence_before();
            mbar_arrive(&tempty_bar[as]);
          }
          if (gcol >= N || kSkipEpilogue) continue;
          const int kind = gcol < ep.q_cols ? 0 : (gcol < nq ? 1 : 2);   // q / k / v columns (warp uniform)
          uint64_t rs2 = pk(1.f, 1.f);
          if (kind < 2) {
            uint64_t acc0 = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < NCHH; ++c) {
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const uint64_t x0 = pku(r[c][2 * i], r[c][2 * i + 1]), x1 = pku(r[c][2 * i + 2], r[c][2 * i + 3]);
                acc0 = fma2(x0, x0, acc0);
                acc1 = fma2(x1, x1, acc1);
              }
            }
            float s0, s1;
            upk(add2(acc0, acc1), s0, s1);
            const float rs = rsqrtf((s0 + s1) / (float)DH + kNormEps);
            if (ep.rstd_out && row < M) ep.rstd_out[row * (nq / DH) + gcol / DH] = rs;
            rs2 = pk(rs, rs);
          }
          const float* sc = smem_scale + kind * 128;
#pragma unroll
          for (int c = 0; c < NCHH; ++c) {
            uint64_t v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = pku(r[c][2 * i], r[c][2 * i + 1]);
            if (kind < 2) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 s4 = *reinterpret_cast<const float4*>(sc + c * 32 + g * 4);
                v[2 * g] = mul2(mul2(v[2 * g], rs2), pk(s4.x, s4.y));
                v[2 * g + 1] = mul2(mul2(v[2 * g + 1], rs2), pk(s4.z, s4.w));
              }
            }
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            slab_store_bf16(slab + sbuf * 2048, lane, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, slab + sbuf * 2048, gcol + c * 32, row0);
              bulk_commit();
            }
            sbuf ^= 1;
          }
        }
      }
    }
    if (EPI != EPI_DIRECT) {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

template <int BN, int EPI, int DH, bool TN = false, int EW = NUM_EPI_WARPS>
static int launch(const void* A, int64_t lda, const void* Wt, int64_t ldw, EpiParams ep, int64_t M,
                  int N, int64_t K, cudaStream_t st, void* aux = nullptr, int64_t ld_aux = 0) {
  using L = SmemLayout<BN>;
  CUtensorMap tmA, tmB, tmC, tmAux;
  if (TN) {
    // operands are [K tokens, M] and [K tokens, N], boxes of 64 tokens x 64 columns
    if (make_map(&tmA, A, K, M, lda, 64)) return 1;
    if (make_map(&tmB, Wt, K, N, ldw, 64)) return 1;
  } else {
    const int64_t kcols = ep.x3 ? 3 * K : K;   // split operands are [rows, 3K]
    if (make_map(&tmA, A, M, kcols, lda, BM)) return 1;
    if (make_map(&tmB, Wt, N, kcols, ldw, BN)) return 1;
  }
  if (EPI == EPI_DIRECT) tmC = tmA;  // unused
  else if (make_map_c(&tmC, ep.C, M, N, ep.ldc)) return 1;
  tmAux = tmC;
  const int aux_kind = ep.aux_pre;   // requested by the caller: 0/1 = pre-activation, 2 = activation derivative
  ep.aux_pre = 0;
  if (EPI == EPI_TMA && aux != nullptr) {
    if (make_map_c(&tmAux, aux, M, N, ld_aux)) return 1;
    ep.aux_pre = aux_kind == 2 ? 2 : 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI, DH, TN, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    SPA3D_REQUIRE(e == cudaSuccess, "gemm_tcgen05: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int splits = 1;
  if (TN) {
    // split the token reduction so that every SM has work: about two waves of work items, at
    // least 16 K-blocks (1024 tokens) per item, no empty split
    const int num_kb = (int)((K + BK - 1) / BK);
    int want = (int)((2 * (int64_t)num_sms() + tiles - 1) / tiles);
    int max_by_len = num_kb / 16 > 0 ? num_kb / 16 : 1;
    splits = want < max_by_len ? want : max_by_len;
    if (splits < 1) splits = 1;
    const int per = (num_kb + splits - 1) / splits;
    splits = (num_kb + per - 1) / per;
  }
  const int64_t items = tiles * splits;
  int grid = (int)(items < num_sms() ? items : num_sms());
  gemm_tcgen05_kernel<BN, EPI, DH, TN, EW><<<grid, 64 + 32 * EW, L::TOTAL, st>>>(tmA, tmB, tmC, tmAux, ep, M, N, K, splits);
  return check_launch("gemm_tcgen05");
}

static bool epi_warps16() {   // A/B switch for measurements: SPA3D_GEMM_EW16=0 keeps the 8-warp epilogue (both are product kernels)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPA3D_GEMM_EW16");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

template <int EPI>
static int launch_bn(int bn, const void* A, int64_t lda, const void* Wt, int64_t ldw, const EpiParams& ep,
                     int64_t M, int N, int K, cudaStream_t st, void* aux = nullptr, int64_t ld_aux = 0) {
  if constexpr (EPI == EPI_TMA) {
    // GELU / plain bf16 outputs on 256-wide tiles: four epilogue warps per scheduler (see the kernel's EW)
    // (short K: the epilogue, not the main loop, sets the pace - measured on M=309248, K=384, N=1536 + GELU: 0.41 -> 0.35 ms;
    //  at K=1280 the main loop dominates and the 8-warp kernel with its double-buffered TMEM loads is 3 % faster; the training
    //  forward's two outputs through the single 2 KB slab of a 16-warp layout measured no faster, so they keep 8 warps.)
    if (bn == 256 && aux == nullptr && K <= 768 && epi_warps16())
      return launch<256, EPI, 32, false, 16>(A, lda, Wt, ldw, ep, M, N, K, st, aux, ld_aux);
  }
  switch (bn) {
    case 256: return launch<256, EPI, 32>(A, lda, Wt, ldw, ep, M, N, K, st, aux, ld_aux);
    case 192: return launch<192, EPI, 32>(A, lda, Wt, ldw, ep, M, N, K, st, aux, ld_aux);
    case 128: return launch<128, EPI, 32>(A, lda, Wt, ldw, ep, M, N, K, st, aux, ld_aux);
    default: return launch<64, EPI, 32>(A, lda, Wt, ldw, ep, M, N, K, st, aux, ld_aux);
  }
}

// tile width: the candidate of {256,192,128} with the least padded N (ties -> the widest); 64 for N <= 64
static int pick_bn(int N) {
  if (N <= 64) return 64;
  int best = 256, best_pad = (N + 255) / 256 * 256;
  for (int bn : {192, 128}) {
    int pad = (N + bn - 1) / bn * bn;
    if (pad < best_pad) { best = bn; best_pad = pad; }
  }
  return best;
}

}  // namespace tc

bool gemm_tcgen05_applicable(const void* A, int64_t lda, const void* Wt, int64_t ldw, int64_t M,
                             int N, int K) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(A) && al16(Wt) && (lda % 8 == 0) && (ldw % 8 == 0) && (K % 8 == 0) && (N % 8 == 0) &&
         M > 0 && M < (1ll << 31);
}

bool gemm_tcgen05_rms_applicable(int N, int dh, int q_cols, int k_cols, int c_dtype) {
  if (c_dtype != SPA3D_BF16) return false;
  if (dh != 64 && dh != 96 && dh != 32 && dh != 128) return false;
  return q_cols % dh == 0 && k_cols % dh == 0 && N % dh == 0;
}

int gemm_tcgen05(const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, int act,
                 const void* residual, int64_t ldr, int r_dtype, void* C, int64_t ldc, int c_dtype,
                 int64_t M, int N, int K, const RmsEpilogue* rms, cudaStream_t st, int res_op, void* aux_pre,
                 int64_t ld_aux, int aux_kind, float* colsum) {
  using namespace tc;
  SPA3D_REQUIRE(c_dtype == SPA3D_F32 || c_dtype == SPA3D_BF16, "gemm_tcgen05: bad C dtype");
  SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % (c_dtype == SPA3D_F32 ? 4 : 8) == 0,
                "gemm_tcgen05: C must be 16-byte aligned with 16-byte row pitch");
  if (residual)
    SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % (r_dtype == SPA3D_F32 ? 4 : 8) == 0,
                  "gemm_tcgen05: residual must be 16-byte aligned with 16-byte row pitch");
  if (bias) SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_tcgen05: bias must be 16-byte aligned");
  EpiParams ep{bias, residual, ldr, r_dtype, C, ldc, c_dtype, act, 0, 0, nullptr, nullptr, 1.f, nullptr, 0, res_op, 0, 0};
  if (aux_pre)
    SPA3D_REQUIRE(c_dtype == SPA3D_BF16 && !residual && (reinterpret_cast<uintptr_t>(aux_pre) & 15) == 0 && ld_aux % 8 == 0,
                  "gemm_tcgen05: the pre-activation side output needs a bf16 C, no residual, 16-byte aligned rows");
  if (res_op) SPA3D_REQUIRE(residual != nullptr && r_dtype == SPA3D_BF16, "gemm_tcgen05: res_op needs the saved bf16 pre-activation / derivative");
  ep.aux_pre = aux_kind;
  ep.colsum = colsum;
  if (colsum) SPA3D_REQUIRE(residual != nullptr && res_op == 2 && c_dtype == SPA3D_BF16 && (reinterpret_cast<uintptr_t>(colsum) & 15) == 0,
                            "gemm_tcgen05: fused column sums exist for the saved-derivative backward epilogue (bf16 output) only");
  if (aux_kind == 2) SPA3D_REQUIRE(aux_pre != nullptr && act == SPA3D_ACT_GELU_TANH, "gemm_tcgen05: gelu'(z) side output needs the GELU epilogue");
  if (rms && rms->dh > 0) {
    SPA3D_REQUIRE(c_dtype == SPA3D_BF16 && !bias && !residual && act == 0, "gemm_tcgen05: fused RMSNorm is bf16, no bias/act/residual");
    ep.q_cols = rms->q_cols; ep.k_cols = rms->k_cols;
    ep.scale_q = rms->scale_q; ep.scale_k = rms->scale_k; ep.q_mul = rms->q_mul; ep.rstd_out = rms->rstd_out;
    // every half tile must hold whole heads
    const bool wide = N % 256 == 0;
    switch (rms->dh) {
      case 96: return launch<192, EPI_RMS, 96>(A, lda, Wt, ldw, ep, M, N, K, st);
      case 128: return launch<256, EPI_RMS, 128>(A, lda, Wt, ldw, ep, M, N, K, st);
      case 64: return wide ? launch<256, EPI_RMS, 64>(A, lda, Wt, ldw, ep, M, N, K, st)
                           : launch<128, EPI_RMS, 64>(A, lda, Wt, ldw, ep, M, N, K, st);
      case 32: return wide ? launch<256, EPI_RMS, 32>(A, lda, Wt, ldw, ep, M, N, K, st)
                           : launch<64, EPI_RMS, 32>(A, lda, Wt, ldw, ep, M, N, K, st);
      default: SPA3D_REQUIRE(false, "gemm_tcgen05: fused RMSNorm needs a head width of 32, 64, 96 or 128");
    }
  }
  const int bn = pick_bn(N);
  if (c_dtype == SPA3D_BF16 && !residual) return launch_bn<EPI_TMA>(bn, A, lda, Wt, ldw, ep, M, N, K, st, aux_pre, ld_aux);
  return launch_bn<EPI_DIRECT>(bn, A, lda, Wt, ldw, ep, M, N, K, st);
}

// "bf16 x 3" accurate contraction: C[M,N] (fp32) = A[M,K] . W[N,K]^T + bias (+ residual) with A3 / W3 the [rows, 3K] bf16 splits
// of fp32 operands (spa3d_split3).  Six tcgen05 products per K block into one fp32 accumulator reproduce the fp32 product
// to ~2^-22: the tensor-core form of the "fp32-accumulate" mode (north_star tolerance 1e-4), 6x the MMA work of the bf16 path.
bool gemm_tcgen05_x3_applicable(const void* A3, int64_t lda, const void* W3, int64_t ldw, int64_t M, int N, int K) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(A3) && al16(W3) && lda % 8 == 0 && ldw % 8 == 0 && K % tc::BK == 0 && N % 8 == 0 && M > 0 && M < (1ll << 31);
}

int gemm_tcgen05_x3(const void* A3, int64_t lda, const void* W3, int64_t ldw, const float* bias, const void* residual, int64_t ldr,
                    int r_dtype, float* C, int64_t ldc, int64_t M, int N, int K, cudaStream_t st) {
  using namespace tc;
  SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0, "gemm_x3: C must be 16-byte aligned with 16-byte row pitch");
  if (residual)
    SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % (r_dtype == SPA3D_F32 ? 4 : 8) == 0,
                  "gemm_x3: residual must be 16-byte aligned with 16-byte row pitch");
  if (bias) SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_x3: bias must be 16-byte aligned");
  EpiParams ep{bias, residual, ldr, r_dtype, C, ldc, SPA3D_F32, 0, 0, 0, nullptr, nullptr, 1.f, nullptr, 0, 0, 0, 1};
  return launch_bn<EPI_DIRECT>(pick_bn(N), A3, lda, W3, ldw, ep, M, N, K, st);
}

// Weight gradient on the tensor cores: dW[N,K] += dY[M,N]^T . X[M,K]  (fp32 accumulate into dW;
// the reduction over the M tokens is split across CTAs, partial tiles are added atomically).
bool gemm_tcgen05_dw_applicable(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M,
                                int N, int K, const void* dW, int64_t lddw) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(dY) && al16(X) && al16(dW) && (lddy % 8 == 0) && (ldx % 8 == 0) && (lddw % 4 == 0) &&
         (N % 8 == 0) && (K % 8 == 0) && M > 0 && M < (1ll << 31);
}

int gemm_tcgen05_dw(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                    int64_t M, int N, int K, cudaStream_t st) {
  using namespace tc;
  EpiParams ep{nullptr, nullptr, 0, 0, dW, lddw, SPA3D_F32, 0, 0, 0, nullptr, nullptr, 1.f, nullptr, 0, 0, 1, 0};
  // output tile = 128 rows of dW (N) x BN columns of dW (K); BN a multiple of the 64-column box
  int bn = K <= 64 ? 64 : 256, best_pad = K <= 64 ? 64 : (K + 255) / 256 * 256;
  if (K > 64) {
    for (int c : {192, 128}) {
      int pad = (K + c - 1) / c * c;
      if (pad < best_pad) { bn = c; best_pad = pad; }
    }
  }
  switch (bn) {
    case 256: return launch<256, EPI_DIRECT, 32, true>(dY, lddy, X, ldx, ep, N, K, M, st);
    case 192: return launch<192, EPI_DIRECT, 32, true>(dY, lddy, X, ldx, ep, N, K, M, st);
    case 128: return launch<128, EPI_DIRECT, 32, true>(dY, lddy, X, ldx, ep, N, K, M, st);
    default: return launch<64, EPI_DIRECT, 32, true>(dY, lddy, X, ldx, ep, N, K, M, st);
  }
}

}  // namespace spa3d
