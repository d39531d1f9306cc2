// bf16 GEMM on the 5th-gen tensor cores: C = epilogue(A[M,K] . Wt[N,K]^T + bias) (+ residual).
//
// This is the throughput path for every Dense / DenseGeneral of the 3DSPA hot path
// (attention.py:106-107,154-183; track_autoencoder_3d.py:73-115), 90+ % of its FLOPs.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor 2D loads of the A (128 x 64) and Wt
//                               (BN x 64) K-blocks into a STAGES-deep 128B-swizzled smem ring,
//                               completion signalled on mbarriers (expect_tx).
//   warp 1      MMA issuer    : one thread issues tcgen05.mma.cta_group::1.kind::f16
//                               (M=128, N=BN, K=16) x4 per K-block, accumulating fp32 in TMEM;
//                               tcgen05.commit releases smem slots and publishes accumulators.
//   warps 2..5  epilogue      : tcgen05.ld 32 lanes x 32 columns at a time (thread = output row),
//                               bias / tanh-GELU / residual / per-head RMSNorm in registers, then
//                               each warp stages its 32 rows x 128 B in a 128B-swizzled smem slab
//                               (conflict-free st.shared.v4) and one lane issues a TMA store
//                               (cp.async.bulk.tensor.2d.global.shared), double buffered.  A
//                               row-per-thread st.global would touch 32 cache lines per
//                               instruction and made the epilogue, not the MMAs, the pace-setter.
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA and masked in
// the epilogue, so M, N need no padding (K % 8 == 0 for the 16-byte TMA stride rule).
//
// Operand reuse (CL = 2): the kernel is L2->smem bound, not MMA bound (a 128 x 256 tile needs 48 KB
// of operands per 512 MMA cycles).  With CL = 2 the two CTAs of a cluster work on vertically
// adjacent M tiles of the same N tile; each loads half of the weight tile and multicasts it to
// both (cp.async.bulk.tensor ... .multicast::cluster), so weight traffic from L2 halves.  The smem
// slot of a stage is released by BOTH CTAs' MMA warps (tcgen05.commit ... multicast::cluster).
//
// Fused per-head RMSNorm (attention.py:166-167): with BN a multiple of the head width every
// accumulator row holds whole heads in one thread, so q/k normalisation (+ q/sqrt(Dh)) is two
// passes over TMEM inside the epilogue and the separate in-place pass over q,k disappears.
#include <cuda.h>

#include "common.cuh"

namespace spa3d {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int NUM_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                               uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
// (cute::UMMA::SmemDescriptor: start>>4 @0, LBO>>4 @16, SBO>>4 @32, version=1 @46, layout @61)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: c=f32, a=b=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int tmem_cols_for(int bn) {
  int need = 2 * bn;
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)
  static constexpr int SCALE_BYTES = 2 * 128 * 4;  // RMSNorm scales (q, k)
  static constexpr int BUDGET = 227 * 1024 - EPI_BYTES - SCALE_BYTES - 256 - 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
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
  // scale_k, the rest is stored as is.  rms_dh == 0 disables it.
  int rms_dh;
  int q_cols;
  int k_cols;
  const float* scale_q;
  const float* scale_k;
  float q_mul;
  float* rstd_out;   // [M, (q_cols+k_cols)/rms_dh] or null
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// flax nn.gelu(approximate=True) with the hardware tanh (bf16 outputs only)
__device__ __forceinline__ float gelu_fast(float x) {
  const float k = 0.7978845608028654f;
  float u = k * fmaf(0.044715f * x, x * x, x);
  return 0.5f * x * (1.0f + tanh_fast(u));
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// bias / GELU / residual on 8 consecutive columns of one output row (registers only)
__device__ __forceinline__ void epi_math8(const EpiParams& ep, float (&v)[8], int64_t row, int col, bool row_ok) {
  if (ep.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (ep.act == SPA3D_ACT_GELU_TANH) {
    if (ep.c_dtype == SPA3D_BF16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_fast(v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_tanh(v[i]);
    }
  }
  if (ep.residual && row_ok) {
    if (ep.r_dtype == SPA3D_F32) {
      const float* rp = reinterpret_cast<const float*>(ep.residual) + row * ep.ldr + col;
      const float4 r0 = *reinterpret_cast<const float4*>(rp);
      const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    } else {
      const uint4 rr = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.residual) + row * ep.ldr + col);
      const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        v[2 * i] += __low2float(h);
        v[2 * i + 1] += __high2float(h);
      }
    }
  }
}

template <int BN, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, EpiParams ep, int64_t M, int N, int K) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int TMEM_COLS = tmem_cols_for(BN);
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in every CTA of a cluster (multicast writes land at the same smem offset)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * L::A_BYTES;
  uint8_t* smem_epi = smem + STAGES * L::STAGE_BYTES;                    // [4 warps][2][4096], 1024-aligned
  float* smem_scale = reinterpret_cast<float*>(smem_epi + L::EPI_BYTES);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + L::EPI_BYTES + L::SCALE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t m_tiles = (M + BM - 1) / BM;
  const int64_t m_groups = (m_tiles + CL - 1) / CL;
  const int64_t num_groups = m_groups * n_tiles;      // a group = CL vertically adjacent tiles
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const int64_t cluster_id = blockIdx.x / CL;
  const int64_t num_clusters = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL);   // one tcgen05.commit from every CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
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
  if (ep.rms_dh > 0 && threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < ep.rms_dh; i += NUM_THREADS - 64) {
      smem_scale[i] = ep.scale_q ? ep.scale_q[i] * ep.q_mul : 0.f;
      smem_scale[128 + i] = ep.scale_k ? ep.scale_k[i] : 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast lands
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t g = cluster_id; g < num_groups; g += num_clusters) {
        const int m_blk = (int)(g / n_tiles) * CL + (int)crank, n_blk = (int)(g % n_tiles);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(smem_a + stage * L::A_BYTES, &tmA, kb * BK, m_blk * BM, &full_bar[stage]);
          if (CL == 1) {
            tma_load_2d(smem_b + stage * L::B_BYTES, &tmB, kb * BK, n_blk * BN, &full_bar[stage]);
          } else {
            // this CTA fetches rows [crank*BN/CL, (crank+1)*BN/CL) of the weight tile for everyone
            tma_load_2d_mc(smem_b + stage * L::B_BYTES + crank * (L::B_BYTES / CL), &tmB, kb * BK,
                           n_blk * BN + (int)crank * (BN / CL), &full_bar[stage], MC_MASK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t g = cluster_id; g < num_groups; g += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * L::A_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * L::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128B swizzle atom: +2 in the >>4 field
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          // smem slot reusable (in every CTA that multicasts into it) once these MMAs have read it
          if (CL == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_mc(&empty_bar[stage], MC_MASK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    uint8_t* slab = smem_epi + (size_t)quarter * 2 * 4096;
    const bool out_f32 = ep.c_dtype == SPA3D_F32;
    const int box_cols = out_f32 ? 32 : 64;     // 128 bytes of output per row and box
    int sbuf = 0;
    int it = 0;
    for (int64_t g = cluster_id; g < num_groups; g += num_clusters, ++it) {
      const int m_blk = (int)(g / n_tiles) * CL + (int)crank, n_blk = (int)(g % n_tiles);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      const int row0 = m_blk * BM + quarter * 32;
      const int64_t row = (int64_t)row0 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
      // fused per-head RMSNorm, pass 1: 1/rms of every normalised head of this row
      float rstd_h[8];
      const int dh = ep.rms_dh;
      const int nq = ep.q_cols + ep.k_cols;
      if (dh > 0) {
#pragma unroll 1
        for (int hh = 0; hh * dh < BN; ++hh) {
          const int gcol = n_blk * BN + hh * dh;
          float rs = 1.f;
          if (gcol < nq && gcol < N) {
            float ss = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < dh; c0 += 32) {
              uint32_t r[32];
              tmem_ld32(taddr + (uint32_t)(hh * dh + c0), r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float x = __uint_as_float(r[i]);
                ss = fmaf(x, x, ss);
              }
            }
            rs = rsqrtf(ss / (float)dh + kNormEps);
            if (ep.rstd_out && row_ok) ep.rstd_out[row * (nq / dh) + gcol / dh] = rs;
          }
          rstd_h[hh & 7] = rs;
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n_blk * BN + c0;
        if (col0 >= N) break;
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        const int cb = c0 % box_cols;  // column offset inside the current 128-byte box
        if (cb == 0) {
          // the buffer we are about to fill was handed to a TMA store two boxes ago
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
        }
        float mul = 1.f;
        const float* sc = nullptr;
        if (dh > 0 && col0 < nq) {
          mul = rstd_h[(c0 / dh) & 7];
          sc = smem_scale + (col0 < ep.q_cols ? 0 : 128) + (c0 % dh);
        }
        uint8_t* srow = slab + sbuf * 4096 + lane * 128;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = __uint_as_float(r[gq * 8 + i]);
            v[i] = sc ? x * mul * sc[gq * 8 + i] : x;
          }
          const int col = col0 + gq * 8;
          if (col < N) epi_math8(ep, v, row, col, row_ok);
          if (out_f32) {
            // 8 floats = two 16-byte chunks j = 2*gq, 2*gq+1 of the 128-byte row; swizzle j ^ (row & 7)
            const int j0 = 2 * gq, j1 = 2 * gq + 1;
            *reinterpret_cast<float4*>(srow + ((j0 ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(srow + ((j1 ^ (lane & 7)) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            const int j = (cb >> 3) + gq;  // 8 bf16 = one 16-byte chunk
            *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        const bool box_done = (cb + 32 == box_cols) || (c0 + 32 >= BN) || (col0 + 32 >= N);
        if (box_done) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, slab + sbuf * 4096, col0 - cb, row0);
            bulk_commit();
          }
          sbuf ^= 1;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into its smem
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// output map: [M, N] of bf16 / f32, box = {128 bytes of columns, 32 rows}, 128B swizzle
static int make_map_c(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int c_dtype) {
  EncodeTiledFn fn = get_encode_fn();
  SPA3D_REQUIRE(fn != nullptr, "gemm_tcgen05: cuTensorMapEncodeTiled not available from the driver");
  const int esz = c_dtype == SPA3D_F32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, c_dtype == SPA3D_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPA3D_REQUIRE(r == CUDA_SUCCESS, "gemm_tcgen05: cuTensorMapEncodeTiled (C) failed (%d)", (int)r);
  return 0;
}

// 2D bf16 row-major [rows, cols] with row pitch ld (elements); box = {64 cols, box_rows}
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SPA3D_REQUIRE(fn != nullptr, "gemm_tcgen05: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPA3D_REQUIRE(r == CUDA_SUCCESS, "gemm_tcgen05: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int g_cluster = 2;  // weight-tile multicast width (1 = off); SPA3D_GEMM_CLUSTER overrides

static int cluster_width() {
  static bool init = false;
  if (!init) {
    const char* e = getenv("SPA3D_GEMM_CLUSTER");
    if (e) g_cluster = atoi(e) == 1 ? 1 : 2;
    init = true;
  }
  return g_cluster;
}

template <int BN, int CL>
static int launch(const void* A, int64_t lda, const void* Wt, int64_t ldw, const EpiParams& ep,
                  int64_t M, int N, int K, cudaStream_t st) {
  using L = SmemLayout<BN>;
  CUtensorMap tmA, tmB, tmC;
  if (make_map(&tmA, A, M, K, lda, BM)) return 1;
  if (make_map(&tmB, Wt, N, K, ldw, BN / CL)) return 1;
  if (make_map_c(&tmC, ep.C, M, N, ep.ldc, ep.c_dtype)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    SPA3D_REQUIRE(e == cudaSuccess, "gemm_tcgen05: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t m_tiles = (M + BM - 1) / BM;
  int64_t groups = ((m_tiles + CL - 1) / CL) * ((N + BN - 1) / BN);
  int max_clusters = num_sms() / CL;
  int clusters = (int)(groups < max_clusters ? groups : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, CL>, tmA, tmB, tmC, ep, M, N, K);
  SPA3D_REQUIRE(e == cudaSuccess, "gemm_tcgen05 launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_tcgen05");
}

template <int BN>
static int launch_cl(const void* A, int64_t lda, const void* Wt, int64_t ldw, const EpiParams& ep,
                     int64_t M, int N, int K, cudaStream_t st) {
  // multicast pays only when there are at least two M tiles to pair up
  if (cluster_width() == 2 && M > BM) return launch<BN, 2>(A, lda, Wt, ldw, ep, M, N, K, st);
  return launch<BN, 1>(A, lda, Wt, ldw, ep, M, N, K, st);
}

}  // namespace tc

bool gemm_tcgen05_applicable(const void* A, int64_t lda, const void* Wt, int64_t ldw, int64_t M,
                             int N, int K) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(A) && al16(Wt) && (lda % 8 == 0) && (ldw % 8 == 0) && (K % 8 == 0) && (N % 8 == 0) &&
         M > 0 && M < (1ll << 31);
}

bool gemm_tcgen05_rms_applicable(int N, int dh, int q_cols, int k_cols) {
  if (dh != 64 && dh != 96 && dh != 32 && dh != 128) return false;
  return q_cols % dh == 0 && k_cols % dh == 0 && N % dh == 0;
}

int gemm_tcgen05(const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, int act,
                 const void* residual, int64_t ldr, int r_dtype, void* C, int64_t ldc, int c_dtype,
                 int64_t M, int N, int K, const RmsEpilogue* rms, cudaStream_t st) {
  using namespace tc;
  SPA3D_REQUIRE(c_dtype == SPA3D_F32 || c_dtype == SPA3D_BF16, "gemm_tcgen05: bad C dtype");
  SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % (c_dtype == SPA3D_F32 ? 4 : 8) == 0,
                "gemm_tcgen05: C must be 16-byte aligned with 16-byte row pitch");
  if (residual)
    SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % (r_dtype == SPA3D_F32 ? 4 : 8) == 0,
                  "gemm_tcgen05: residual must be 16-byte aligned with 16-byte row pitch");
  if (bias) SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_tcgen05: bias must be 16-byte aligned");
  EpiParams ep{bias, residual, ldr, r_dtype, C, ldc, c_dtype, act, 0, 0, 0, nullptr, nullptr, 1.f, nullptr};
  if (rms && rms->dh > 0) {
    ep.rms_dh = rms->dh; ep.q_cols = rms->q_cols; ep.k_cols = rms->k_cols;
    ep.scale_q = rms->scale_q; ep.scale_k = rms->scale_k; ep.q_mul = rms->q_mul; ep.rstd_out = rms->rstd_out;
    // tile width must hold whole heads
    const int dh = rms->dh;
    if (dh == 96) return launch_cl<192>(A, lda, Wt, ldw, ep, M, N, K, st);
    if (dh == 64 || dh == 128 || dh == 32) {
      if (N % 256 == 0) return launch_cl<256>(A, lda, Wt, ldw, ep, M, N, K, st);
      return launch_cl<128>(A, lda, Wt, ldw, ep, M, N, K, st);
    }
    SPA3D_REQUIRE(false, "gemm_tcgen05: fused RMSNorm needs a head width of 32, 64, 96 or 128");
  }
  // tile width: the widest of {256,192,128} that divides N, else the narrowest tile covering N
  if (N % 256 == 0) return launch_cl<256>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N % 192 == 0) return launch_cl<192>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N % 128 == 0) return launch_cl<128>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N <= 64) return launch_cl<64>(A, lda, Wt, ldw, ep, M, N, K, st);
  // 96-wide tiles only for f32 outputs: a bf16 store box spans 64 columns and must not straddle tiles
  if (N <= 96 && c_dtype == SPA3D_F32) return launch_cl<96>(A, lda, Wt, ldw, ep, M, N, K, st);
  return launch_cl<128>(A, lda, Wt, ldw, ep, M, N, K, st);
}

}  // namespace spa3d
