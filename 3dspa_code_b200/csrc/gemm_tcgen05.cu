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
//                               bias / tanh-GELU / residual in registers, 128-bit global stores.
// The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Out-of-range rows / columns / K are zero-filled by TMA and masked in
// the epilogue, so M, N need no padding (K % 8 == 0 for the 16-byte TMA stride rule).
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

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
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
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    EpiParams ep, int64_t M, int N, int K) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int TMEM_COLS = tmem_cols_for(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * L::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t m_tiles = (M + BM - 1) / BM;
  const int64_t num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
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
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          tma_load_2d(smem_a + stage * L::A_BYTES, &tmA, kb * BK, m_blk * BM, &full_bar[stage]);
          tma_load_2d(smem_b + stage * L::B_BYTES, &tmB, kb * BK, n_blk * BN, &full_bar[stage]);
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
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = (int)(tile / n_tiles), n_blk = (int)(tile % n_tiles);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      const int64_t row = (int64_t)m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c0;
        if (row_ok && col0 < N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // groups of 8 columns
            const int col = col0 + g * 8;
            if (col >= N) break;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            if (ep.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (ep.act == SPA3D_ACT_GELU_TANH) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = gelu_tanh(v[i]);
            }
            if (ep.residual) {
              if (ep.r_dtype == SPA3D_F32) {
                const float* rp = reinterpret_cast<const float*>(ep.residual) + row * ep.ldr + col;
                const float4 r0 = *reinterpret_cast<const float4*>(rp);
                const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
                v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              } else {
                const uint4 rr = *reinterpret_cast<const uint4*>(
                    reinterpret_cast<const bf16*>(ep.residual) + row * ep.ldr + col);
                const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
                  v[2 * i] += __low2float(h);
                  v[2 * i + 1] += __high2float(h);
                }
              }
            }
            if (ep.c_dtype == SPA3D_F32) {
              float* cp = reinterpret_cast<float*>(ep.C) + row * ep.ldc + col;
              *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.C) + row * ep.ldc + col) =
                  make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[as]);
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

template <int BN>
static int launch(const void* A, int64_t lda, const void* Wt, int64_t ldw, const EpiParams& ep,
                  int64_t M, int N, int K, cudaStream_t st) {
  using L = SmemLayout<BN>;
  CUtensorMap tmA, tmB;
  if (make_map(&tmA, A, M, K, lda, BM)) return 1;
  if (make_map(&tmB, Wt, N, K, ldw, BN)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    SPA3D_REQUIRE(e == cudaSuccess, "gemm_tcgen05: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  gemm_tcgen05_kernel<BN><<<grid, NUM_THREADS, L::TOTAL, st>>>(tmA, tmB, ep, M, N, K);
  return check_launch("gemm_tcgen05");
}

}  // namespace tc

bool gemm_tcgen05_applicable(const void* A, int64_t lda, const void* Wt, int64_t ldw, int64_t M,
                             int N, int K) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(A) && al16(Wt) && (lda % 8 == 0) && (ldw % 8 == 0) && (K % 8 == 0) && (N % 8 == 0) &&
         M > 0 && M < (1ll << 31);
}

int gemm_tcgen05(const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, int act,
                 const void* residual, int64_t ldr, int r_dtype, void* C, int64_t ldc, int c_dtype,
                 int64_t M, int N, int K, cudaStream_t st) {
  using namespace tc;
  SPA3D_REQUIRE(c_dtype == SPA3D_F32 || c_dtype == SPA3D_BF16, "gemm_tcgen05: bad C dtype");
  SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % (c_dtype == SPA3D_F32 ? 4 : 8) == 0,
                "gemm_tcgen05: C must be 16-byte aligned with 16-byte row pitch");
  if (residual)
    SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % (r_dtype == SPA3D_F32 ? 4 : 8) == 0,
                  "gemm_tcgen05: residual must be 16-byte aligned with 16-byte row pitch");
  if (bias) SPA3D_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_tcgen05: bias must be 16-byte aligned");
  EpiParams ep{bias, residual, ldr, r_dtype, C, ldc, c_dtype, act};
  // tile width: the widest of {256,192,128} that divides N, else the narrowest tile covering N
  if (N % 256 == 0) return launch<256>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N % 192 == 0) return launch<192>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N % 128 == 0) return launch<128>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N <= 64) return launch<64>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N <= 96) return launch<96>(A, lda, Wt, ldw, ep, M, N, K, st);
  if (N % 96 == 0 && N < 512) return launch<96>(A, lda, Wt, ldw, ep, M, N, K, st);
  return launch<128>(A, lda, Wt, ldw, ep, M, N, K, st);
}

}  // namespace spa3d
