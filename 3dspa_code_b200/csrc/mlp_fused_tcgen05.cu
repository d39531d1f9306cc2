// Fused MLP sub-block of the per-track transformer, inference form (attention.py:102-108):
//   y = a + gelu_tanh(LN(a) . W1 + b1) . W2 + b2        with LN(a) given (bf16), width D = 384, hidden width Hd (1536)
// as ONE kernel: the Hd-wide hidden activation never leaves the SM.  Unfused, MLP_in writes and MLP_out re-reads
// 2 x M x Hd x 2 B = 1.9 GB per layer at M = 309,248 tokens - the largest single item of the 10 GB a layer moves.
//
// One persistent CTA per SM owns a 128-token tile:
//   * A = LN(a) tile [128 x 384] bf16 stays RESIDENT in shared memory (6 K-blocks, 96 KB, TMA, 128B swizzle);
//   * the hidden dimension is walked in chunks of 128: GEMM1 chunk  S_j = A . W1[j]^T  (6 K-blocks x 4 UMMAs, N = 128) into a
//     128-column TMEM accumulator; 16 epilogue warps (one per TMEM lane quarter and 32-column part - the thread-level
//     parallelism the GELU epilogue needs, see gemm_tcgen05.cu) add b1, apply tanh-GELU and write the bf16 chunk H_j as a
//     K-major 128B-swizzled operand tile (32 KB); GEMM2 chunk  Y += H_j . W2[:, j]^T  (2 K-blocks x 2 halves of N = 192) accumulates
//     into a 384-column TMEM accumulator (128 + 384 = all 512 columns);
//   * the MMA warp issues GEMM1(j+1) BEFORE GEMM2(j), so the tensor pipe works on the next chunk while the epilogue warps turn
//     chunk j into H_j; W1 / W2 tiles stream from L2 (2.4 MB of weights, resident there) through a 4-stage 24 KB ring;
//   * final epilogue: Y + b2 + residual (fp32) -> fp32 out, transposed through shared memory so that global accesses are
//     64-byte-per-row coalesced.
//
// STATUS (round 2): correct (tests/test_gpu_kernels.py::test_mlp_fused_matches_unfused_and_fp64) and exported, but NOT used by the
// model by default: on the per-track shape (M = 309,248) it takes 0.805 ms against 0.783 ms for the two separate GEMMs (0.344 +
// 0.417 when timed alone).  The bound is the weight stream, not HBM: a chunk consumes 192 KB of W1 / W2 tiles per 3,072 cycles of
// MMA work, and with the 96 KB token tile resident the ring can hold only 96 KB - about one L2 + TMA latency of look-ahead - so the
// tensor pipe waits for tiles about half of the time.  The fix is a B operand shared by a CTA pair (cta_group::2 halves the weight
// bytes each SM must stage), not a different schedule inside one CTA.
#include "tc_ptx.cuh"

namespace spa3d {
namespace tm {

using namespace tc;

constexpr int EW = 16;                       // epilogue warps
constexpr int THREADS = 64 + 32 * EW;        // 576
constexpr int CH = 128;                      // hidden chunk
constexpr int WSTAGES = 4;
constexpr int WSTAGE_BYTES = 192 * BK * 2;   // 24 KB: one W2 half tile (192 x 64) or one W1 tile (128 x 64, 16 KB)

struct MlpParams {
  const float* b1;        // [Hd]
  const float* b2;        // [D]
  const float* residual;  // [M, D] fp32 (ldr)
  float* out;             // [M, D] fp32 (ldo)
  int64_t ldr, ldo, M;
  int Hd;
};

__device__ __forceinline__ void bias32(uint64_t (&v)[16], float breg) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    v[i] = add2(v[i], pk(__shfl_sync(0xffffffffu, breg, 2 * i), __shfl_sync(0xffffffffu, breg, 2 * i + 1)));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

template <int D>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  constexpr int KB1 = D / BK;                // K-blocks of GEMM1 (6)
  constexpr int KB2 = CH / BK;               // K-blocks of one GEMM2 chunk (2)
  constexpr int NH = D / 2;                  // GEMM2 output half (192 columns per UMMA)
  constexpr int A_BYTES = KB1 * BM * BK * 2; // 96 KB
  constexpr int H_BYTES = KB2 * BM * BK * 2; // 32 KB
  static_assert(D == 384, "accumulator budget: 128 (hidden chunk) + D columns of TMEM");
  constexpr uint32_t C_ACC1 = 0, C_ACC2 = CH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sH = sA + A_BYTES;
  uint8_t* sW = sH + H_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WSTAGES * WSTAGE_BYTES);
  uint64_t* a_full = bars, *a_empty = bars + 1, *acc1_full = bars + 2, *acc1_empty = bars + 3, *h_full = bars + 4, *h_empty = bars + 5;
  uint64_t* acc2_full = bars + 6, *acc2_empty = bars + 7, *w_full = bars + 8, *w_empty = bars + 8 + WSTAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8 + 2 * WSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NJ = p.Hd / CH;
  const int64_t m_tiles = (p.M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
    mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(acc1_full, 1); mbar_init(acc1_empty, EW * 32);
    mbar_init(h_full, EW * 32); mbar_init(h_empty, 1); mbar_init(acc2_full, 1); mbar_init(acc2_empty, EW * 32);
    for (int i = 0; i < WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer: A tile, then the weight tiles in the order the MMA warp consumes them =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t wph = 0;
      auto load_w1 = [&](int j) {
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&w_empty[stage], wph ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], (uint32_t)(CH * BK * 2));
          tma_load_2d(sW + stage * WSTAGE_BYTES, &tmW1, kb * BK, j * CH, &w_full[stage]);
          if (++stage == WSTAGES) { stage = 0; wph ^= 1; }
        }
      };
      auto load_w2 = [&](int j) {
        for (int kb2 = 0; kb2 < KB2; ++kb2)
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&w_empty[stage], wph ^ 1);
            mbar_arrive_expect_tx(&w_full[stage], (uint32_t)(NH * BK * 2));
            tma_load_2d(sW + stage * WSTAGE_BYTES, &tmW2, j * CH + kb2 * BK, hf * NH, &w_full[stage]);
            if (++stage == WSTAGES) { stage = 0; wph ^= 1; }
          }
      };
      uint32_t tph = 0;
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, tph ^= 1) {
        mbar_wait(a_empty, tph ^ 1);
        mbar_arrive_expect_tx(a_full, (uint32_t)A_BYTES);
        for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sA + kb * (BM * BK * 2), &tmA, kb * BK, (int)(t * BM), a_full);
        load_w1(0);
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ) load_w1(j + 1);
          load_w2(j);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t id1 = make_idesc(CH), id2 = make_idesc(NH);
      int stage = 0;
      uint32_t wph = 0, tph = 0;
      uint32_t c1 = 0, c2 = 0;   // chunks issued so far: GEMM1 / GEMM2 (phases of the accumulator and H barriers)
      auto gemm1 = [&]() {
        mbar_wait(acc1_empty, (c1 & 1) ^ 1);     // the epilogue warps have read the previous chunk out of the accumulator
        tcgen05_fence_after();
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&w_full[stage], wph);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(sA + kb * (BM * BK * 2)));
          const uint64_t db = make_smem_desc(smem_u32(sW + stage * WSTAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base + C_ACC1, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), id1, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&w_empty[stage]);
          if (++stage == WSTAGES) { stage = 0; wph ^= 1; }
        }
        umma_commit(acc1_full);
        ++c1;
      };
      auto gemm2 = [&](int j) {
        mbar_wait(h_full, c2 & 1);               // H_j is in shared memory
        tcgen05_fence_after();
        for (int kb2 = 0; kb2 < KB2; ++kb2)
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&w_full[stage], wph);
            tcgen05_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(sH + kb2 * (BM * BK * 2)));
            const uint64_t db = make_smem_desc(smem_u32(sW + stage * WSTAGE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base + C_ACC2 + (uint32_t)(hf * NH), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), id2, (j > 0 || kb2 > 0 || k > 0) ? 1u : 0u);
            umma_commit(&w_empty[stage]);
            if (++stage == WSTAGES) { stage = 0; wph ^= 1; }
          }
        umma_commit(h_empty);
        ++c2;
      };
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, tph ^= 1) {
        mbar_wait(a_full, tph);
        tcgen05_fence_after();
        gemm1();
        for (int j = 0; j < NJ; ++j) {
          if (j + 1 < NJ) {
            gemm1();
            if (j + 2 == NJ) umma_commit(a_empty);   // the last GEMM1 of this tile has been issued: A may be reloaded once it retires
          }
          if (j == 0) {
            if (NJ == 1) umma_commit(a_empty);
            mbar_wait(acc2_empty, tph ^ 1);          // the final epilogue of the previous tile has drained Y
            tcgen05_fence_after();
          }
          gemm2(j);
        }
        umma_commit(acc2_full);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter
    const int part = ew >> 2;              // 32-column part of the 128-wide hidden chunk / 96-column part of Y
    const int rloc = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t c = 0, tph = 0;               // global chunk counter / tile phase
    // H_j operand tile: K-block = part / 2, 16-byte chunk index ((part & 1) * 4 + i) ^ (row & 7) within the 128-byte row
    uint8_t* hrow = sH + (part >> 1) * (BM * BK * 2) + rloc * 128;
    const int hx = rloc & 7, hc0 = (part & 1) * 4;
    uint8_t* slab = sH + ew * 2048;        // final epilogue: one [32 rows x 16 fp32] slab per warp
    for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, tph ^= 1) {
      for (int j = 0; j < NJ; ++j, ++c) {
        const float breg = __ldg(p.b1 + j * CH + part * 32 + lane);
        mbar_wait(acc1_full, c & 1);
        tcgen05_fence_after();
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_off + C_ACC1 + (uint32_t)(part * 32), r);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(acc1_empty);
        uint64_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pku(r[2 * i], r[2 * i + 1]);
        bias32(v, breg);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(gelu2_fast(v[i]));
        mbar_wait(h_empty, (c & 1) ^ 1);   // GEMM2 of the previous chunk has read H
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts128(smem_u32(hrow) + (uint32_t)(((hc0 + i) ^ hx) << 4), w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        fence_proxy_async_smem();
        mbar_arrive(h_full);
      }
      // ---- final epilogue: Y + b2 + residual -> out, 96 columns per warp in six 16-column pieces through a 2 KB slab ----
      mbar_wait(acc2_full, tph);
      tcgen05_fence_after();
      const int64_t row0 = t * BM + quarter * 32;
      const int cr = lane >> 2, cc = lane & 3;           // coalesced view: 8 rows x 4 lanes x 16 B per pass
#pragma unroll 1
      for (int pc = 0; pc < 6; ++pc) {
        const int col0 = part * 96 + pc * 16;
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_off + C_ACC2 + (uint32_t)col0, r);
        tmem_ld_wait();
        if (pc == 5) {
          tcgen05_fence_before();
          mbar_arrive(acc2_empty);
        }
        const uint32_t srow = smem_u32(slab) + lane * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(srow + (uint32_t)((i ^ (lane & 3)) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        __syncwarp();
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + cc * 4));
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int rr = ps * 8 + cr;
          const int64_t grow = row0 + rr;
          float4 a = lds128(smem_u32(slab) + rr * 64 + ((cc ^ (rr & 3)) << 4));
          if (grow < p.M) {
            const float4 rs = *reinterpret_cast<const float4*>(p.residual + grow * p.ldr + col0 + cc * 4);
            a.x += b4.x + rs.x; a.y += b4.y + rs.y; a.z += b4.z + rs.z; a.w += b4.w + rs.w;
            *reinterpret_cast<float4*>(p.out + grow * p.ldo + col0 + cc * 4) = a;
          }
        }
        __syncwarp();
      }
      // the slabs live in the H operand region: nobody may write the next tile's first H chunk before every warp is done with them
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tm
}  // namespace spa3d

using namespace spa3d;

extern "C" {

int spa3d_mlp_fused_applicable(int D, int Hd) { return (D == 384 && Hd >= 128 && Hd % 128 == 0) ? 1 : 0; }

int spa3d_mlp_fused(const void* A, int64_t lda, const void* W1t, int64_t ldw1, const float* b1, const void* W2t, int64_t ldw2,
                    const float* b2, const float* residual, int64_t ldr, float* out, int64_t ldo, int64_t M, int D, int Hd,
                    void* stream) {
  using namespace spa3d::tm;
  SPA3D_REQUIRE(spa3d_mlp_fused_applicable(D, Hd), "mlp_fused: width 384 and a hidden width that is a multiple of 128 only (D=%d, Hd=%d)", D, Hd);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  SPA3D_REQUIRE(al16(A) && al16(W1t) && al16(W2t) && al16(b1) && al16(b2) && al16(residual) && al16(out) && lda % 8 == 0 && ldw1 % 8 == 0 &&
                    ldw2 % 8 == 0 && ldr % 4 == 0 && ldo % 4 == 0,
                "mlp_fused: operands must be 16-byte aligned with 16-byte row pitches");
  SPA3D_REQUIRE(b1 && b2 && residual, "mlp_fused: biases and residual are required");
  if (M == 0) return 0;
  SPA3D_REQUIRE(M < (1ll << 31), "mlp_fused: too many rows");
  CUtensorMap tmA, tmW1, tmW2;
  if (tc::make_map(&tmA, A, M, D, lda, tc::BM)) return 1;
  if (tc::make_map(&tmW1, W1t, Hd, D, ldw1, CH)) return 1;
  if (tc::make_map(&tmW2, W2t, D, Hd, ldw2, D / 2)) return 1;
  constexpr int SMEM = 6 * 128 * 64 * 2 + 2 * 128 * 64 * 2 + WSTAGES * WSTAGE_BYTES + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "mlp_fused: smem attribute (%d B): %s", SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  MlpParams p{b1, b2, residual, out, ldr, ldo, M, Hd};
  const int64_t m_tiles = (M + tc::BM - 1) / tc::BM;
  const int grid = (int)(m_tiles < num_sms() ? m_tiles : num_sms());
  stat_add(ST_GEMM_TCGEN05);
  mlp_fused_kernel<384><<<grid, THREADS, SMEM, (cudaStream_t)stream>>>(tmA, tmW1, tmW2, p);
  return check_launch("mlp_fused");
}

}  // extern "C"
