// Fused track embedding (track_autoencoder_3d.py:123-149 + track_autoencoder.py:18-38):
//   x[r'] = [ sin-Fourier(x,y,z,t/T) | dino[r] | depth[r] ] . W_embed + (b_track + b_dino + b_depth)
// for every (track, frame) row r, written at row r' = r + r/T + 1 of the token matrix (slot 0 of every
// 151-row sequence is the learned read-out token, :161-165).
//
// The reference sums three Dense outputs; that equals one contraction of the concatenated features
// with the three kernels stacked along K.  Here the concatenation never exists in HBM:
// the A operand of the tcgen05 GEMM is produced on the fly.  8 producer warps read the fp32 DINO and
// depth features with 128-bit loads (16 lanes per row segment, 512 contiguous bytes per warp
// instruction), round to bf16 and write the 128B-swizzled K-major tile the tensor core reads; the
// four Fourier K-blocks (one per coordinate, 32 sin + 32 cos features) are evaluated in registers
// with sin.approx (bf16 path: |err| ~1e-4 at the largest argument, well under bf16 rounding).  The
// weight tile arrives by TMA.  One CTA owns all W output columns of its 128 rows (two 192-column
// accumulators in TMEM), so every input byte is read from HBM exactly once: 2048*150*(1024*4+12)
// = 1.26 GB per clip, the kernel is HBM-bound.
//
//   warp 0      TMA producer (weights)     warp 1      MMA issuer
//   warps 2..5  epilogue: TMEM -> +bias -> smem transpose -> coalesced fp32 stores (row remap)
//   warps 6..13 A-tile producers
#include <cuda_pipeline.h>

#include "tc_ptx.cuh"

namespace spa3d {
namespace te {

using namespace tc;

constexpr int NUM_PROD_WARPS = 8;
constexpr int NUM_EPI = 4;
constexpr int THREADS = 64 + 32 * NUM_EPI + 32 * NUM_PROD_WARPS;   // 448
constexpr int STAGES = 3;

struct EmbedParams {
  const float* tracks;   // [R, 3]
  const float* dino;     // [R, Dd] or null
  const float* depth;    // [R, Dz] or null
  const float* bias;     // [W] summed biases
  float* out;            // [R + R/T, W] token matrix (row remap r -> r + r/T + 1)
  bf16* acat;            // optional [R + R/T, K] bf16 copy of the concatenated features (training: operand of dW), same row remap
  int64_t lda;
  int64_t ldo;
  int64_t R;
  int T, Dd, Dz, W;
  float inv_scale;       // 1 / track_scale_factor
  float inv_T;           // 1 / T
  float fs[32];          // 2^(i/3)
  int l2_prefetch;       // pull whole feature rows into L2 one tile ahead (SPA3D_EMBED_L2_PREFETCH=1 enables; measured slower)
};

template <int NB, int BNH, bool ACAT>   // W = NB * BNH output columns, BNH <= 256; ACAT: also store the bf16 features
__global__ void __launch_bounds__(THREADS, 1)
embed_fused_kernel(const __grid_constant__ CUtensorMap tmB, const EmbedParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = NB * BNH * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_epi = smem + STAGES * STAGE_BYTES;            // [4 warps][4096]
  float* smem_bias = reinterpret_cast<float*>(smem_epi + NUM_EPI * 4096);   // [NB*BNH]
  float* smem_trk = smem_bias + NB * BNH;                                     // [2][128 rows][3] coordinates of a tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_trk + 2 * BM * 3);
  uint64_t* full_bar = bars;                 // [STAGES]  1 TMA arrive (expect_tx) + 8 producer warps
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // accumulator complete
  uint64_t* tempty_bar = bars + 2 * STAGES + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_dino0 = 4, kb_depth0 = 4 + p.Dd / 64;
  const int num_kb = kb_depth0 + p.Dz / 64;
  const int64_t m_tiles = (p.R + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1 + NUM_PROD_WARPS);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, NUM_EPI * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < NB * BNH; i += THREADS) smem_bias[i] = p.bias[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== weights by TMA =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], B_BYTES);
#pragma unroll
          for (int h = 0; h < NB; ++h)
            tma_load_2d(smem_b + stage * B_BYTES + h * (BNH * BK * 2), &tmB, kb * BK, h * BNH, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BNH);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        mbar_wait(tempty_bar, (it & 1) ^ 1);   // the epilogue has drained the previous tile
        tcgen05_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_BYTES));
#pragma unroll
          for (int h = 0; h < NB; ++h) {
            const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * B_BYTES + h * (BNH * BK * 2)));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base + (uint32_t)(h * BNH), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else if (warp < 2 + NUM_EPI) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    uint8_t* slab = smem_epi + (size_t)(warp - 2) * 4096;
    const int cr = lane >> 3, cc = lane & 7;
    int it = 0;
    for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int64_t row0 = t * BM + quarter * 32;
      // output rows of the 8 coalesced passes (row remap r -> r + r/T + 1), once per tile
      float* orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r_ = row0 + i * 4 + cr;
        orow[i] = r_ < p.R ? p.out + (r_ + (int64_t)((uint32_t)r_ / (uint32_t)p.T) + 1) * p.ldo + cc * 4 : nullptr;
      }
      mbar_wait(tfull_bar, it & 1);
      tcgen05_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
      constexpr int NCH = NB * BNH / 32;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c == NCH - 1) {
          tcgen05_fence_before();
          mbar_arrive(tempty_bar);
        }
        const int col0 = c * 32;
        uint8_t* srow = slab + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const float4 b = *reinterpret_cast<const float4*>(smem_bias + col0 + cc * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + cr;
          if (orow[i] != nullptr) {
            float4 a = *reinterpret_cast<const float4*>(slab + rr * 128 + ((cc ^ (rr & 7)) << 4));
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            *reinterpret_cast<float4*>(orow[i] + col0) = a;
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== A-tile producers (8 warps, 256 threads) =====================
    // Software pipelined over the flat (tile, K-block) sequence: the 128-bit loads of item i+1 are
    // in flight while item i is converted and written to shared memory (two register buffers), so
    // every producer thread keeps 2 x 8 x 16 B of HBM reads outstanding.
    const int ptid = threadIdx.x - (2 + NUM_EPI) * 32;   // 0..255
    const int rsub = ptid >> 4, l16 = ptid & 15;         // 16 lanes per row segment, 16 rows per pass
    const int64_t my_tiles = blockIdx.x < m_tiles ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // feature K-blocks: straight-line load / convert code with no data-dependent branch between the
    // issue of a load and its use (a branch there makes ptxas wait for every outstanding load)
    auto issue = [&](int64_t tile_iter, int kb, float4 (&v)[8]) {
      const int64_t row_base = (blockIdx.x + tile_iter * gridDim.x) * BM;
      const bool is_dino = kb < kb_depth0;
      const float* src = is_dino ? p.dino : p.depth;
      const int64_t ld = is_dino ? p.Dd : p.Dz;
      const int col = (is_dino ? kb - kb_dino0 : kb - kb_depth0) * 64 + l16 * 4;
#pragma unroll
      for (int ps = 0; ps < 8; ++ps) {
        const int64_t r_ = row_base + ps * 16 + rsub;
        const float* ptr = src + (r_ < p.R ? r_ : 0) * ld + col;      // clamped: always a valid address
        v[ps] = __ldcs(reinterpret_cast<const float4*>(ptr));
      }
    };
    auto acquire = [&](int64_t idx) -> uint8_t* {
      const int stage = (int)(idx % STAGES);
      const uint32_t phase = (uint32_t)((idx / STAGES) & 1);
      if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1);
      __syncwarp();
      return smem_a + stage * A_BYTES;
    };
    auto publish = [&](int64_t idx) {
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[(int)(idx % STAGES)]);
    };
    auto consume = [&](int64_t tile_iter, int kb, const float4 (&v)[8]) {
      const int64_t idx = tile_iter * num_kb + kb;
      const int64_t row_base = (blockIdx.x + tile_iter * gridDim.x) * BM;
      uint8_t* at = acquire(idx);
#pragma unroll
      for (int ps = 0; ps < 8; ++ps) {
        const int row = ps * 16 + rsub;
        const bool ok = row_base + row < p.R;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(ok ? v[ps].x : 0.f, ok ? v[ps].y : 0.f);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(ok ? v[ps].z : 0.f, ok ? v[ps].w : 0.f);
        const uint2 w2 = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(at + row * 128 + (((l16 >> 1) ^ (row & 7)) << 4) + (l16 & 1) * 8) = w2;
        if (ACAT && ok) {   // 16 lanes x 8 B = one 128-byte line of the row
          const uint32_t r32 = (uint32_t)(row_base + row);
          *reinterpret_cast<uint2*>(p.acat + ((int64_t)r32 + r32 / (uint32_t)p.T + 1) * p.lda + kb * 64 + l16 * 4) = w2;
        }
      }
      publish(idx);
    };
    // coordinates (x, y, z) of the tile's 128 rows: 1.5 KB copied global -> shared with cp.async one
    // tile ahead (no registers held, no latency in the Fourier blocks)
    auto load_tracks = [&](int64_t tile_iter) {
      if (ptid < BM * 3 / 4) {
        const int64_t f0 = (blockIdx.x + tile_iter * gridDim.x) * BM * 3 + ptid * 4;   // first of 4 floats
        float* dst = smem_trk + (tile_iter & 1) * (BM * 3) + ptid * 4;
        if (f0 + 4 <= p.R * 3) {
          __pipeline_memcpy_async(dst, p.tracks + f0, 16);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[e] = f0 + e < p.R * 3 ? p.tracks[f0 + e] : 0.f;
        }
      }
      __pipeline_commit();
    };
    auto fourier = [&](int64_t tile_iter, int kb) {
      // Fourier features of coordinate kb: features f = l16*4 .. +3 (f < 32: sin(v s_f), else sin(v s_f + pi/2))
      const int64_t idx = tile_iter * num_kb + kb;
      const int64_t row_base = (blockIdx.x + tile_iter * gridDim.x) * BM;
      uint8_t* at = acquire(idx);
      const float* trk = smem_trk + (tile_iter & 1) * (BM * 3);
      const int f0 = (l16 * 4) & 31;
      const float ph = l16 >= 8 ? 1.57079632679489661923f : 0.f;
      const float s0 = p.fs[f0], s1 = p.fs[f0 + 1], s2 = p.fs[f0 + 2], s3 = p.fs[f0 + 3];
#pragma unroll
      for (int ps = 0; ps < 8; ++ps) {
        const int row = ps * 16 + rsub;
        const int64_t r_ = row_base + row;
        float x = kb < 3 ? trk[row * 3 + kb] : (float)((uint32_t)r_ % (uint32_t)p.T) * p.inv_T;
        x *= p.inv_scale;
        float o0, o1, o2, o3;
        asm("sin.approx.f32 %0, %1;" : "=f"(o0) : "f"(__fadd_rn(__fmul_rn(x, s0), ph)));
        asm("sin.approx.f32 %0, %1;" : "=f"(o1) : "f"(__fadd_rn(__fmul_rn(x, s1), ph)));
        asm("sin.approx.f32 %0, %1;" : "=f"(o2) : "f"(__fadd_rn(__fmul_rn(x, s2), ph)));
        asm("sin.approx.f32 %0, %1;" : "=f"(o3) : "f"(__fadd_rn(__fmul_rn(x, s3), ph)));
        if (r_ >= p.R) o0 = o1 = o2 = o3 = 0.f;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
        const uint2 w2 = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(at + row * 128 + (((l16 >> 1) ^ (row & 7)) << 4) + (l16 & 1) * 8) = w2;
        if (ACAT && r_ < p.R) {
          const uint32_t r32 = (uint32_t)r_;
          *reinterpret_cast<uint2*>(p.acat + ((int64_t)r32 + r32 / (uint32_t)p.T + 1) * p.lda + kb * 64 + l16 * 4) = w2;
        }
      }
      publish(idx);
    };
    // Optional experiment (off): pull every row's whole feature vector (3 KB + 1 KB contiguous) into L2 one
    // tile ahead so the 256-byte-per-row slices become L2 hits.  Measured slower than plain demand loads.
    auto prefetch_tile = [&](int64_t tile_iter) {
      if (tile_iter >= my_tiles || !p.l2_prefetch) return;
      const int64_t r_ = (blockIdx.x + tile_iter * gridDim.x) * BM + (ptid >> 1);
      if (r_ >= p.R) return;
      const float* src = (ptid & 1) ? p.depth : p.dino;
      const int64_t w = (ptid & 1) ? p.Dz : p.Dd;
      if (src != nullptr && w > 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + r_ * w), "r"((uint32_t)(w * 4)) : "memory");
    };

    float4 va[8], vb[8];
    const bool has_feat = num_kb > 4;
    if (my_tiles > 0) {
      prefetch_tile(0);
      load_tracks(0);
      if (has_feat) issue(0, 4, va);
    }
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      prefetch_tile(ti + 1);
      __pipeline_wait_prior(0);                              // this tile's coordinates have landed ...
      asm volatile("bar.sync 2, 256;" ::: "memory");        // ... for every producer warp
      if (ti + 1 < my_tiles) load_tracks(ti + 1);           // other half of the double buffer
      for (int kb = 0; kb < 4; ++kb) fourier(ti, kb);
      for (int kb = 4; kb < num_kb; kb += 2) {
        if (kb + 1 < num_kb) issue(ti, kb + 1, vb);
        consume(ti, kb, va);
        if (kb + 1 < num_kb) {
          if (kb + 2 < num_kb) issue(ti, kb + 2, va);
          else if (ti + 1 < my_tiles) issue(ti + 1, 4, va);   // first feature block of the next tile, over its Fourier phase
          consume(ti, kb + 1, vb);
        } else if (ti + 1 < my_tiles) {
          issue(ti + 1, 4, va);
        }
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

template <int NB, int BNH, bool ACAT>
static int launch_v(const CUtensorMap& tmB, const EmbedParams& p, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + NB * BNH * BK * 2) + NUM_EPI * 4096 + NB * BNH * 4 + 2 * BM * 3 * 4 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(embed_fused_kernel<NB, BNH, ACAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "embed_fused: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t m_tiles = (p.R + BM - 1) / BM;
  const int grid = (int)(m_tiles < num_sms() ? m_tiles : num_sms());
  embed_fused_kernel<NB, BNH, ACAT><<<grid, THREADS, SMEM, st>>>(tmB, p);
  return check_launch("embed_fused");
}

template <int NB, int BNH>
static int launch(const CUtensorMap& tmB, const EmbedParams& p, cudaStream_t st) {
  return p.acat != nullptr ? launch_v<NB, BNH, true>(tmB, p, st) : launch_v<NB, BNH, false>(tmB, p, st);
}

}  // namespace te
}  // namespace spa3d

using namespace spa3d;

extern "C" {

int spa3d_embed_fused_applicable(int W, int K_total, int dino_dim, int depth_dim, int coords) {
  if (coords != 3 || K_total != 256 + dino_dim + depth_dim) return 0;
  if (dino_dim % 64 || depth_dim % 64) return 0;
  return (W == 384 || W == 256 || W == 192 || W == 128) ? 1 : 0;   // W <= 384: 3 stages of A + W weight rows fit in smem
}

int spa3d_embed_fused(const float* tracks, const float* dino, const float* depth, const void* Wt, int64_t ldw,
                      const float* bias, float* out, int64_t ldo, void* a_cat, int64_t lda, int64_t rows, int T,
                      int dino_dim, int depth_dim, int W, int num_freq, float track_scale_factor, void* stream) {
  using namespace spa3d::te;
  SPA3D_REQUIRE(num_freq == 32, "embed_fused: 32 frequencies per coordinate (one 64-column K block each)");
  SPA3D_REQUIRE(spa3d_embed_fused_applicable(W, 256 + dino_dim + depth_dim, dino_dim, depth_dim, 3), "embed_fused: unsupported widths");
  SPA3D_REQUIRE((dino_dim == 0) == (dino == nullptr) && (depth_dim == 0) == (depth == nullptr), "embed_fused: feature pointer / width mismatch");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  SPA3D_REQUIRE(al16(dino) && al16(depth) && al16(Wt) && al16(out) && al16(bias) && ldw % 8 == 0 && ldo % 4 == 0,
                "embed_fused: operands must be 16-byte aligned");
  if (rows == 0) return 0;
  SPA3D_REQUIRE(rows < (1ll << 31), "embed_fused: too many rows");
  EmbedParams p;
  p.tracks = tracks; p.dino = dino; p.depth = depth; p.bias = bias; p.out = out; p.ldo = ldo; p.R = rows;
  p.acat = reinterpret_cast<bf16*>(a_cat); p.lda = lda;
  SPA3D_REQUIRE(a_cat == nullptr || (lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a_cat) & 7) == 0), "embed_fused: a_cat must be 8-byte aligned");
  p.T = T; p.Dd = dino_dim; p.Dz = depth_dim; p.W = W; p.inv_scale = (float)(1.0 / (double)track_scale_factor); p.inv_T = (float)(1.0 / (double)T);
  for (int i = 0; i < 32; ++i) p.fs[i] = (float)pow(2.0, (double)i / 3.0);
  {
    const char* e = getenv("SPA3D_EMBED_L2_PREFETCH");
    p.l2_prefetch = (e && atoi(e) == 1) ? 1 : 0;   // measured: 0.54 ms with, 0.48 ms without - off by default
  }
  const int K = 256 + dino_dim + depth_dim;
  CUtensorMap tmB;
  const int bnh = W > 256 ? W / 2 : W;
  if (tc::make_map(&tmB, Wt, W, K, ldw, bnh)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (W) {
    case 384: return launch<2, 192>(tmB, p, st);
    case 256: return launch<1, 256>(tmB, p, st);
    case 192: return launch<1, 192>(tmB, p, st);
    default: return launch<1, 128>(tmB, p, st);
  }
}

}  // extern "C"
