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
  // SAMPLE variant ("project, then sample": the DINO projection is linear, so the patch map is projected ONCE per clip
  // and every track row blends four projected patch rows instead of projecting its own blended 768-vector)
  const bf16* proj;      // [T*Hp*Wp, W] projected DINO patch map
  const float* trk2d;    // [R, 2] pixel coordinates of the row's track point
  const float* dfeat;    // [R, 4] (d, d/10, d_t - d_{t-1}, 0): the three live depth-feature channels (inference.py:437-443), or null
  const float* wdep;     // [3, W] rows 0..2 of the depth projection kernel (fp32), or null
  int Hp, Wp;
  float scale_w, scale_h;   // Wp / video_W, Hp / video_H (inference.py:367-368)
  int debug;                // -DSPA3D_EMBED_DEBUG=n builds only (timing experiments): 1 = all gathers hit one patch row, 2 = no stores; 0 in the product
};

template <int NB, int BNH, bool ACAT, bool SAMPLE>   // W = NB * BNH output columns, BNH <= 256; ACAT: also store the bf16 features
__global__ void __launch_bounds__(THREADS, 1)
embed_fused_kernel(const __grid_constant__ CUtensorMap tmB, const EmbedParams p) {
  // SAMPLE re-balances the warps: its A operand is the four Fourier K-blocks only (no feature streaming), while its epilogue
  // gathers and blends projected patch rows, so it runs 8 epilogue warps (two per TMEM lane quarter, alternating 32-column
  // chunks) and 4 producer warps on a 2-stage ring; the streaming variant keeps 4 + 8 on 3 stages.
  constexpr int NEPI = SAMPLE ? 8 : NUM_EPI;               // epilogue warps 2 .. 2+NEPI-1
  constexpr int NPW = SAMPLE ? 4 : NUM_PROD_WARPS;         // producer warps
  constexpr int PT = NPW * 32;                             // producer threads
  constexpr int RPP = PT / 16;                             // rows per producer pass (16 lanes per row segment)
  constexpr int PASSES = BM / RPP;
  constexpr int STAGES = SAMPLE ? 2 : te::STAGES;
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = NB * BNH * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_epi = smem + STAGES * STAGE_BYTES;            // [4 warps][4096]
  float* smem_bias = reinterpret_cast<float*>(smem_epi + NEPI * 4096);   // [NB*BNH]
  float* smem_trk = smem_bias + NB * BNH;                                     // [2][128 rows][3] coordinates of a tile
  float* smem_wdep = smem_trk + 2 * BM * 3;                                  // SAMPLE: [3][W] depth-projection rows
  uint32_t* smem_smp = reinterpret_cast<uint32_t*>(smem_wdep + (SAMPLE ? 3 * NB * BNH : 0));   // SAMPLE: [8 warps][32 rows][12 words]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_smp + (SAMPLE ? NEPI * 32 * 12 : 0));
  uint64_t* full_bar = bars;                 // [STAGES]  1 TMA arrive (expect_tx) + 8 producer warps
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // accumulator complete
  uint64_t* tempty_bar = bars + 2 * STAGES + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_dino0 = 4, kb_depth0 = 4 + p.Dd / 64;
  const int num_kb = kb_depth0 + p.Dz / 64;
  // SAMPLE: tiles are frame-major - 128 tracks of ONE frame - so every gather of a tile (and of the ~148 tiles in flight)
  // hits the same 1 MB slice of the projected map, which stays in L2 and is read from HBM once.  Row i of tile tt is
  // track n = (tt % nt_n) * 128 + i at frame t = tt / nt_n; its data lives at flat row n*T + t like in the other variant.
  const int Ntrk = (int)(p.R / p.T);
  const int nt_n = (Ntrk + BM - 1) / BM;
  const int64_t m_tiles = SAMPLE ? (int64_t)p.T * nt_n : (p.R + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1 + NPW);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, NEPI * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < NB * BNH; i += THREADS) smem_bias[i] = p.bias[i];
  if constexpr (SAMPLE)
    for (int i = threadIdx.x; i < 3 * NB * BNH; i += THREADS) smem_wdep[i] = p.wdep != nullptr ? p.wdep[i] : 0.f;
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
  } else if (warp < 2 + NEPI) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int c_first = (SAMPLE && warp >= 6) ? 1 : 0;   // SAMPLE: warps 2..5 take the even chunks, 6..9 the odd ones
    constexpr int c_step = SAMPLE ? 2 : 1;
    uint8_t* slab = smem_epi + (size_t)(warp - 2) * 4096;
    const int cr = lane >> 3, cc = lane & 7;
    int it = 0;
    for (int64_t t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int64_t row0 = t * BM + quarter * 32;
      const int tf = SAMPLE ? (int)(t / nt_n) : 0;                               // SAMPLE: the tile's frame ...
      const int n0 = SAMPLE ? (int)(t % nt_n) * BM + quarter * 32 : 0;           // ... and this warp's first track
      // output rows of the 8 coalesced passes (row remap r -> r + r/T + 1), once per tile
      float* orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if constexpr (SAMPLE) {
          const int n = n0 + i * 4 + cr;
          orow[i] = n < Ntrk ? p.out + ((int64_t)n * (p.T + 1) + tf + 1) * p.ldo + cc * 4 : nullptr;
        } else {
          const int64_t r_ = row0 + i * 4 + cr;
          orow[i] = r_ < p.R ? p.out + (r_ + (int64_t)((uint32_t)r_ / (uint32_t)p.T) + 1) * p.ldo + cc * 4 : nullptr;
        }
      }
      uint32_t* smp = smem_smp + (warp - 2) * (32 * 12);
      if constexpr (SAMPLE) {
        // lane = row: bilinear set-up of this row's track point on the patch grid, once per tile
        const int n = n0 + lane;
        const int64_t r_ = (int64_t)n * p.T + tf;
        uint32_t w[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (n < Ntrk) {
          const float2 xy = *reinterpret_cast<const float2*>(p.trk2d + r_ * 2);
          const int t = tf;
          const Bilin b = bilin_setup(__fmul_rn(xy.x, p.scale_w), __fmul_rn(xy.y, p.scale_h), p.Wp, p.Hp);
          const float omx = __fsub_rn(1.f, b.wx), omy = __fsub_rn(1.f, b.wy);
          const int fb = t * p.Hp * p.Wp;
          if (p.debug & 1) { w[0] = w[1] = w[2] = w[3] = 0; } else {
          w[0] = (uint32_t)((fb + b.y0 * p.Wp + b.x0) * p.W);
          w[1] = (uint32_t)((fb + b.y0 * p.Wp + b.x1) * p.W);
          w[2] = (uint32_t)((fb + b.y1 * p.Wp + b.x0) * p.W);
          w[3] = (uint32_t)((fb + b.y1 * p.Wp + b.x1) * p.W); }
          w[4] = __float_as_uint(omx * omy);
          w[5] = __float_as_uint(b.wx * omy);
          w[6] = __float_as_uint(omx * b.wy);
          w[7] = __float_as_uint(b.wx * b.wy);
          if (p.dfeat != nullptr) {
            const float4 df = *reinterpret_cast<const float4*>(p.dfeat + r_ * 4);
            w[8] = __float_as_uint(df.x);
            w[9] = __float_as_uint(df.y);
            w[10] = __float_as_uint(df.z);
          }
        }
        __syncwarp();   // the previous tile's passes have read their parameters
#pragma unroll
        for (int j = 0; j < 3; ++j) *reinterpret_cast<uint4*>(smp + lane * 12 + j * 4) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        __syncwarp();
      }
      mbar_wait(tfull_bar, it & 1);
      tcgen05_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
      constexpr int NCH = NB * BNH / 32;
      // SAMPLE: the gathers of the four projected patch rows are issued half a chunk (4 passes x 4 neighbours = 16 loads of
      // 8 B per lane) ahead of their use and unconditionally (rows past the end carry offset 0), so their L2 latency runs
      // under the TMEM read, the transpose and the other half's arithmetic instead of once per pass.
      auto gather = [&](int c, int half, uint2 (&g)[4][4]) {
        const bf16* pj = p.proj + c * 32 + cc * 4;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 offf = lds128(smem_u32(smp) + (((half * 4 + i4) * 4 + cr) * 12) * 4);
          const uint4 off = make_uint4(__float_as_uint(offf.x), __float_as_uint(offf.y), __float_as_uint(offf.z), __float_as_uint(offf.w));
          g[i4][0] = __ldg(reinterpret_cast<const uint2*>(pj + off.x));
          g[i4][1] = __ldg(reinterpret_cast<const uint2*>(pj + off.y));
          g[i4][2] = __ldg(reinterpret_cast<const uint2*>(pj + off.z));
          g[i4][3] = __ldg(reinterpret_cast<const uint2*>(pj + off.w));
        }
      };
      auto finish = [&](int c, int half, const uint2 (&g)[4][4], const float4& b) {
        const int col0 = c * 32;
        const float4 w0 = *reinterpret_cast<const float4*>(smem_wdep + col0 + cc * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(smem_wdep + NB * BNH + col0 + cc * 4);
        const float4 w2 = *reinterpret_cast<const float4*>(smem_wdep + 2 * NB * BNH + col0 + cc * 4);
        auto lo = [](uint32_t u) { return __uint_as_float(u << 16); };
        auto hi = [](uint32_t u) { return __uint_as_float(u & 0xffff0000u); };
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int i = half * 4 + i4, rr = i * 4 + cr;
          const float4 wq = lds128(smem_u32(smp) + (rr * 12 + 4) * 4);
          const float4 dq = lds128(smem_u32(smp) + (rr * 12 + 8) * 4);
          float4 a = lds128(smem_u32(slab) + rr * 128 + ((cc ^ (rr & 7)) << 4));
          a.x += b.x + lo(g[i4][0].x) * wq.x + lo(g[i4][1].x) * wq.y + lo(g[i4][2].x) * wq.z + lo(g[i4][3].x) * wq.w;
          a.y += b.y + hi(g[i4][0].x) * wq.x + hi(g[i4][1].x) * wq.y + hi(g[i4][2].x) * wq.z + hi(g[i4][3].x) * wq.w;
          a.z += b.z + lo(g[i4][0].y) * wq.x + lo(g[i4][1].y) * wq.y + lo(g[i4][2].y) * wq.z + lo(g[i4][3].y) * wq.w;
          a.w += b.w + hi(g[i4][0].y) * wq.x + hi(g[i4][1].y) * wq.y + hi(g[i4][2].y) * wq.z + hi(g[i4][3].y) * wq.w;
          a.x += dq.x * w0.x + dq.y * w1.x + dq.z * w2.x;
          a.y += dq.x * w0.y + dq.y * w1.y + dq.z * w2.y;
          a.z += dq.x * w0.z + dq.y * w1.z + dq.z * w2.z;
          a.w += dq.x * w0.w + dq.y * w1.w + dq.z * w2.w;
          if (orow[i] != nullptr && !((p.debug & 2) && a.x != 12345.678f)) *reinterpret_cast<float4*>(orow[i] + col0) = a;
        }
      };
      static_assert(!SAMPLE || NCH % 2 == 0, "SAMPLE splits the chunks between two warps per quarter");
#pragma unroll 1
      for (int c = c_first; c < NCH; c += c_step) {
        uint2 ga[4][4], gb[4][4];
        if constexpr (SAMPLE) gather(c, 0, ga);
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c + c_step >= NCH) {   // this warp's last read of the accumulator
          tcgen05_fence_before();
          mbar_arrive(tempty_bar);
        }
        const int col0 = c * 32;
        const uint32_t srow = smem_u32(slab) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(srow + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const float4 b = lds128(smem_u32(smem_bias) + (col0 + cc * 4) * 4);
        if constexpr (SAMPLE) {
          gather(c, 1, gb);
          finish(c, 0, ga, b);
          finish(c, 1, gb, b);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + cr;
            if (orow[i] != nullptr) {
              float4 a = lds128(smem_u32(slab) + rr * 128 + ((cc ^ (rr & 7)) << 4));
              a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
              *reinterpret_cast<float4*>(orow[i] + col0) = a;
            }
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
    const int ptid = threadIdx.x - (2 + NEPI) * 32;      // 0..PT-1
    const int rsub = ptid >> 4, l16 = ptid & 15;         // 16 lanes per row segment, RPP rows per pass
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
      for (int ps = 0; ps < 8; ++ps) {   // streaming variant: PASSES == 8
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
      if constexpr (SAMPLE) {
        // frame-major tile: the 128 rows are 128 tracks at one frame, T*12 bytes apart
        const int64_t tile = blockIdx.x + tile_iter * gridDim.x;
        const int tf = (int)(tile / nt_n), n = (int)(tile % nt_n) * BM + ptid;
        if (ptid < BM) {   // one thread per row (PT == BM): three 4-byte copies
          float* dst = smem_trk + (tile_iter & 1) * (BM * 3) + ptid * 3;
          const float* src = p.tracks + ((int64_t)(n < Ntrk ? n : 0) * p.T + tf) * 3;
          __pipeline_memcpy_async(dst, src, 4);
          __pipeline_memcpy_async(dst + 1, src + 1, 4);
          __pipeline_memcpy_async(dst + 2, src + 2, 4);
        }
        __pipeline_commit();
        return;
      }
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
      for (int ps = 0; ps < PASSES; ++ps) {
        const int row = ps * RPP + rsub;
        int64_t r_ = row_base + row;
        float x;
        if constexpr (SAMPLE) {
          const int64_t tile = blockIdx.x + tile_iter * gridDim.x;
          x = kb < 3 ? trk[row * 3 + kb] : (float)(int)(tile / nt_n) * p.inv_T;
          r_ = ((int)(tile % nt_n) * BM + row) < Ntrk ? 0 : p.R;   // only used for the validity test below
        } else {
          x = kb < 3 ? trk[row * 3 + kb] : (float)((uint32_t)r_ % (uint32_t)p.T) * p.inv_T;
        }
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
      asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");   // ... for every producer warp
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

template <int NB, int BNH, bool ACAT, bool SAMPLE = false>
static int launch_v(const CUtensorMap& tmB, const EmbedParams& p, cudaStream_t st) {
  constexpr int NEPI = SAMPLE ? 8 : NUM_EPI;
  constexpr int SMEM = (SAMPLE ? 2 : STAGES) * (BM * BK * 2 + NB * BNH * BK * 2) + NEPI * 4096 + NB * BNH * 4 + 2 * BM * 3 * 4 + 256 + 1024 +
                       (SAMPLE ? 3 * NB * BNH * 4 + NEPI * 32 * 12 * 4 : 0);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(embed_fused_kernel<NB, BNH, ACAT, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SPA3D_REQUIRE(e == cudaSuccess, "embed_fused: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t m_tiles = SAMPLE ? (int64_t)p.T * ((p.R / p.T + BM - 1) / BM) : (p.R + BM - 1) / BM;
  const int grid = (int)(m_tiles < num_sms() ? m_tiles : num_sms());
  embed_fused_kernel<NB, BNH, ACAT, SAMPLE><<<grid, THREADS, SMEM, st>>>(tmB, p);
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
  p.proj = nullptr; p.trk2d = nullptr; p.dfeat = nullptr; p.wdep = nullptr; p.Hp = p.Wp = 0; p.scale_w = p.scale_h = 0.f; p.debug = 0;
  SPA3D_REQUIRE(a_cat == nullptr || (lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a_cat) & 7) == 0), "embed_fused: a_cat must be 8-byte aligned");
  p.T = T; p.Dd = dino_dim; p.Dz = depth_dim; p.W = W; p.inv_scale = (float)(1.0 / (double)track_scale_factor); p.inv_T = (float)(1.0 / (double)T);
  for (int i = 0; i < 32; ++i) p.fs[i] = (float)pow(2.0, (double)i / 3.0);
  p.l2_prefetch = 0;   // measured: 0.54 ms with the L2 prefetch of the next tile's features, 0.48 ms without
  spa3d::stat_add(spa3d::ST_EMBED_FUSED);
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

/* K1, "project then sample" form (inference.py:543-590): tokens for track points whose DINO / depth features
 * are never materialised.  out[r'] = Fourier(xyz[r], t/T) . Wt[:, :256]^T + bias
 *                                   + sum_k w_k(r) * proj[patch_k(r)]                  (bilinear blend of PROJECTED patch rows)
 *                                   + dfeat[r, 0..2] . wdep[0..2]                      (the three live depth channels)
 * proj = bf16(dino_map) . W_dino^T is computed once per clip by the caller (spa3d_gemm). */
int spa3d_embed_sampled(const float* xyz, const float* tracks_2d, const float* dfeat, const void* proj, const void* Wt, int64_t ldw,
                        const float* wdep, const float* bias, float* out, int64_t ldo, int64_t rows, int T, int Hp, int Wp,
                        int video_H, int video_W, int W, int num_freq, float track_scale_factor, void* stream) {
  using namespace spa3d::te;
  SPA3D_REQUIRE(num_freq == 32, "embed_sampled: 32 frequencies per coordinate (one 64-column K block each)");
  SPA3D_REQUIRE(W == 384 || W == 256 || W == 192 || W == 128, "embed_sampled: unsupported token width %d", W);
  SPA3D_REQUIRE(xyz && tracks_2d && proj && Wt && bias && out, "embed_sampled: NULL operand");
  SPA3D_REQUIRE((dfeat == nullptr) == (wdep == nullptr), "embed_sampled: dfeat and wdep come together");
  SPA3D_REQUIRE(Hp > 0 && Wp > 0 && video_H > 0 && video_W > 0 && T > 0, "embed_sampled: bad geometry");
  SPA3D_REQUIRE((int64_t)T * Hp * Wp * W < (1ll << 31), "embed_sampled: projected map too large for 32-bit offsets");
  auto al = [](const void* q, int a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  SPA3D_REQUIRE(al(Wt, 16) && al(out, 16) && al(bias, 16) && al(proj, 8) && al(tracks_2d, 8) && al(dfeat, 16) && al(wdep, 16) &&
                    ldw % 8 == 0 && ldo % 4 == 0, "embed_sampled: operand alignment");
  if (rows == 0) return 0;
  SPA3D_REQUIRE(rows < (1ll << 31) && rows % T == 0, "embed_sampled: rows must be tracks x T");
  EmbedParams p;
  p.tracks = xyz; p.dino = nullptr; p.depth = nullptr; p.bias = bias; p.out = out; p.ldo = ldo; p.R = rows;
  p.acat = nullptr; p.lda = 0;
  p.T = T; p.Dd = 0; p.Dz = 0; p.W = W; p.inv_scale = (float)(1.0 / (double)track_scale_factor); p.inv_T = (float)(1.0 / (double)T);
  for (int i = 0; i < 32; ++i) p.fs[i] = (float)pow(2.0, (double)i / 3.0);
  p.l2_prefetch = 0;
#ifdef SPA3D_EMBED_DEBUG
  p.debug = SPA3D_EMBED_DEBUG;   // timing experiments only, compile-time (1 = all gathers hit one patch row, 2 = no stores)
#else
  p.debug = 0;
#endif
  p.proj = reinterpret_cast<const bf16*>(proj); p.trk2d = tracks_2d; p.dfeat = dfeat; p.wdep = wdep; p.Hp = Hp; p.Wp = Wp;
  p.scale_w = (float)((double)Wp / (double)video_W);
  p.scale_h = (float)((double)Hp / (double)video_H);
  CUtensorMap tmB;
  const int bnh = W > 256 ? W / 2 : W;
  if (tc::make_map(&tmB, Wt, W, 256, ldw, bnh)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (W) {
    case 384: return launch_v<2, 192, false, true>(tmB, p, st);
    case 256: return launch_v<1, 256, false, true>(tmB, p, st);
    case 192: return launch_v<1, 192, false, true>(tmB, p, st);
    default: return launch_v<1, 128, false, true>(tmB, p, st);
  }
}

}  // extern "C"
