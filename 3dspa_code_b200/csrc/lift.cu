// K0: feature lifting - bilinear depth sample + pinhole unprojection, bilinear DINO patch
// sample, 256-channel depth feature.  Replaces the Python double loops of
// /root/reference/inference.py:287-336, :339-395, :398-447 with one bandwidth-bound gather
// kernel.  Arithmetic is float32 in the reference's operation order with explicit
// round-to-nearest intrinsics (no FMA contraction), so results are bit-identical to the NumPy
// code for float32 inputs.
//
// Two forms (same arithmetic, same bits): the per-point kernel - one warp per track point, points enumerated frame-major
// (p = t*N + n) so the CTAs in flight gather from the same frame's DINO map (37x37x768 f32 = 4.2 MB, L2-resident), lanes stride
// over channels with 128-bit loads / stores - and, below it, the cell-binned form the product uses when the patch features are
// requested: the points of a frame are sorted by patch cell and a cell's four corner rows are read once for all of its points.
#include <algorithm>

#include "common.cuh"

namespace spa3d {

// z00*(1-wx)*(1-wy) + z01*wx*(1-wy) + z10*(1-wx)*wy + z11*wx*wy, left to right
__device__ __forceinline__ float blend(float v00, float v01, float v10, float v11, float wx, float wy,
                                       float omx, float omy) {
  float a = __fmul_rn(__fmul_rn(v00, omx), omy);
  float b = __fmul_rn(__fmul_rn(v01, wx), omy);
  float c = __fmul_rn(__fmul_rn(v10, omx), wy);
  float d = __fmul_rn(__fmul_rn(v11, wx), wy);
  return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
}

__device__ __forceinline__ float sample_depth(const float* __restrict__ depth, int t, int H, int W,
                                              float x, float y) {
  Bilin b = bilin_setup(x, y, W, H);
  const float* dp = depth + (int64_t)t * H * W;
  float omx = __fsub_rn(1.f, b.wx), omy = __fsub_rn(1.f, b.wy);
  return blend(dp[b.y0 * W + b.x0], dp[b.y0 * W + b.x1], dp[b.y1 * W + b.x0], dp[b.y1 * W + b.x1],
               b.wx, b.wy, omx, omy);
}

template <typename TO>
__device__ __forceinline__ void store4(TO* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
#ifdef SPA3D_LIFT_PLAIN_STORES
  *reinterpret_cast<float4*>(p) = v;
#else
  __stcs(reinterpret_cast<float4*>(p), v);   // streaming: written once, keep the L2 for the DINO maps
#endif
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  __stcs(reinterpret_cast<uint2*>(p), u);
}

template <typename TO>
__global__ void __launch_bounds__(256)
lift_sample_kernel(const float* __restrict__ tracks, const float* __restrict__ depth,
                   const float* __restrict__ dino, float* __restrict__ xyz, TO* __restrict__ dino_out,
                   TO* __restrict__ depth_out, int N, int T, int H, int W, int Hp, int Wp, int D,
                   int Cd, float scale_w, float scale_h, float fx, float fy, float cx, float cy) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= (int64_t)N * T) return;
  const int t = (int)(p / N), n = (int)(p % N);
  const int64_t pt = (int64_t)n * T + t;
  const float x = tracks[pt * 2], y = tracks[pt * 2 + 1];

  if (depth != nullptr) {
    float z = sample_depth(depth, t, H, W, x, y);
    if (xyz != nullptr && lane == 0) {
      xyz[pt * 3 + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(x, cx), z), fx);
      xyz[pt * 3 + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(y, cy), z), fy);
      xyz[pt * 3 + 2] = z;
    }
    if (depth_out != nullptr) {
      float grad = 0.f;
      if (t > 0) {
        float zp = sample_depth(depth, t - 1, H, W, tracks[(pt - 1) * 2], tracks[(pt - 1) * 2 + 1]);
        grad = __fsub_rn(z, zp);
      }
      TO* o = depth_out + pt * Cd;
      if ((Cd & 3) == 0) {   // 128-bit stores: (z, z/10, dz, 0) then zeros (inference.py:437-443)
        for (int c = lane * 4; c < Cd; c += 128)
          store4<TO>(o + c, c == 0 ? make_float4(z, __fdiv_rn(z, 10.f), grad, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
      } else {
        for (int c = lane; c < Cd; c += 32) {
          float v = c == 0 ? z : (c == 1 ? __fdiv_rn(z, 10.f) : (c == 2 ? grad : 0.f));
          stf<TO>(o + c, v);
        }
      }
    }
  }

  if (dino != nullptr && dino_out != nullptr) {
    float px = __fmul_rn(x, scale_w), py = __fmul_rn(y, scale_h);
    Bilin b = bilin_setup(px, py, Wp, Hp);
    float omx = __fsub_rn(1.f, b.wx), omy = __fsub_rn(1.f, b.wy);
    const float* base = dino + (int64_t)t * Hp * Wp * D;
    const float* f00 = base + ((int64_t)b.y0 * Wp + b.x0) * D;
    const float* f01 = base + ((int64_t)b.y0 * Wp + b.x1) * D;
    const float* f10 = base + ((int64_t)b.y1 * Wp + b.x0) * D;
    const float* f11 = base + ((int64_t)b.y1 * Wp + b.x1) * D;
    TO* o = dino_out + pt * D;
    if ((D & 255) == 0) {
      // two 128-channel groups per trip (768 channels = 3 trips): eight independent 16-byte gathers are in flight per lane before
      // the first blend - the kernel is bound by the latency of these L2 gathers (ncu: 74 % of the samples on the long
      // scoreboard with four in flight), not by the bytes it writes
      for (int c = lane * 4; c < D; c += 256) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(f00 + c));
        const float4 bq = __ldg(reinterpret_cast<const float4*>(f01 + c));
        const float4 cq = __ldg(reinterpret_cast<const float4*>(f10 + c));
        const float4 d = __ldg(reinterpret_cast<const float4*>(f11 + c));
        const float4 a2 = __ldg(reinterpret_cast<const float4*>(f00 + c + 128));
        const float4 bq2 = __ldg(reinterpret_cast<const float4*>(f01 + c + 128));
        const float4 cq2 = __ldg(reinterpret_cast<const float4*>(f10 + c + 128));
        const float4 d2 = __ldg(reinterpret_cast<const float4*>(f11 + c + 128));
        float4 r, r2;
        r.x = blend(a.x, bq.x, cq.x, d.x, b.wx, b.wy, omx, omy);
        r.y = blend(a.y, bq.y, cq.y, d.y, b.wx, b.wy, omx, omy);
        r.z = blend(a.z, bq.z, cq.z, d.z, b.wx, b.wy, omx, omy);
        r.w = blend(a.w, bq.w, cq.w, d.w, b.wx, b.wy, omx, omy);
        r2.x = blend(a2.x, bq2.x, cq2.x, d2.x, b.wx, b.wy, omx, omy);
        r2.y = blend(a2.y, bq2.y, cq2.y, d2.y, b.wx, b.wy, omx, omy);
        r2.z = blend(a2.z, bq2.z, cq2.z, d2.z, b.wx, b.wy, omx, omy);
        r2.w = blend(a2.w, bq2.w, cq2.w, d2.w, b.wx, b.wy, omx, omy);
        store4<TO>(o + c, r);
        store4<TO>(o + c + 128, r2);
      }
    } else if ((D & 3) == 0) {
      for (int c = lane * 4; c < D; c += 128) {
        float4 a = __ldg(reinterpret_cast<const float4*>(f00 + c));
        float4 bq = __ldg(reinterpret_cast<const float4*>(f01 + c));
        float4 cq = __ldg(reinterpret_cast<const float4*>(f10 + c));
        float4 d = __ldg(reinterpret_cast<const float4*>(f11 + c));
        float4 r;
        r.x = blend(a.x, bq.x, cq.x, d.x, b.wx, b.wy, omx, omy);
        r.y = blend(a.y, bq.y, cq.y, d.y, b.wx, b.wy, omx, omy);
        r.z = blend(a.z, bq.z, cq.z, d.z, b.wx, b.wy, omx, omy);
        r.w = blend(a.w, bq.w, cq.w, d.w, b.wx, b.wy, omx, omy);
        store4<TO>(o + c, r);
      }
    } else {
      for (int c = lane; c < D; c += 32)
        stf<TO>(o + c, blend(f00[c], f01[c], f10[c], f11[c], b.wx, b.wy, omx, omy));
    }
  }
}

// ---- geometry only: one THREAD per point -------------------------------------------------------------------------
// The fused maps path (spa3d_embed_sampled) needs the lifted coordinates and the narrow depth feature (d, d/10, d_t - d_{t-1}, 0)
// but no per-point patch features: a warp per point would spend its time on one lane's chain of dependent gathers.  Points are
// enumerated frame-major like above; a thread writes its point's xyz and its Cd <= 8 feature values itself.
template <typename TO>
__global__ void __launch_bounds__(256)
lift_points_kernel(const float* __restrict__ tracks, const float* __restrict__ depth, float* __restrict__ xyz, TO* __restrict__ depth_out,
                   int N, int T, int H, int W, int Cd, float fx, float fy, float cx, float cy) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)N * T) return;
  const int t = (int)(p / N);
  const int64_t pt = (int64_t)(p % N) * T + t;
  const float x = tracks[pt * 2], y = tracks[pt * 2 + 1];
  const float z = sample_depth(depth, t, H, W, x, y);
  if (xyz != nullptr) {
    xyz[pt * 3 + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(x, cx), z), fx);
    xyz[pt * 3 + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(y, cy), z), fy);
    xyz[pt * 3 + 2] = z;
  }
  if (depth_out != nullptr) {
    float grad = 0.f;
    if (t > 0) grad = __fsub_rn(z, sample_depth(depth, t - 1, H, W, tracks[(pt - 1) * 2], tracks[(pt - 1) * 2 + 1]));
    TO* o = depth_out + pt * Cd;
    stf<TO>(o, z);
    stf<TO>(o + 1, __fdiv_rn(z, 10.f));
    stf<TO>(o + 2, grad);
    for (int c = 3; c < Cd; ++c) stf<TO>(o + c, 0.f);
  }
}

// ---- cell-binned form of the same gather ----------------------------------------------------------------------
// What bounds the kernel above is the 4x read amplification of bilinear sampling on the L2 -> SM path: every point pulls its
// four corner rows (4 x D floats = 12 KB at D = 768) through the L1 for 3 KB of output - 9.4 GB of gathers per cfg4 clip, of
// which the L1 absorbs a third (ncu: 4.5 GB L2 -> L1, 74 % of the samples waiting on those loads, L2 / DRAM at 38 % / 51 %).
// Measured and not kept on the way here: a persistent grid (same 0.78 ms), eight gathers in flight per lane (0.77), one scalar
// phase per 8 points (0.87) and a cp.async.bulk version staging the four rows of every point in shared memory, one point
// ahead (1.58 ms: no L1 reuse at all, 7.4 GB through the L2).
// All points of a frame whose sample falls into the same patch cell share the four corner rows.  A pre-pass bins the points
// of every frame by cell (counting sort in shared memory, one CTA per frame); the gather then walks (frame, cell) items: a
// warp loads the cell's four corner rows ONCE into registers and blends them for every point of the cell, so the corner rows
// are read once per occupied cell instead of once per point.  Same per-point arithmetic, same bits.
// Cell key: the UNclamped floor of the sample position clamped to [-1, Wp-1] x [-1, Hp-1] - every point with that key has the
// same clamped corners (x0, x1, y0, y1) in bilin_setup.
__device__ __forceinline__ int lift_cell(float x, float y, float scale_w, float scale_h, int Wp, int Hp) {
  const int kx = min(max((int)floorf(__fmul_rn(x, scale_w)), -1), Wp - 1);
  const int ky = min(max((int)floorf(__fmul_rn(y, scale_h)), -1), Hp - 1);
  return (ky + 1) * (Wp + 1) + (kx + 1);
}

// one CTA per frame: rec[t*N + i] = the frame's points sorted by cell, one 32-byte record each - (n, x, y, x_prev, y_prev): what the
// gather's scalar phase needs, so that it starts from ONE coalesced load instead of the chain order -> track coordinates ->
// previous-frame coordinates; starts[t*(cells+1) + c] = first position of cell c
struct __align__(16) LiftRec {
  int n;
  float x, y, xp, yp;
  int pad[3];
};
static_assert(sizeof(LiftRec) == 32, "record size");

__global__ void __launch_bounds__(256)
lift_bin_kernel(const float* __restrict__ tracks, LiftRec* __restrict__ rec, int* __restrict__ starts, int N, int T, int Hp, int Wp,
                float scale_w, float scale_h) {
  extern __shared__ int bin_smem[];
  const int cells = (Wp + 1) * (Hp + 1);
  int* hist = bin_smem;            // [cells]: counts, then running cursors
  int* part = bin_smem + cells;    // [256]: per-thread chunk sums
  const int t = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < cells; c += 256) hist[c] = 0;
  __syncthreads();
  for (int n = tid; n < N; n += 256) {
    const int64_t pt = (int64_t)n * T + t;
    atomicAdd(&hist[lift_cell(tracks[pt * 2], tracks[pt * 2 + 1], scale_w, scale_h, Wp, Hp)], 1);
  }
  __syncthreads();
  // exclusive scan: every thread owns a run of consecutive cells
  const int per = (cells + 255) / 256, c0 = min(tid * per, cells), c1 = min(c0 + per, cells);
  int sum = 0;
  for (int c = c0; c < c1; ++c) sum += hist[c];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; }
  }
  __syncthreads();
  int run = part[tid];
  int* st = starts + (int64_t)t * (cells + 1);
  for (int c = c0; c < c1; ++c) {
    const int v = hist[c];
    hist[c] = run;
    st[c] = run;
    run += v;
  }
  if (tid == 0) st[cells] = N;
  __syncthreads();
  for (int n = tid; n < N; n += 256) {
    const int64_t pt = (int64_t)n * T + t;
    const float x = tracks[pt * 2], y = tracks[pt * 2 + 1];
    const int pos = atomicAdd(&hist[lift_cell(x, y, scale_w, scale_h, Wp, Hp)], 1);
    uint4 lo, hi = make_uint4(0, 0, 0, 0);
    lo.x = (uint32_t)n;
    lo.y = __float_as_uint(x);
    lo.z = __float_as_uint(y);
    lo.w = t > 0 ? __float_as_uint(tracks[(pt - 1) * 2]) : 0u;
    hi.x = t > 0 ? __float_as_uint(tracks[(pt - 1) * 2 + 1]) : 0u;
    uint4* dst = reinterpret_cast<uint4*>(rec + (int64_t)t * N + pos);
    dst[0] = lo;
    dst[1] = hi;
  }
}

#ifndef LIFT_BINNED_CTAS
#define LIFT_BINNED_CTAS 2
#endif
// NJ 128-channel groups per pass (D is walked in passes of 128 * NJ channels; 4 * NJ float4 of corner rows live in registers)
template <typename TO, int NJ>
__global__ void __launch_bounds__(256, LIFT_BINNED_CTAS)
lift_sample_binned_kernel(const float* __restrict__ tracks, const float* __restrict__ depth, const float* __restrict__ dino,
                          const LiftRec* __restrict__ rec, const int* __restrict__ starts, float* __restrict__ xyz,
                          TO* __restrict__ dino_out, TO* __restrict__ depth_out, int N, int T, int H, int W, int Hp, int Wp, int D,
                          int Cd, float scale_w, float scale_h, float fx, float fy, float cx, float cy) {
  const int lane = threadIdx.x & 31;
  const int cells = (Wp + 1) * (Hp + 1);
  const int64_t items = (int64_t)T * cells, nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < items; it += nw) {
    const int t = (int)(it / cells), cell = (int)(it % cells);
    const int* st = starts + (int64_t)t * (cells + 1) + cell;
    const int s0 = __ldg(st), s1 = __ldg(st + 1);
    if (s0 == s1) continue;
    const int kx = cell % (Wp + 1) - 1, ky = cell / (Wp + 1) - 1;
    const int x0 = min(max(kx, 0), Wp - 1), x1 = min(max(kx + 1, 0), Wp - 1), y0 = min(max(ky, 0), Hp - 1), y1 = min(max(ky + 1, 0), Hp - 1);
    const float* base = dino + (int64_t)t * Hp * Wp * D;
    const float* f00 = base + ((int64_t)y0 * Wp + x0) * D;
    const float* f01 = base + ((int64_t)y0 * Wp + x1) * D;
    const float* f10 = base + ((int64_t)y1 * Wp + x0) * D;
    const float* f11 = base + ((int64_t)y1 * Wp + x1) * D;
    // the corner rows of the first pass are requested before the scalar phase: its chain of dependent loads (point record ->
    // depth gathers of both frames) and the row gathers are then in flight together
    float4 a[NJ], bq[NJ], cq[NJ], d[NJ];
    auto load_rows = [&](int cb) {
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const int c = cb + jj * 128 + lane * 4;
        a[jj] = __ldg(reinterpret_cast<const float4*>(f00 + c));
        bq[jj] = __ldg(reinterpret_cast<const float4*>(f01 + c));
        cq[jj] = __ldg(reinterpret_cast<const float4*>(f10 + c));
        d[jj] = __ldg(reinterpret_cast<const float4*>(f11 + c));
      }
    };
    load_rows(0);
    for (int b0 = s0; b0 < s1; b0 += 32) {
      // ---- scalar phase: lane j owns point b0 + j of the cell ----
      const int m = min(32, s1 - b0);
      int n = 0;
      float x = 0.f, y = 0.f, z = 0.f, grad = 0.f, wx = 0.f, wy = 0.f;
      if (lane < m) {
        const uint4* rp = reinterpret_cast<const uint4*>(rec + (int64_t)t * N + b0 + lane);
        const uint4 lo = __ldg(rp);
        const uint32_t ypb = __ldg(reinterpret_cast<const uint32_t*>(rp + 1));
        n = (int)lo.x;
        x = __uint_as_float(lo.y);
        y = __uint_as_float(lo.z);
        const int64_t pt = (int64_t)n * T + t;
        const Bilin b = bilin_setup(__fmul_rn(x, scale_w), __fmul_rn(y, scale_h), Wp, Hp);
        wx = b.wx;
        wy = b.wy;
        if (depth != nullptr) {
          // both frames' gathers are independent of each other: issued together
          const float zc = sample_depth(depth, t, H, W, x, y);
          const float zp = (depth_out != nullptr && t > 0) ? sample_depth(depth, t - 1, H, W, __uint_as_float(lo.w), __uint_as_float(ypb)) : 0.f;
          z = zc;
          if (xyz != nullptr) {
            xyz[pt * 3 + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(x, cx), z), fx);
            xyz[pt * 3 + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(y, cy), z), fy);
            xyz[pt * 3 + 2] = z;
          }
          if (depth_out != nullptr && t > 0) grad = __fsub_rn(z, zp);
        }
      }
      // ---- depth-feature rows: (z, z/10, dz, 0) then zeros (inference.py:437-443) ----
      if (depth != nullptr && depth_out != nullptr) {
        for (int j = 0; j < m; ++j) {
          const float zj = __shfl_sync(0xffffffffu, z, j), gj = __shfl_sync(0xffffffffu, grad, j);
          TO* o = depth_out + ((int64_t)__shfl_sync(0xffffffffu, n, j) * T + t) * Cd;
          for (int c = lane * 4; c < Cd; c += 128)
            store4<TO>(o + c, c == 0 ? make_float4(zj, __fdiv_rn(zj, 10.f), gj, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
      }
      // ---- feature rows: the cell's corner rows are loaded once per pass and blended for each of its points ----
      for (int cb = 0; cb < D; cb += 128 * NJ) {
        if (cb != 0 || b0 != s0) load_rows(cb);
        for (int j = 0; j < m; ++j) {
          const float wxj = __shfl_sync(0xffffffffu, wx, j), wyj = __shfl_sync(0xffffffffu, wy, j);
          const float omx = __fsub_rn(1.f, wxj), omy = __fsub_rn(1.f, wyj);
          TO* o = dino_out + ((int64_t)__shfl_sync(0xffffffffu, n, j) * T + t) * D + cb + lane * 4;
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) {
            float4 r;
            r.x = blend(a[jj].x, bq[jj].x, cq[jj].x, d[jj].x, wxj, wyj, omx, omy);
            r.y = blend(a[jj].y, bq[jj].y, cq[jj].y, d[jj].y, wxj, wyj, omx, omy);
            r.z = blend(a[jj].z, bq[jj].z, cq[jj].z, d[jj].z, wxj, wyj, omx, omy);
            r.w = blend(a[jj].w, bq[jj].w, cq[jj].w, d[jj].w, wxj, wyj, omx, omy);
            store4<TO>(o + jj * 128, r);
          }
        }
      }
    }
  }
}

}  // namespace spa3d

namespace spa3d {
static int lift_sample_impl(const float* tracks_2d, const float* depth, const float* dino, float* xyz, void* dino_out, void* depth_out,
                            int out_dtype, int N, int T, int H, int W, int Hp, int Wp, int D, int Cd, int video_H, int video_W,
                            const float* intrinsics, void* workspace, int64_t workspace_bytes, void* stream) {
  int64_t pts = (int64_t)N * T;
  if (pts == 0) return 0;
  SPA3D_REQUIRE(tracks_2d != nullptr, "lift_sample: tracks_2d is NULL");
  SPA3D_REQUIRE(depth == nullptr || (H > 0 && W > 0), "lift_sample: bad depth shape");
  SPA3D_REQUIRE(dino == nullptr || (Hp > 0 && Wp > 0 && D > 0 && video_H > 0 && video_W > 0),
                "lift_sample: bad dino shape");
  SPA3D_REQUIRE(depth_out == nullptr || Cd >= 3, "lift_sample: depth feature width must be >= 3");
  // scale_w = W_patches / W is a Python float in the reference; the multiply happens in f32
  float scale_w = dino ? (float)((double)Wp / (double)video_W) : 0.f;
  float scale_h = dino ? (float)((double)Hp / (double)video_H) : 0.f;
  float fx, fy, cx, cy;
  if (intrinsics) {
    fx = intrinsics[0]; fy = intrinsics[1]; cx = intrinsics[2]; cy = intrinsics[3];
  } else {  // inference.py:303-306
    fx = fy = (float)(H > W ? H : W);
    cx = (float)((double)W / 2.0);
    cy = (float)((double)H / 2.0);
  }
  cudaStream_t st = (cudaStream_t)stream;
  // cell-binned form: needs the patch features, channel groups of 128, vector-aligned rows and the caller's workspace
  static int use_binned = -1;   // A/B switch for measurements: SPA3D_LIFT_BINNED=0 keeps the per-point kernel (both are product kernels)
  if (use_binned < 0) {
    const char* e = getenv("SPA3D_LIFT_BINNED");
    use_binned = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int64_t cells = (int64_t)(Wp + 1) * (Hp + 1);
  const bool binned = use_binned && dino && dino_out && D % 128 == 0 && (depth_out == nullptr || Cd % 4 == 0) && workspace != nullptr &&
                      workspace_bytes >= spa3d_lift_workspace_bytes(N, T, Hp, Wp) && (cells + 256) * 4 <= 200 * 1024 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0;
  if (binned) {
    LiftRec* rec = reinterpret_cast<LiftRec*>(workspace);
    int* starts = reinterpret_cast<int*>(rec + pts);
    const size_t bin_smem = (size_t)(cells + 256) * 4;
    static bool attr_set = false;
    if (!attr_set && bin_smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(lift_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      SPA3D_REQUIRE(e == cudaSuccess, "lift_sample: smem attribute: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    lift_bin_kernel<<<T, 256, bin_smem, st>>>(tracks_2d, rec, starts, N, T, Hp, Wp, scale_w, scale_h);
    if (check_launch("lift_bin")) return 1;
    const int sms = num_sms();
    const int64_t items = (int64_t)T * cells;
    const unsigned grid = (unsigned)std::min<int64_t>((items + 7) / 8, (int64_t)sms * LIFT_BINNED_CTAS);
    // measured on the cfg4 shape (D = 768): 3 groups per pass at two CTAs per SM 0.624 ms; 2 groups 0.643; 1 group 0.71 - 0.84; all six
    // groups in one pass (168 registers, 12 warps per SM) 0.633 - 0.645; three CTAs per SM (80 registers) 0.638 with 2 groups
    const int nj = D % 384 == 0 ? 3 : (D % 256 == 0 ? 2 : 1);
    SPA3D_DISPATCH(out_dtype, TO, {
      if (nj == 3)
        lift_sample_binned_kernel<TO, 3><<<grid, 256, 0, st>>>(tracks_2d, depth, dino, rec, starts, xyz, (TO*)dino_out, (TO*)depth_out, N, T, H, W, Hp, Wp, D, Cd, scale_w, scale_h, fx, fy, cx, cy);
      else if (nj == 2)
        lift_sample_binned_kernel<TO, 2><<<grid, 256, 0, st>>>(tracks_2d, depth, dino, rec, starts, xyz, (TO*)dino_out, (TO*)depth_out, N, T, H, W, Hp, Wp, D, Cd, scale_w, scale_h, fx, fy, cx, cy);
      else
        lift_sample_binned_kernel<TO, 1><<<grid, 256, 0, st>>>(tracks_2d, depth, dino, rec, starts, xyz, (TO*)dino_out, (TO*)depth_out, N, T, H, W, Hp, Wp, D, Cd, scale_w, scale_h, fx, fy, cx, cy);
    });
    return check_launch("lift_sample_binned");
  }
  if (depth != nullptr && (dino == nullptr || dino_out == nullptr) && (depth_out == nullptr || Cd <= 8)) {   // geometry only
    SPA3D_DISPATCH(out_dtype, TO, {
      lift_points_kernel<TO><<<(unsigned)((pts + 255) / 256), 256, 0, st>>>(tracks_2d, depth, xyz, (TO*)depth_out, N, T, H, W, Cd, fx, fy, cx, cy);
    });
    return check_launch("lift_points");
  }
  unsigned blocks = (unsigned)((pts + 7) / 8);
  SPA3D_DISPATCH(out_dtype, TO, {
    lift_sample_kernel<TO><<<blocks, 256, 0, st>>>(tracks_2d, depth, dino, xyz, (TO*)dino_out, (TO*)depth_out, N, T, H, W, Hp, Wp, D, Cd, scale_w, scale_h, fx, fy, cx, cy);
  });
  return check_launch("lift_sample");
}
}  // namespace spa3d

extern "C" int64_t spa3d_lift_workspace_bytes(int N, int T, int Hp, int Wp) {
  return 32 * (int64_t)N * T + 4 * (int64_t)T * ((int64_t)(Wp + 1) * (Hp + 1) + 1);   // point records + cell offsets per frame
}

extern "C" int spa3d_lift_sample(const float* tracks_2d, const float* depth, const float* dino,
                                 float* xyz, void* dino_out, void* depth_out, int out_dtype, int N,
                                 int T, int H, int W, int Hp, int Wp, int D, int Cd, int video_H,
                                 int video_W, const float* intrinsics, void* stream) {
  return spa3d::lift_sample_impl(tracks_2d, depth, dino, xyz, dino_out, depth_out, out_dtype, N, T, H, W, Hp, Wp, D, Cd, video_H, video_W,
                                 intrinsics, nullptr, 0, stream);
}

extern "C" int spa3d_lift_sample_ws(const float* tracks_2d, const float* depth, const float* dino,
                                    float* xyz, void* dino_out, void* depth_out, int out_dtype, int N,
                                    int T, int H, int W, int Hp, int Wp, int D, int Cd, int video_H,
                                    int video_W, const float* intrinsics, void* workspace, int64_t workspace_bytes, void* stream) {
  return spa3d::lift_sample_impl(tracks_2d, depth, dino, xyz, dino_out, depth_out, out_dtype, N, T, H, W, Hp, Wp, D, Cd, video_H, video_W,
                                 intrinsics, workspace, workspace_bytes, stream);
}
