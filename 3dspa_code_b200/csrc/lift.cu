// K0: feature lifting - bilinear depth sample + pinhole unprojection, bilinear DINO patch
// sample, 256-channel depth feature.  Replaces the Python double loops of
// /root/reference/inference.py:287-336, :339-395, :398-447 with one bandwidth-bound gather
// kernel.  Arithmetic is float32 in the reference's operation order with explicit
// round-to-nearest intrinsics (no FMA contraction), so results are bit-identical to the NumPy
// code for float32 inputs.
//
// Mapping: one warp per track point; points are enumerated frame-major (p = t*N + n) so the
// CTAs in flight at any moment gather from the same frame's DINO map (37x37x768 f32 = 4.2 MB),
// which therefore stays L2-resident; lanes stride over channels with 128-bit loads/stores.
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
  __stcs(reinterpret_cast<float4*>(p), v);   // streaming: written once, keep the L2 for the DINO maps
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

}  // namespace spa3d

extern "C" int spa3d_lift_sample(const float* tracks_2d, const float* depth, const float* dino,
                                 float* xyz, void* dino_out, void* depth_out, int out_dtype, int N,
                                 int T, int H, int W, int Hp, int Wp, int D, int Cd, int video_H,
                                 int video_W, const float* intrinsics, void* stream) {
  using namespace spa3d;
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
  unsigned blocks = (unsigned)((pts + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_DISPATCH(out_dtype, TO, {
    lift_sample_kernel<TO><<<blocks, 256, 0, st>>>(tracks_2d, depth, dino, xyz, (TO*)dino_out, (TO*)depth_out, N, T, H, W, Hp, Wp, D, Cd, scale_w, scale_h, fx, fy, cx, cy);
  });
  return check_launch("lift_sample");
}
