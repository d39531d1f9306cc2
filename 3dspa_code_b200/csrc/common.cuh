// Shared helpers for the 3DSPA B200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>
#include <cstdio>

#include "../../include/spa3d_b200.h"

namespace spa3d {

// ---- error plumbing (thread-local message, int status; never throws across the C ABI) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SPA3D_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      spa3d::set_error(__VA_ARGS__);  \
      return 1;                       \
    }                                 \
  } while (0)

// ---- dispatch counters (spa3d_stats): which implementation every contraction / attention call took ----
// A bf16 operand that does not meet the tensor-core kernels' alignment rules is still computed (SIMT), but 10-100x
// slower; the counters make that visible (bench.py asserts the *_bf16_fallback counters stay zero).
enum StatId {
  ST_GEMM_TCGEN05 = 0, ST_GEMM_X3, ST_GEMM_SIMT, ST_GEMM_BF16_FALLBACK,
  ST_GEMM_DW_TCGEN05, ST_GEMM_DW_SIMT, ST_GEMM_DW_BF16_FALLBACK,
  ST_ATTN_TCGEN05, ST_ATTN_CROSS_TCGEN05, ST_ATTN_Q1, ST_ATTN_SIMT, ST_ATTN_BF16_FALLBACK,
  ST_EMBED_FUSED, ST_COUNT
};
void stat_add(int id);

// ---- dtype helpers ----
using bf16 = __nv_bfloat16;

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// dispatch a generic lambda on a dtype code: f(T{}) with T in {float, bf16}
#define SPA3D_DISPATCH(code, T, ...)                               \
  do {                                                             \
    if ((code) == SPA3D_F32) {                                     \
      using T = float;                                             \
      __VA_ARGS__;                                                 \
    } else if ((code) == SPA3D_BF16) {                             \
      using T = spa3d::bf16;                                       \
      __VA_ARGS__;                                                 \
    } else {                                                       \
      spa3d::set_error("bad dtype code %d", (int)(code));          \
      return 1;                                                    \
    }                                                              \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// bilinear sampling set-up shared by the lifting kernel and the sampled embedding (inference.py:307-322)
struct Bilin {
  int x0, y0, x1, y1;
  float wx, wy;
};

__device__ __forceinline__ Bilin bilin_setup(float x, float y, int W, int H) {
  Bilin b;
  float fx = floorf(x), fy = floorf(y);
  int x0 = (int)fx, y0 = (int)fy;
  b.wx = __fsub_rn(x, fx);  // weights use the UNclamped floor (inference.py:312)
  b.wy = __fsub_rn(y, fy);
  b.x1 = min(max(x0 + 1, 0), W - 1);
  b.y1 = min(max(y0 + 1, 0), H - 1);
  b.x0 = min(max(x0, 0), W - 1);
  b.y0 = min(max(y0, 0), H - 1);
  return b;
}

// flax nn.gelu(approximate=True)
// Logit of a masked key in the bf16 attention kernels: bf16(-1e30) widened to fp32, the value the
// tcgen05 forward adds as a key bias (attention_tc.cu).  Any finite logit is absorbed by rounding, so
// masked keys are one constant as with the reference's finfo.min, and the backward kernels that
// recompute P from the saved row maximum must use the very same constant.
__device__ __forceinline__ float masked_logit_bf16() { return __uint_as_float(0xF14A0000u); }

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  float u = k * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k = 0.7978845608028654f;
  float x2 = x * x;
  float u = k * (x + 0.044715f * x * x2);
  float t = tanhf(u);
  float du = k * (1.0f + 3.0f * 0.044715f * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

constexpr float kNormEps = 1e-6f;  // flax LayerNorm / RMSNorm epsilon

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// internal entry points implemented in the other translation units
int gemm_simt(const void* A, int64_t sam, int64_t sak, int a_dtype, const void* B, int64_t sbk,
              int64_t sbn, int b_dtype, const float* bias, int act, const void* residual,
              int64_t ldr, int r_dtype, void* C, int64_t ldc, int c_dtype, int64_t M, int N,
              int64_t K, int accumulate, cudaStream_t st);
struct RmsEpilogue {  // fused per-head RMSNorm of the leading q_cols / k_cols output columns
  int dh, q_cols, k_cols;
  const float* scale_q;
  const float* scale_k;
  float q_mul;
  float* rstd_out;  // [M, (q_cols+k_cols)/dh] or null
};
int gemm_tcgen05(const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias,
                 int act, const void* residual, int64_t ldr, int r_dtype, void* C, int64_t ldc,
                 int c_dtype, int64_t M, int N, int K, const RmsEpilogue* rms, cudaStream_t st,
                 int res_op = 0, void* aux_pre = nullptr, int64_t ld_aux = 0, int aux_kind = 1, float* colsum = nullptr);
bool gemm_tcgen05_rms_applicable(int N, int dh, int q_cols, int k_cols, int c_dtype);
int head_rmsnorm_fwd_impl(void* buf, int64_t ld, int dtype, const float* scale, float out_mul,
                          float* rstd_out, int64_t rstd_ld, int64_t rows, int heads, int Dh,
                          cudaStream_t st);
bool gemm_tcgen05_dw_applicable(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M,
                                int N, int K, const void* dW, int64_t lddw);
int gemm_tcgen05_dw(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw,
                    int64_t M, int N, int K, cudaStream_t st);
bool gemm_tcgen05_x3_applicable(const void* A3, int64_t lda, const void* W3, int64_t ldw, int64_t M, int N, int K);
int gemm_tcgen05_x3(const void* A3, int64_t lda, const void* W3, int64_t ldw, const float* bias, const void* residual, int64_t ldr,
                    int r_dtype, float* C, int64_t ldc, int64_t M, int N, int K, cudaStream_t st);
bool gemm_tcgen05_applicable(const void* A, int64_t lda, const void* Wt, int64_t ldw, int64_t M,
                             int N, int K);

}  // namespace spa3d
