// fp32-accumulate SIMT GEMM with arbitrary operand strides.
//
// This is the "fp32-accumulate" accurate mode of the hot path (north_star: 1e-4 parity with the
// reference's float32 arithmetic) and the generic fallback used by the backward pass for operand
// layouts the tcgen05 kernel does not take.  It is NOT the throughput path: bf16 contractions
// go through gemm_tcgen05.cu.
#include "common.cuh"

namespace spa3d {

constexpr int SBM = 64, SBN = 64, SBK = 16, SPAD = 4;

template <typename TA, typename TB, typename TC, typename TR, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sam, int64_t sak, const TB* __restrict__ B,
                 int64_t sbk, int64_t sbn, const float* __restrict__ bias, int act,
                 const TR* __restrict__ residual, int64_t ldr, TC* __restrict__ C, int64_t ldc,
                 int64_t M, int N, int64_t K, int accumulate) {
  __shared__ __align__(16) float As[SBK][SBM + SPAD];
  __shared__ __align__(16) float Bs[SBK][SBN + SPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid * 4 + i;
      int m, k;
      if (A_KC) { k = idx % SBK; m = idx / SBK; } else { m = idx % SBM; k = idx / SBM; }
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = ldf<TA>(A + gm * sam + gk * sak);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid * 4 + i;
      int n, k;
      if (B_NC) { n = idx % SBN; k = idx / SBN; } else { k = idx % SBK; n = idx / SBK; }
      int64_t gk = k0 + k;
      int gn = n0 + n;
      float v = 0.f;
      if (gn < N && gk < K) v = ldf<TB>(B + gk * sbk + (int64_t)gn * sbn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (accumulate) v += ldf<TC>(C + gm * ldc + gn);
      if (bias) v += bias[gn];
      if (act == SPA3D_ACT_GELU_TANH) v = gelu_tanh(v);
      if (residual) v += ldf<TR>(residual + gm * ldr + gn);
      stf<TC>(C + gm * ldc + gn, v);
    }
  }
}

template <typename TA, typename TB, typename TC, typename TR>
static int launch_simt(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbk,
                       int64_t sbn, const float* bias, int act, const void* residual, int64_t ldr,
                       void* C, int64_t ldc, int64_t M, int N, int64_t K, int accumulate,
                       cudaStream_t st) {
  dim3 grid((unsigned)((M + SBM - 1) / SBM), (unsigned)((N + SBN - 1) / SBN));
  bool akc = (sak == 1), bnc = (sbn == 1);
#define L(AK, BN_)                                                                              \
  gemm_simt_kernel<TA, TB, TC, TR, AK, BN_><<<grid, 256, 0, st>>>(                              \
      (const TA*)A, sam, sak, (const TB*)B, sbk, sbn, bias, act, (const TR*)residual, ldr, (TC*)C, \
      ldc, M, N, K, accumulate)
  if (akc && bnc) L(true, true);
  else if (akc) L(true, false);
  else if (bnc) L(false, true);
  else L(false, false);
#undef L
  return check_launch("gemm_simt");
}

int gemm_simt(const void* A, int64_t sam, int64_t sak, int a_dtype, const void* B, int64_t sbk,
              int64_t sbn, int b_dtype, const float* bias, int act, const void* residual,
              int64_t ldr, int r_dtype, void* C, int64_t ldc, int c_dtype, int64_t M, int N,
              int64_t K, int accumulate, cudaStream_t st) {
  if (!residual) r_dtype = SPA3D_F32;
  SPA3D_DISPATCH(a_dtype, TA, SPA3D_DISPATCH(b_dtype, TB, SPA3D_DISPATCH(c_dtype, TC, SPA3D_DISPATCH(r_dtype, TR, {
    return launch_simt<TA, TB, TC, TR>(A, sam, sak, B, sbk, sbn, bias, act, residual, ldr, C, ldc, M, N, K, accumulate, st);
  }))));
  return 0;
}

}  // namespace spa3d
