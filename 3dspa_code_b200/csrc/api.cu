// C-ABI plumbing: version, thread-local error string, GEMM dispatch.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace spa3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_stats[ST_COUNT];
static const char* const kStatNames[ST_COUNT] = {
    "gemm_tcgen05", "gemm_x3", "gemm_simt", "gemm_bf16_fallback", "gemm_dw_tcgen05", "gemm_dw_simt", "gemm_dw_bf16_fallback",
    "attention_tcgen05", "attention_cross_tcgen05", "attention_q1", "attention_simt", "attention_bf16_fallback",
    "embed_fused"};
void stat_add(int id) { __atomic_fetch_add(&g_stats[id], 1ll, __ATOMIC_RELAXED); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // namespace spa3d

extern "C" {

int spa3d_version(void) { return 100; }

const char* spa3d_last_error(void) { return spa3d::g_err; }

int spa3d_stats(int64_t* out, int n) {
  for (int i = 0; i < n && i < spa3d::ST_COUNT; ++i) out[i] = (int64_t)__atomic_load_n(&spa3d::g_stats[i], __ATOMIC_RELAXED);
  return spa3d::ST_COUNT;
}
int spa3d_stats_reset(void) {
  for (int i = 0; i < spa3d::ST_COUNT; ++i) __atomic_store_n(&spa3d::g_stats[i], 0ll, __ATOMIC_RELAXED);
  return 0;
}
const char* spa3d_stat_name(int i) { return (i >= 0 && i < spa3d::ST_COUNT) ? spa3d::kStatNames[i] : ""; }

// scratch sizes: the library never allocates; these are the buffers a caller passes in
int64_t spa3d_gemm_workspace_bytes(int64_t M, int N, int K) { (void)M; (void)N; (void)K; return 0; }   // split-K partials are added atomically into C
int64_t spa3d_sumsq_workspace_bytes(void) { return (int64_t)SPA3D_SUMSQ_WORKSPACE * 4; }
int64_t spa3d_attention_bwd_workspace_bytes(int64_t batch, int heads, int Lq) { return batch * heads * (int64_t)Lq * 4; }   // delta
int64_t spa3d_attention_stats_bytes(int64_t batch, int heads, int Lq) { return batch * heads * (int64_t)Lq * 2 * 4; }      // (max, sum) per row
int64_t spa3d_layernorm_bwd_workspace_bytes(int num_partials, int d) { return (int64_t)num_partials * d * 4; }             // d-scale partial rows

int spa3d_gemm(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
               const float* bias, int act, const void* residual, int64_t ldr, int r_dtype,
               void* C, int64_t ldc, int c_dtype, int64_t M, int N, int K, int impl,
               void* stream) {
  using namespace spa3d;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  if (M == 0) return 0;
  // the tcgen05 epilogue evaluates GELU with the hardware tanh (2^-11): fine for bf16 outputs only
  bool tc_ok = (a_dtype == SPA3D_BF16) && gemm_tcgen05_applicable(A, lda, Wt, ldw, M, N, K) &&
               !(act == SPA3D_ACT_GELU_TANH && c_dtype == SPA3D_F32);
  if (impl == SPA3D_GEMM_TCGEN05) {
    SPA3D_REQUIRE(tc_ok, "gemm: tcgen05 path not applicable (dtype=%d lda=%lld ldw=%lld K=%d)",
                  a_dtype, (long long)lda, (long long)ldw, K);
  }
  if (impl != SPA3D_GEMM_SIMT && tc_ok) {
    stat_add(ST_GEMM_TCGEN05);
    return gemm_tcgen05(A, lda, Wt, ldw, bias, act, residual, ldr, r_dtype, C, ldc, c_dtype, M, N,
                        K, nullptr, st);
  }
  // SIMT fp32-accumulate path: B(k,n) = Wt[n*ldw + k]
  stat_add(ST_GEMM_SIMT);
  if (impl == SPA3D_GEMM_AUTO && a_dtype == SPA3D_BF16) stat_add(ST_GEMM_BF16_FALLBACK);   // bf16 operands that missed the tensor-core path
  return gemm_simt(A, lda, 1, a_dtype, Wt, 1, ldw, a_dtype, bias, act, residual, ldr, r_dtype, C,
                   ldc, c_dtype, M, N, K, 0, st);
}

int spa3d_gemm_x3_applicable(int64_t M, int N, int K) { return (K % 64 == 0 && N % 8 == 0 && M > 0) ? 1 : 0; }

int spa3d_gemm_x3(const void* A3, int64_t lda, const void* W3, int64_t ldw, const float* bias, const void* residual, int64_t ldr,
                  int r_dtype, float* C, int64_t ldc, int64_t M, int N, int K, void* stream) {
  using namespace spa3d;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_x3: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  if (M == 0) return 0;
  SPA3D_REQUIRE(gemm_tcgen05_x3_applicable(A3, lda, W3, ldw, M, N, K), "gemm_x3: needs K %% 64 == 0, N %% 8 == 0 and 16-byte aligned rows");
  stat_add(ST_GEMM_X3);
  return gemm_tcgen05_x3(A3, lda, W3, ldw, bias, residual, ldr, r_dtype, C, ldc, M, N, K, (cudaStream_t)stream);
}

int spa3d_gemm_rmsnorm(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype, void* C,
                       int64_t ldc, int c_dtype, int64_t M, int N, int K, int Dh, int q_cols,
                       int k_cols, const float* scale_q, const float* scale_k, float q_mul,
                       float* rstd_out, int impl, void* stream) {
  using namespace spa3d;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0 && Dh > 0, "gemm_rmsnorm: bad shape");
  SPA3D_REQUIRE(q_cols % Dh == 0 && k_cols % Dh == 0 && q_cols + k_cols <= N, "gemm_rmsnorm: bad head split");
  if (M == 0) return 0;
  bool tc_ok = (a_dtype == SPA3D_BF16) && gemm_tcgen05_applicable(A, lda, Wt, ldw, M, N, K) &&
               gemm_tcgen05_rms_applicable(N, Dh, q_cols, k_cols, c_dtype);
  if (impl == SPA3D_GEMM_TCGEN05) SPA3D_REQUIRE(tc_ok, "gemm_rmsnorm: tcgen05 path not applicable");
  if (impl != SPA3D_GEMM_SIMT && tc_ok) {
    RmsEpilogue rms{Dh, q_cols, k_cols, scale_q, scale_k, q_mul, rstd_out};
    stat_add(ST_GEMM_TCGEN05);
    return gemm_tcgen05(A, lda, Wt, ldw, nullptr, 0, nullptr, 0, 0, C, ldc, c_dtype, M, N, K, &rms, st);
  }
  // unfused: contraction, then the in-place normalisation pass over the q and k column blocks
  int rc = spa3d_gemm(A, lda, Wt, ldw, a_dtype, nullptr, 0, nullptr, 0, 0, C, ldc, c_dtype, M, N, K,
                      impl == SPA3D_GEMM_SIMT ? SPA3D_GEMM_SIMT : SPA3D_GEMM_AUTO, stream);
  if (rc) return rc;
  const int nh = (q_cols + k_cols) / Dh;
  const size_t esz = c_dtype == SPA3D_F32 ? 4 : 2;
  if (q_cols) {
    rc = head_rmsnorm_fwd_impl(C, ldc, c_dtype, scale_q, q_mul, rstd_out, nh, M, q_cols / Dh, Dh, st);
    if (rc) return rc;
  }
  if (k_cols)
    rc = head_rmsnorm_fwd_impl((char*)C + (size_t)q_cols * esz, ldc, c_dtype, scale_k, 1.f,
                               rstd_out ? rstd_out + q_cols / Dh : nullptr, nh, M, k_cols / Dh, Dh, st);
  return rc;
}

int spa3d_gemm_gelu(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
                    const float* bias, void* Z, int64_t ldz, void* H, int64_t ldh, int64_t M, int N, int K,
                    int save_grad, int impl, void* stream) {
  using namespace spa3d;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_gelu: bad shape");
  if (M == 0) return 0;
  bool tc_ok = (a_dtype == SPA3D_BF16) && gemm_tcgen05_applicable(A, lda, Wt, ldw, M, N, K) &&
               (reinterpret_cast<uintptr_t>(Z) & 15) == 0 && ldz % 8 == 0;
  if (impl == SPA3D_GEMM_TCGEN05) SPA3D_REQUIRE(tc_ok, "gemm_gelu: tcgen05 path not applicable");
  if (save_grad) SPA3D_REQUIRE(tc_ok && impl != SPA3D_GEMM_SIMT, "gemm_gelu: the gelu'(z) side output exists on the tcgen05 path only");
  if (impl != SPA3D_GEMM_SIMT && tc_ok) {
    stat_add(ST_GEMM_TCGEN05);
    return gemm_tcgen05(A, lda, Wt, ldw, bias, SPA3D_ACT_GELU_TANH, nullptr, 0, 0, H, ldh, a_dtype, M, N, K, nullptr, st,
                        0, Z, ldz, save_grad ? 2 : 1);
  }
  int rc = spa3d_gemm(A, lda, Wt, ldw, a_dtype, bias, 0, nullptr, 0, 0, Z, ldz, a_dtype, M, N, K,
                      impl == SPA3D_GEMM_SIMT ? SPA3D_GEMM_SIMT : SPA3D_GEMM_AUTO, stream);
  if (rc) return rc;
  return spa3d_gelu_fwd(Z, ldz, a_dtype, H, ldh, a_dtype, M, N, stream);
}

int spa3d_gemm_gelu_bwd(const void* dH_in, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
                        const void* Z, int64_t ldz, void* dZ, int64_t lddz, int64_t M, int N, int K,
                        int z_is_grad, float* dz_colsum, int impl, void* stream) {
  using namespace spa3d;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_gelu_bwd: bad shape");
  if (M == 0) return 0;
  bool tc_ok = (a_dtype == SPA3D_BF16) && gemm_tcgen05_applicable(dH_in, lda, Wt, ldw, M, N, K) &&
               (reinterpret_cast<uintptr_t>(Z) & 15) == 0 && ldz % 8 == 0 &&
               (reinterpret_cast<uintptr_t>(dZ) & 15) == 0 && lddz % 8 == 0;
  if (impl == SPA3D_GEMM_TCGEN05) SPA3D_REQUIRE(tc_ok, "gemm_gelu_bwd: tcgen05 path not applicable");
  if (impl != SPA3D_GEMM_SIMT && tc_ok) {
    stat_add(ST_GEMM_TCGEN05);
    return gemm_tcgen05(dH_in, lda, Wt, ldw, nullptr, 0, Z, ldz, a_dtype, dZ, lddz, a_dtype, M, N, K, nullptr, st,
                        z_is_grad ? 2 : 1, nullptr, 0, 1, dz_colsum);
  }
  SPA3D_REQUIRE(!z_is_grad && !dz_colsum, "gemm_gelu_bwd: the saved-derivative form and the fused column sums exist on the tcgen05 path only");
  int rc = spa3d_gemm(dH_in, lda, Wt, ldw, a_dtype, nullptr, 0, nullptr, 0, 0, dZ, lddz, a_dtype, M, N, K,
                      impl == SPA3D_GEMM_SIMT ? SPA3D_GEMM_SIMT : SPA3D_GEMM_AUTO, stream);
  if (rc) return rc;
  return spa3d_gelu_bwd(Z, ldz, a_dtype, dZ, lddz, a_dtype, dZ, lddz, a_dtype, M, N, stream);
}

int spa3d_gemm_dw(const void* dY, int64_t lddy, const void* X, int64_t ldx, int dtype, float* dW,
                  int64_t lddw, int64_t M, int N, int K, int accumulate, int impl, void* stream) {
  using namespace spa3d;
  cudaStream_t st = (cudaStream_t)stream;
  SPA3D_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_dw: bad shape");
  bool tc_ok = dtype == SPA3D_BF16 && gemm_tcgen05_dw_applicable(dY, lddy, X, ldx, M, N, K, dW, lddw);
  if (impl == SPA3D_GEMM_TCGEN05) SPA3D_REQUIRE(tc_ok || M == 0, "gemm_dw: tcgen05 path not applicable");
  if (!accumulate) {
    cudaError_t e = cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)K * 4, (size_t)N, st);
    SPA3D_REQUIRE(e == cudaSuccess, "gemm_dw: memset: %s", cudaGetErrorString(e));
  }
  if (M == 0) return 0;
  if (impl != SPA3D_GEMM_SIMT && tc_ok) {
    stat_add(ST_GEMM_DW_TCGEN05);
    return gemm_tcgen05_dw(dY, lddy, X, ldx, dW, lddw, M, N, K, st);
  }
  stat_add(ST_GEMM_DW_SIMT);
  if (impl == SPA3D_GEMM_AUTO && dtype == SPA3D_BF16) stat_add(ST_GEMM_DW_BF16_FALLBACK);
  // SIMT: A(n,m) = dY[m*lddy + n], B(m,k) = X[m*ldx + k]
  return gemm_simt(dY, 1, lddy, dtype, X, ldx, 1, dtype, nullptr, 0, nullptr, 0, 0, dW, lddw, SPA3D_F32, N,
                   K, M, 1, st);
}

int spa3d_gemm_strided(const void* A, int64_t sam, int64_t sak, int a_dtype, const void* B,
                       int64_t sbk, int64_t sbn, int b_dtype, void* C, int64_t ldc, int c_dtype,
                       int64_t M, int N, int64_t K, int accumulate, void* stream) {
  using namespace spa3d;
  SPA3D_REQUIRE(!accumulate || c_dtype == SPA3D_F32, "gemm_strided: accumulate needs f32 C");
  if (M == 0 || N == 0) return 0;
  return gemm_simt(A, sam, sak, a_dtype, B, sbk, sbn, b_dtype, nullptr, 0, nullptr, 0, 0, C, ldc,
                   c_dtype, M, N, K, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
