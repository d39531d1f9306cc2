// C-ABI entry points for the attention cores; dispatch between the tensor-core kernel
// (bf16, short sequences held in one tile) and the fp32 SIMT kernel.
#include "common.cuh"

namespace spa3d {
int attention_fwd_simt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                       float* lse_out, int64_t batch, int heads, int Lq, int Lk, int Dh,
                       cudaStream_t st);
int attention_bwd_simt(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                       void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                       int dtype, const uint8_t* key_mask, const float* lse, float* delta_ws,
                       int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st);
bool attention_fwd_mma_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk,
                                  int64_t ldv, int64_t ldo);
int attention_fwd_mma(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse_out,
                      int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st);
bool attention_fwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                 const void* q, const void* k, const void* v, const void* o);
int attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int L, int Dh,
                     cudaStream_t st);
bool attention_bwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t lddo,
                                 int64_t lddq, int64_t lddk, int64_t lddv);
int attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                     int Dh, cudaStream_t st);
int attention_delta(const void* o, int64_t ldo, const void* d_o, int64_t lddo, int dtype, float* delta_ws,
                    int64_t batch, int heads, int Lq, int Dh, cudaStream_t st);
bool attention_bwd_mma_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv,
                                  int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv);
int attention_bwd_mma(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk,
                      void* dv, int64_t lddv, const uint8_t* key_mask, const float* stats,
                      const float* delta, int64_t batch, int heads, int Lq, int Lk, int Dh,
                      cudaStream_t st);
bool attention_q1_applicable(int dtype, int Lq, int Lk, int Dh);
int attention_q1_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int Lk, int Dh,
                     cudaStream_t st);
int attention_q1_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, int64_t batch, int heads, int Lk, int Dh, cudaStream_t st);
}  // namespace spa3d

extern "C" {

int spa3d_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                        float* lse_out, int64_t batch, int heads, int Lq, int Lk, int Dh,
                        void* stream) {
  using namespace spa3d;
  if (batch == 0 || Lq == 0) return 0;
  SPA3D_REQUIRE(Lk > 0, "attention: Lk must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (attention_q1_applicable(dtype, Lq, Lk, Dh)) {   // one query per sequence: the pruned last layers
    stat_add(ST_ATTN_Q1);
    return attention_q1_fwd(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lk, Dh, st);
  }
  if (attention_fwd_tc_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, ldo, q, k, v, o)) {
    stat_add(ST_ATTN_TCGEN05);
    return attention_fwd_tc(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Dh, st);
  }
  if (attention_fwd_mma_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, ldo)) {
    stat_add(ST_ATTN_MMA);
    return attention_fwd_mma(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Lk,
                             Dh, st);
  }
  stat_add(ST_ATTN_SIMT);
  if (dtype == SPA3D_BF16) stat_add(ST_ATTN_BF16_FALLBACK);
  return attention_fwd_simt(q, ldq, k, ldk, v, ldv, o, ldo, dtype, key_mask, lse_out, batch, heads,
                            Lq, Lk, Dh, st);
}

int spa3d_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                        void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                        int dtype, const uint8_t* key_mask, const float* lse, float* delta_ws,
                        int64_t batch, int heads, int Lq, int Lk, int Dh, void* stream) {
  using namespace spa3d;
  if (batch == 0 || Lq == 0) return 0;
  SPA3D_REQUIRE(lse != nullptr && delta_ws != nullptr, "attention_bwd: lse/delta_ws required");
  if (attention_q1_applicable(dtype, Lq, Lk, Dh)) {   // recomputes the softmax: neither lse nor delta is read
    stat_add(ST_ATTN_Q1);
    return attention_q1_bwd(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, batch, heads, Lk, Dh,
                            (cudaStream_t)stream);
  }
  if (attention_bwd_tc_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, lddo, lddq, lddk, lddv)) {
    stat_add(ST_ATTN_TCGEN05);
    // delta is computed inside the kernel from P and dP
    return attention_bwd_tc(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, lse, delta_ws,
                            batch, heads, Lq, Dh, (cudaStream_t)stream);
  }
  if (attention_bwd_mma_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, lddo, lddq, lddk, lddv)) {
    stat_add(ST_ATTN_MMA);
    int rc = attention_delta(o, ldo, d_o, lddo, dtype, delta_ws, batch, heads, Lq, Dh, (cudaStream_t)stream);
    if (rc) return rc;
    return attention_bwd_mma(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, lse,
                             delta_ws, batch, heads, Lq, Lk, Dh, (cudaStream_t)stream);
  }
  stat_add(ST_ATTN_SIMT);
  if (dtype == SPA3D_BF16) stat_add(ST_ATTN_BF16_FALLBACK);
  return attention_bwd_simt(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv,
                            dtype, key_mask, lse, delta_ws, batch, heads, Lq, Lk, Dh,
                            (cudaStream_t)stream);
}

}  // extern "C"
