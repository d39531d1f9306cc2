// C-ABI entry points for the attention cores; dispatch between the tcgen05 kernels (bf16: short self-attention sequences
// held in one tile, and the chunked latents<-tracks cross-attention), the one-query kernel of the pruned last layers, and the
// fp32 SIMT kernels (the accurate mode; also any bf16 shape the tensor-core kernels do not cover, counted as a fallback).
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
bool attention_fwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                 const void* q, const void* k, const void* v, const void* o);
int attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int L, int Dh,
                     cudaStream_t st);
bool attention_bwd_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t lddo,
                                 int64_t lddq, int64_t lddk, int64_t lddv);
int attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, const float* stats, const float* delta, int64_t batch, int heads, int L,
                     int Dh, cudaStream_t st);
int attention_delta(const void* o, int64_t ldo, const void* d_o, int64_t lddo, int dtype, float* delta_ws,
                    int64_t batch, int heads, int Lq, int Dh, cudaStream_t st);
bool attention_q1_applicable(int dtype, int Lq, int Lk, int Dh);
int attention_q1_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, const uint8_t* key_mask, float* stats, int64_t batch, int heads, int Lk, int Dh,
                     cudaStream_t st);
int attention_q1_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                     int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                     const uint8_t* key_mask, int64_t batch, int heads, int Lk, int Dh, cudaStream_t st);
bool attention_cross_tc_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo_or_grads);
int64_t attention_cross_workspace_floats(int64_t batch, int heads, int Lk, int Dh);
int attention_cross_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                           const uint8_t* key_mask, float* stats, float* workspace, int64_t batch, int heads, int Lq, int Lk, int Dh,
                           cudaStream_t st);
int attention_cross_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o, int64_t lddo,
                           void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, const uint8_t* key_mask, const float* stats,
                           const float* delta, float* workspace, int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st);
}  // namespace spa3d

extern "C" {

// ---- latents<-tracks cross-attention on tcgen05 (track_autoencoder_3d.py:200-201; attention.py:92-100) ------------------
int spa3d_attention_cross_applicable(int dtype, int Lq, int Lk, int Dh) {
  return spa3d::attention_cross_tc_applicable(dtype, Lq, Lk, Dh, 8, 8, 8, 8) ? 1 : 0;
}

int64_t spa3d_attention_cross_workspace_bytes(int64_t batch, int heads, int Lq, int Lk, int Dh) {
  (void)Lq;
  return spa3d::attention_cross_workspace_floats(batch, heads, Lk, Dh) * 4;
}

int spa3d_attention_cross_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                              int dtype, const uint8_t* key_mask, float* lse_out, float* workspace, int64_t batch, int heads, int Lq,
                              int Lk, int Dh, void* stream) {
  using namespace spa3d;
  if (batch == 0 || Lq == 0) return 0;
  SPA3D_REQUIRE(Lk > 0, "attention_cross_fwd: Lk must be > 0");
  if (workspace != nullptr && attention_cross_tc_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, ldo)) {
    stat_add(ST_ATTN_CROSS_TCGEN05);
    return attention_cross_fwd_tc(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, workspace, batch, heads, Lq, Lk, Dh, (cudaStream_t)stream);
  }
  return spa3d_attention_fwd(q, ldq, k, ldk, v, ldv, o, ldo, dtype, key_mask, lse_out, batch, heads, Lq, Lk, Dh, stream);
}

int spa3d_attention_cross_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                              const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                              int dtype, const uint8_t* key_mask, const float* lse, float* delta_ws, float* workspace, int64_t batch,
                              int heads, int Lq, int Lk, int Dh, void* stream) {
  using namespace spa3d;
  if (batch == 0 || Lq == 0) return 0;
  SPA3D_REQUIRE(lse != nullptr && delta_ws != nullptr, "attention_cross_bwd: lse/delta_ws required");
  if (workspace != nullptr && attention_cross_tc_applicable(dtype, Lq, Lk, Dh, ldq, ldk, ldv, lddo) && lddq % 8 == 0 && lddk % 8 == 0 &&
      lddv % 8 == 0) {
    stat_add(ST_ATTN_CROSS_TCGEN05);
    int rc = attention_delta(o, ldo, d_o, lddo, dtype, delta_ws, batch, heads, Lq, Dh, (cudaStream_t)stream);
    if (rc) return rc;
    return attention_cross_bwd_tc(q, ldq, k, ldk, v, ldv, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, lse, delta_ws, workspace, batch,
                                  heads, Lq, Lk, Dh, (cudaStream_t)stream);
  }
  return spa3d_attention_bwd(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, dtype, key_mask, lse, delta_ws, batch,
                             heads, Lq, Lk, Dh, stream);
}

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
    return attention_bwd_tc(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv, key_mask, lse, delta_ws,
                            batch, heads, Lq, Dh, (cudaStream_t)stream);
  }
  stat_add(ST_ATTN_SIMT);
  if (dtype == SPA3D_BF16) stat_add(ST_ATTN_BF16_FALLBACK);
  return attention_bwd_simt(q, ldq, k, ldk, v, ldv, o, ldo, d_o, lddo, dq, lddq, dk, lddk, dv, lddv,
                            dtype, key_mask, lse, delta_ws, batch, heads, Lq, Lk, Dh,
                            (cudaStream_t)stream);
}

}  // extern "C"
