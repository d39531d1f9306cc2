/*
 * spa3d_b200.h - C ABI of lib3dspa_b200.so: the B200 (sm_100a) kernels behind the 3DSPA hot path.
 *
 * The reference (TheProParadox/3dspa_code) is pure JAX/Flax: it has no plugin / FFI layer.  The
 * boundary it offers is the Flax module call `TrackAutoEncoder3D.apply({'params': p}, batch)`
 * (track_autoencoder_3d.py:309-357) whose body XLA lowers to device kernels.  Each entry point
 * below replaces one group of those XLA-emitted ops and is shaped the
 * way an XLA FFI custom call (or a torch custom op) binds a kernel: caller-owned device buffers
 * as plain pointers, integer sizes / leading dimensions, scalar attributes, an explicit
 * cudaStream_t, and an int status (0 = ok).  The library never allocates, frees or synchronises
 * and holds no mutable global state; every call is re-entrant.  See INTEGRATION.md for the
 * XLA-FFI and torch custom-op bindings.
 *
 * Conventions
 *   - all matrices are row-major; `ld*` = elements between consecutive rows.
 *   - dtype codes: SPA3D_F32 = 0, SPA3D_BF16 = 1.
 *   - "tokens" = rows of the flattened [sequences x length, width] activation matrix.
 *   - stream is a cudaStream_t passed as void*.
 *   - on failure the call returns non-zero; spa3d_last_error() gives a thread-local message.
 */
#ifndef SPA3D_B200_H_
#define SPA3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPA3D_F32 0
#define SPA3D_BF16 1

#define SPA3D_ACT_NONE 0
#define SPA3D_ACT_GELU_TANH 1 /* flax nn.gelu(approximate=True), attention.py:106 */

#define SPA3D_SUMSQ_WORKSPACE 1024 /* floats of scratch for spa3d_sumsq (deterministic two-stage sum) */
#define SPA3D_GEMM_AUTO 0   /* tcgen05 when operands are bf16 and shapes allow, else SIMT fp32 */
#define SPA3D_GEMM_SIMT 1   /* fp32 FMA path (the "fp32-accumulate" accurate mode)           */
#define SPA3D_GEMM_TCGEN05 2 /* force the tcgen05/TMEM/TMA kernel; error if not applicable   */

int spa3d_version(void);
const char* spa3d_last_error(void);

/* ---- dispatch counters -----------------------------------------------------------------------
 * Every contraction / attention entry point picks an implementation (tcgen05, SIMT) from the
 * operand dtypes and alignments.  spa3d_stats copies up to n process-wide counters into out and returns
 * how many exist; spa3d_stat_name(i) names counter i.  The "*_bf16_fallback" counters count bf16 calls that
 * were computed on the fp32 SIMT kernels because an operand missed the tensor-core path's alignment rules
 * (correct, but 10-100x slower): a deployment should see zeros there (bench.py asserts it).
 * The counters are the only process-wide mutable state of the library (relaxed atomics). */
int spa3d_stats(int64_t* out, int n);
int spa3d_stats_reset(void);
const char* spa3d_stat_name(int i);

/* ---- scratch sizes ---------------------------------------------------------------------------
 * The library never allocates: scratch is caller-owned (XLA FFI: a ScratchAllocator / extra result
 * buffer; otherwise the host framework's allocator).  Pure functions of the shapes. */
int64_t spa3d_gemm_workspace_bytes(int64_t M, int N, int K);                        /* 0: split-K partials are added atomically into C */
int64_t spa3d_sumsq_workspace_bytes(void);                                           /* spa3d_sumsq `workspace` */
int64_t spa3d_attention_bwd_workspace_bytes(int64_t batch, int heads, int Lq);      /* spa3d_attention_bwd `delta_ws` */
int64_t spa3d_attention_stats_bytes(int64_t batch, int heads, int Lq);              /* spa3d_attention_fwd `lse_out` (max, sum per row) */
int64_t spa3d_layernorm_bwd_workspace_bytes(int num_partials, int d);               /* spa3d_layernorm_bwd `dscale_partial` */
int64_t spa3d_lift_workspace_bytes(int N, int T, int Hp, int Wp);                   /* spa3d_lift_sample_ws `workspace` */

/* ---- K0: feature lifting (inference.py:287-447) ------------------------------------------
 * One pass over N x T track points: bilinear depth sample + pinhole unprojection
 * (lift_2d_to_3d, :287-336), bilinear DINO patch sample (:339-395) and the 256-channel depth
 * feature (:398-447: ch0=d, ch1=d/10, ch2=d_t-d_{t-1}, rest 0).  float32 arithmetic in the
 * reference's operation order (bit-exact with the NumPy code for float32 inputs).
 *   tracks_2d [N,T,2] f32 (pixels)   depth [T,H,W] f32   dino [T,Hp,Wp,D] f32
 *   xyz [N,T,3] f32   dino_out [N,T,D] (f32|bf16)   depth_out [N,T,Cd] (f32|bf16)
 * Any of depth/dino/xyz/dino_out/depth_out may be NULL to skip that output.
 * intrinsics = {fx, fy, cx, cy} as float (host values). */
int spa3d_lift_sample(const float* tracks_2d, const float* depth, const float* dino,
                      float* xyz, void* dino_out, void* depth_out, int out_dtype,
                      int N, int T, int H, int W, int Hp, int Wp, int D, int Cd,
                      int video_H, int video_W, const float* intrinsics, void* stream);
/* The same call with caller-owned scratch (spa3d_lift_workspace_bytes; 32-byte point records sorted by cell per frame + cell offsets): the
 * points of every frame are first binned by DINO patch cell, and the four corner rows of a cell are then read ONCE for
 * all of its points instead of once per point (bilinear sampling re-reads 4 bytes per byte written otherwise).  Used when
 * dino / dino_out are given and D % 128 == 0; any other call (or workspace == NULL) takes the per-point kernel of
 * spa3d_lift_sample.  Bit-identical results either way. */
int spa3d_lift_sample_ws(const float* tracks_2d, const float* depth, const float* dino,
                         float* xyz, void* dino_out, void* depth_out, int out_dtype,
                         int N, int T, int H, int W, int Hp, int Wp, int D, int Cd,
                         int video_H, int video_W, const float* intrinsics,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a1: SinusoidalEmbedding (track_autoencoder.py:18-38) --------------------------------
 * out[r, c*2F + f] = sin(x[r,c]/scale * 2^(f/3)),  out[r, c*2F + F + f] = sin(... + pi/2).
 * x [rows, C] f32 (ldx).  If append_time > 0 an extra coordinate (r % append_time)/append_time
 * is appended (the fr_id of track_autoencoder_3d.py:126-131).  If tail_zero != 0 one more
 * coordinate equal to 0 is appended (the query_frame // 150.0 feature, :268-269, defect D6).
 * exact != 0: correctly-rounded float32 sine (double-precision evaluation).
 * out [*, ldo] (f32|bf16), written at column offset 0.  out_row_group > 0 leaves room for one
 * extra leading row per group of that many input rows: output row = r + r/out_row_group + 1
 * (the read-out token slot of track_autoencoder_3d.py:161-165). */
int spa3d_fourier_features(const float* x, int64_t ldx, void* out, int64_t ldo, int out_dtype,
                           int64_t rows, int C, int num_freq, float scale_factor,
                           int append_time, int tail_zero, int exact, int out_row_group,
                           void* stream);

/* ---- K1 fused: track embedding without staging the concatenated features -----------------------
 * (track_autoencoder_3d.py:123-149): for every (track, frame) row r of tracks [rows,3] (f32),
 *   out[r + r/T + 1, 0:W] = [Fourier(x,y,z,t/T) | dino[r] | depth[r]] . Wt^T + bias
 * Wt [W, 256+dino_dim+depth_dim] bf16 (the three Flax kernels stacked along K, transposed), bias [W]
 * f32 (sum of the three biases), dino / depth f32 [rows, dim] or NULL with dim 0, out f32 with the
 * read-out slot (row 0 of every T+1 rows) left untouched.  a_cat (bf16 [rows + rows/T, 256+dino+depth], lda) may
 * be NULL; otherwise it receives the concatenated bf16 features at the same remapped rows (training keeps
 * them as the operand of the embedding's weight gradient).  bf16 tensor-core path (sin.approx
 * Fourier features); spa3d_embed_fused_applicable tells whether the widths are supported. */
int spa3d_embed_fused_applicable(int W, int K_total, int dino_dim, int depth_dim, int coords);
int spa3d_embed_fused(const float* tracks, const float* dino, const float* depth, const void* Wt,
                      int64_t ldw, const float* bias, float* out, int64_t ldo, void* a_cat, int64_t lda,
                      int64_t rows, int T, int dino_dim, int depth_dim, int W, int num_freq,
                      float track_scale_factor, void* stream);

/* K1, "project then sample" form (replaces sample_dino_features_for_tracks + sample_depth_features_for_tracks
 * + embed_track_pos_visible, inference.py:339-447,543-557 + track_autoencoder_3d.py:123-149, for the inference pipeline):
 * the DINO projection is linear, so the caller projects the PATCH MAP once per clip (proj = bf16(dino_map) . W_dino^T,
 * [T*Hp*Wp, W] bf16, through spa3d_gemm) and every (track, frame) row r adds the bilinear blend of four projected patch
 * rows; the [N,T,768] / [N,T,256] per-track features never exist.
 *   out[r + r/T + 1] = Fourier(xyz[r], t/T) . Wt[:, 0:256]^T + bias + sum_k w_k(r) proj[patch_k(r)] + dfeat[r,0:3] . wdep
 * xyz [rows,3] f32 (spa3d_lift_sample), tracks_2d [rows,2] f32 pixels, dfeat [rows,4] f32 = (d, d/10, d_t - d_{t-1}, 0)
 * (spa3d_lift_sample with Cd = 4) or NULL, wdep [3,W] f32 = rows 0..2 of the depth projection kernel or NULL, Wt the
 * stacked embedding kernel (only its first 256 columns are read; ldw in elements), bias [W] f32 (sum of the biases in play). */
int spa3d_embed_sampled(const float* xyz, const float* tracks_2d, const float* dfeat, const void* proj, const void* Wt,
                        int64_t ldw, const float* wdep, const float* bias, float* out, int64_t ldo, int64_t rows, int T,
                        int Hp, int Wp, int video_H, int video_W, int W, int num_freq, float track_scale_factor,
                        void* stream);

/* Row-wise dtype conversion / strided copy: dst[r', 0:cols] = (dst_dtype) src[r, 0:cols],
 * r' = r (+ r/out_row_group + 1 when out_row_group > 0). */
int spa3d_convert(const void* src, int64_t lds, int src_dtype, void* dst, int64_t ldd,
                  int dst_dtype, int64_t rows, int cols, int out_row_group, void* stream);

/* dst[i*row_stride, 0:cols] = vec[0:cols] for i < rows  (writes the learned read-out token,
 * ParamStateInit, track_autoencoder.py:41-53, into slot 0 of every sequence). */
int spa3d_set_rows(void* dst, int64_t ld, int dst_dtype, int64_t row_stride, const float* vec,
                   int64_t rows, int cols, void* stream);

/* ---- K2/K4/K5/K9/K11: dense contractions ---------------------------------------------------
 * C = epilogue(A[M,K] . W + bias) (+ residual), the Flax Dense / DenseGeneral of
 * attention.py:106-107,154-183 and track_autoencoder_3d.py:73-115.
 *   A      [M,K]  lda, a_dtype
 *   Wt     [N,K]  ldw, a_dtype  - the Flax kernel [K,N] TRANSPOSED (K contiguous)
 *   bias   [N] f32 or NULL;  act applied after bias;  residual [M,N] ldr, r_dtype or NULL,
 *          added after the activation;  C [M,N] ldc, c_dtype.
 * impl selects the kernel (SPA3D_GEMM_*).  The tcgen05 kernel needs bf16 A/Wt, K % 8 == 0 and
 * 16-byte aligned rows. */
int spa3d_gemm(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
               const float* bias, int act, const void* residual, int64_t ldr, int r_dtype,
               void* C, int64_t ldc, int c_dtype, int64_t M, int N, int K, int impl,
               void* stream);

/* Tensor-core form of the accurate ("fp32-accumulate", 1e-4) mode: "bf16 x 3".  spa3d_split3 writes an fp32 matrix as three bf16
 * terms side by side, dst[r, p*K + c] = part p of src[r, c] with x = hi + mid + lo to 24 bits (dst bf16 [rows, 3K]).
 * spa3d_gemm_x3 contracts two split operands: C[M,N] (f32) = A . W^T + bias (+ residual), six tcgen05 products per K block
 * (hi.hi, hi.mid, mid.hi, hi.lo, mid.mid, lo.hi; every bf16 x bf16 product is exact in fp32) accumulated in ONE fp32 TMEM
 * accumulator, smallest terms first.  Needs K % 64 == 0, N % 8 == 0 (spa3d_gemm_x3_applicable); other shapes use the SIMT path. */
int spa3d_split3(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int K, void* stream);
int spa3d_gemm_x3_applicable(int64_t M, int N, int K);
int spa3d_gemm_x3(const void* A3, int64_t lda, const void* W3, int64_t ldw, const float* bias, const void* residual,
                  int64_t ldr, int r_dtype, float* C, int64_t ldc, int64_t M, int N, int K, void* stream);

/* MLP_in with its activation, training form (attention.py:106): Z = A . Wt^T + bias (saved for the
 * backward pass) and H = gelu_tanh(Z), both [M,N] in a_dtype, from one pass over the accumulator.
 * save_grad = 1 (tcgen05 path only): Z receives gelu_tanh'(A . Wt^T + bias) instead - the derivative shares tanh(u) with
 * the activation, and the backward epilogue becomes one multiply. */
int spa3d_gemm_gelu(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
                    const float* bias, void* Z, int64_t ldz, void* H, int64_t ldh, int64_t M, int N,
                    int K, int save_grad, int impl, void* stream);
/* Backward through MLP_out and the activation: dZ = (dY . Wt^T) * gelu_tanh'(Z); dY [M,K], Wt [N,K]
 * (MLP_out's kernel [N,K] as stored by Flax), Z, dZ [M,N], all in a_dtype.  z_is_grad = 1: Z already holds
 * gelu_tanh'(z) (spa3d_gemm_gelu with save_grad = 1) and dZ = (dY . Wt^T) * Z.  dz_colsum ([N] f32, may be NULL; tcgen05 path
 * only) += the column sums of dZ, i.e. the gradient of MLP_in's bias, reduced in the epilogue (no pass over dZ). */
int spa3d_gemm_gelu_bwd(const void* dY, int64_t lddy, const void* Wt, int64_t ldw, int a_dtype,
                        const void* Z, int64_t ldz, void* dZ, int64_t lddz, int64_t M, int N, int K,
                        int z_is_grad, float* dz_colsum, int impl, void* stream);

/* Fused MLP sub-block, inference form (attention.py:102-108): out = residual + gelu_tanh(A . W1 + b1) . W2 + b2 in ONE kernel - the
 * hidden activation [M, Hd] never reaches HBM (unfused it is written by MLP_in and re-read by MLP_out: 2 x M x Hd x 2 bytes).
 * A [M, D] bf16 = LayerNorm(a) (lda); W1t [Hd, D] bf16 and W2t [D, Hd] bf16 are the Flax kernels transposed (K contiguous); b1 [Hd],
 * b2 [D] f32; residual, out [M, D] f32.  D = 384 and Hd % 128 == 0 (spa3d_mlp_fused_applicable): 128 TMEM columns for a hidden
 * chunk plus D columns for the output tile are all 512.  Measured no faster than the two GEMMs on the per-track shape (its weight
 * ring is one L2 latency deep; see the kernel header), so the model does not use it by default. */
int spa3d_mlp_fused_applicable(int D, int Hd);
int spa3d_mlp_fused(const void* A, int64_t lda, const void* W1t, int64_t ldw1, const float* b1, const void* W2t, int64_t ldw2,
                    const float* b2, const float* residual, int64_t ldr, float* out, int64_t ldo, int64_t M, int D, int Hd,
                    void* stream);

/* Weight gradient of a Dense layer (backward of attention.py:106-107,154-183 and
 * track_autoencoder_3d.py:73-115 under jax.value_and_grad, train.py:161-162):
 *   dW[N,K] (+)= dY[M,N]^T . X[M,K]     dY, X in `dtype` (row-major, lddy / ldx), dW f32 (lddw),
 * in the packed [out,in] layout of Wt.  The reduction runs over the M tokens.  bf16 operands take
 * the tcgen05 kernel (token-major = MN-major operands, split-K with vector atomics into dW). */
int spa3d_gemm_dw(const void* dY, int64_t lddy, const void* X, int64_t ldx, int dtype, float* dW,
                  int64_t lddw, int64_t M, int N, int K, int accumulate, int impl, void* stream);

/* General strided fp32-accumulate GEMM used by the backward pass:
 *   C[m,n] (+)= sum_k A(m,k) * B(k,n),  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
 * accumulate != 0 adds into C (C must be f32 then). */
int spa3d_gemm_strided(const void* A, int64_t sam, int64_t sak, int a_dtype,
                       const void* B, int64_t sbk, int64_t sbn, int b_dtype,
                       void* C, int64_t ldc, int c_dtype, int64_t M, int N, int64_t K,
                       int accumulate, void* stream);

/* ---- LayerNorm(use_bias=False) (attention.py:49,76,103; flax eps 1e-6, fast variance) ------
 * y[r,:] = (x[r,:]-mean)*rsqrt(var+1e-6)*scale.  x [rows,d] ldx x_dtype; y ldy y_dtype.
 * mean_out / rstd_out ([rows] f32) may be NULL.  row_stride_tokens > 1 normalises only rows
 * r*row_stride_tokens (used to read token 0 of every sequence, track_autoencoder_3d.py:187,286). */
int spa3d_layernorm_fwd(const void* x, int64_t ldx, int x_dtype, const float* scale,
                        void* y, int64_t ldy, int y_dtype, float* mean_out, float* rstd_out,
                        int64_t rows, int d, void* stream);
/* Backward: dx (+)= rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*scale; dscale_partial
 * [num_partials, d] receives per-block partial sums of dy*xhat.  dx_lowp (bf16 [rows,d], ldl) may be
 * NULL; otherwise it receives a bf16 copy of the final dx (the operand of the next backward GEMMs).
 * accumulate_dscale != 0: dscale_partial is the [d] gradient of the scale itself and every block adds
 * its partial sum into it atomically (no separate reduction pass).  dx_colsum ([d] f32, may be NULL; fp32 fast path with
 * d <= 512 or d in {1024, 1280, 1536} only) += the sum over rows of the FINAL dx: that is the bias gradient of the Dense layer that produced x
 * (attention.py:107,182), so no separate pass over dx is needed for it. */
int spa3d_layernorm_bwd(const void* x, int64_t ldx, int x_dtype, const float* scale,
                        const float* mean, const float* rstd, const void* dy, int64_t lddy,
                        int dy_dtype, void* dx, int64_t lddx, int dx_dtype, int dx_accumulate,
                        void* dx_lowp, int64_t ldl, float* dscale_partial, int num_partials,
                        int accumulate_dscale, float* dx_colsum, int64_t rows, int d, void* stream);

/* ---- per-head RMSNorm of q and k (attention.py:166-167) + q/sqrt(Dh) (flax attention) ------
 * In place on a packed projection buffer: for every row and head h < heads,
 *   buf[r, h*Dh : (h+1)*Dh] = x * rsqrt(mean(x^2)+1e-6) * scale[:] * out_mul.
 * rstd_out[r*rstd_ld + h] f32 may be NULL (saved for backward). */
int spa3d_head_rmsnorm_fwd(void* buf, int64_t ld, int dtype, const float* scale, float out_mul,
                           float* rstd_out, int64_t rstd_ld, int64_t rows, int heads, int Dh,
                           void* stream);
int spa3d_head_rmsnorm_bwd(const void* y, int64_t ldy, int y_dtype, const float* scale,
                           float out_mul, const float* rstd, int64_t rstd_ld, void* dy_inout,
                           int64_t ldd, int d_dtype, float* dscale_partial, int num_partials, int accumulate_dscale,
                           int64_t rows, int heads, int Dh, void* stream);

/* QKV projection with the per-head RMSNorm fused into the GEMM epilogue (attention.py:154-173):
 *   C = A . Wt^T; columns [0,q_cols) normalised per head with scale_q and multiplied by q_mul
 *   (= 1/sqrt(Dh), flax dot_product_attention), columns [q_cols, q_cols+k_cols) with scale_k, the
 *   remaining (value) columns stored as is.  rstd_out [M, (q_cols+k_cols)/Dh] f32 or NULL.
 *   Self-attention: N=3A, q_cols=k_cols=A.  Cross-attention: q projection (q_cols=A, k_cols=0) and
 *   key/value projection (q_cols=0, k_cols=A, N=2A). */
int spa3d_gemm_rmsnorm(const void* A, int64_t lda, const void* Wt, int64_t ldw, int a_dtype,
                       void* C, int64_t ldc, int c_dtype, int64_t M, int N, int K, int Dh,
                       int q_cols, int k_cols, const float* scale_q, const float* scale_k,
                       float q_mul, float* rstd_out, int impl, void* stream);

/* ---- K3/K6: softmax(q k^T [+ key mask]) v  (flax nn.dot_product_attention, attention.py:175)
 * Batched over `batch` independent sequences and `heads` heads.
 *   q [batch*Lq, *] ldq : head h at columns [h*Dh, (h+1)*Dh)   (already RMS-normed and /sqrt(Dh))
 *   k, v [batch*Lk, *] ldk/ldv, same column convention;  o [batch*Lq, heads*Dh] ldo.
 *   key_mask [batch, Lk] uint8 or NULL: 0 => logit replaced by -FLT_MAX (finfo.min semantics,
 *   an all-masked row yields uniform weights).
 *   lse_out [batch, heads, Lq, 2] f32 or NULL: softmax statistics saved for the backward,
 *   (row max, 1/row sum) - kept separate so all-masked rows (max = -FLT_MAX) stay exact.
 * dtype applies to q,k,v,o.  Lq==Lk<=256 with a mask is the per-track temporal self-attention
 * (track_autoencoder_3d.py:182-184); Lq=128, Lk=N is the latents<-tracks cross-attention (:201). */
int spa3d_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                        float* lse_out, int64_t batch, int heads, int Lq, int Lk, int Dh,
                        void* stream);
int spa3d_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                        void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                        int dtype, const uint8_t* key_mask, const float* lse, float* delta_ws,
                        int64_t batch, int heads, int Lq, int Lk, int Dh, void* stream);
/* delta_ws: caller-provided scratch, [batch*heads*Lq] f32. */

/* K6 on the tensor cores: the latents<-tracks cross-attention (track_autoencoder_3d.py:200-201; attention.py:92-100), Lq <= 128
 * queries over Lk keys (Lq != Lk), bf16.  The keys are split into 128-key chunks, one tcgen05 work item per (sequence, head,
 * chunk) so that every SM takes part; the chunks' (max, sum, O) states are merged flash-style by a second small kernel, and
 * in the backward the chunks' partial dQ tiles are summed the same way.  `workspace` is caller-owned scratch of
 * spa3d_attention_cross_workspace_bytes bytes (NULL, or shapes the kernel does not cover, fall back to
 * spa3d_attention_fwd / _bwd).  Arguments otherwise as spa3d_attention_fwd / spa3d_attention_bwd. */
int spa3d_attention_cross_applicable(int dtype, int Lq, int Lk, int Dh);
int64_t spa3d_attention_cross_workspace_bytes(int64_t batch, int heads, int Lq, int Lk, int Dh);
int spa3d_attention_cross_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                              int64_t ldo, int dtype, const uint8_t* key_mask, float* lse_out, float* workspace,
                              int64_t batch, int heads, int Lq, int Lk, int Dh, void* stream);
int spa3d_attention_cross_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              const void* o, int64_t ldo, const void* d_o, int64_t lddo, void* dq, int64_t lddq, void* dk,
                              int64_t lddk, void* dv, int64_t lddv, int dtype, const uint8_t* key_mask, const float* lse,
                              float* delta_ws, float* workspace, int64_t batch, int heads, int Lq, int Lk, int Dh,
                              void* stream);

/* ---- R1 key mask (track_autoencoder_3d.py:167-184, repaired) -------------------------------
 * mask[b,n,0] = 1 (if has_readout); mask[b,n,j] = visible[b,n,t]!=0 && t < boundary[b]. */
int spa3d_build_key_mask(const float* visible, const int32_t* boundary_frame, uint8_t* mask,
                         int B, int N, int T, int has_readout, void* stream);

/* ---- TRAJAN pooling (track_autoencoder.py:230-232) -----------------------------------------
 * out[s,:] = sum_t tok[s,t,:]*(vis[s,t]!=0) / max(1, sum_t (vis[s,t]!=0)).  tok [S*T, W]. */
int spa3d_masked_mean_fwd(const void* tok, int64_t ldt, int tok_dtype, const float* visible,
                          void* out, int64_t ldo, int out_dtype, int64_t S, int T, int W,
                          void* stream);

/* elementwise tanh-GELU: y = gelu(x) (used by the training path, which keeps the pre-activation) */
int spa3d_gelu_fwd(const void* x, int64_t ldx, int x_dtype, void* y, int64_t ldy, int y_dtype,
                   int64_t rows, int cols, void* stream);

/* ---- K7: clip / quantise / noise with straight-through gradient (:251-260) -----------------
 * y = clip(x,-1,1); if discretize: y = round_half_even(y*128)/128 + noise/128 - 1/256.
 * pass_mask (uint8, may be NULL) records |x|<=1 for the straight-through backward. */
int spa3d_quantize_fwd(const float* x, const float* noise, float* y, uint8_t* pass_mask,
                       int64_t n, int discretize, void* stream);

/* straight-through backward of the quantiser: dx = dy where |x| <= 1 (pass_mask), else 0. */
int spa3d_quantize_bwd(const float* dy, const uint8_t* pass_mask, float* dx, int64_t n,
                       void* stream);

/* ---- K10: decoder token assembly (track_autoencoder_3d.py:276-284, append_time_feat :235-246)
 * tokens[b,q,0,:]   = query_emb[b,q,:]                      (D = C + 128 channels)
 * tokens[b,q,1+n,:] = [ lat[b,n,0:C] , lat[b,n,5*t:5*t+128] ]   t = query_frame[b,q]
 * (window entries beyond C are 0, as in the reference's eye() einsum).
 * lat [B,L,C] lat_dtype; query_emb [B*Q, D]; tokens [B*Q*(L+1), D] tok_dtype. */
int spa3d_decoder_tokens_fwd(const void* lat, int lat_dtype, const void* query_emb, int qe_dtype,
                             const int32_t* query_frame, void* tokens, int tok_dtype, int B, int Q,
                             int L, int C, void* stream);
/* d_lat[b,n,c] = sum_q ( d_tok[b,q,1+n,c] + window contributions ); d_qe = d_tok[b,q,0,:]. */
int spa3d_decoder_tokens_bwd(const void* d_tokens, int tok_dtype, const int32_t* query_frame,
                             float* d_lat, float* d_query_emb, int B, int Q, int L, int C,
                             void* stream);

/* ---- K11: output split + loss (track_autoencoder_3d.py:289-301, train.py:96-129) -----------
 * head_out [rows, 4*T] f32 (x|y|z|vis blocks) -> tracks [rows,T,3], visible_logits [rows,T]. */
int spa3d_split_outputs(const float* head_out, float* tracks, float* visible_logits,
                        int64_t rows, int T, int coords, float* certain_logits, void* stream);
/* evaluate_tapvid3d.py:39-59 (convert_predictions_to_tapvid3d_format) for one clip: tracks [Q,T,coords], visible_logits [Q,T]
 * -> out_tracks [T,Q,coords], out_occluded [T,Q] (1 where logit <= 0).  With target_tracks [Q,T,coords] and out_score [T,Q] non-NULL it
 * also writes the per-point reconstruction error |pred - target|_2 (the coords_score array visualize.py:186 consumes). */
int spa3d_to_tapvid3d(const float* tracks, const float* visible_logits, const float* target_tracks, float* out_tracks,
                      uint8_t* out_occluded, float* out_score, int64_t Q, int T, int coords, void* stream);
/* sums[0] += sum |pred-tgt|*vis, sums[1] += sum BCE(logit,vis), sums[2] += sum vis.
 * (sums must be zeroed by the caller; the three scalars of compute_loss_3d follow on host or
 * after an all-reduce of sums across ranks.) */
int spa3d_loss_fwd(const float* head_out, const float* target_tracks, const float* target_vis,
                   float* sums, int64_t rows, int T, void* stream);
/* d_head_out for total = l1_w*pos + bce_w*vis with the GLOBAL normaliser inv_denom. */
int spa3d_loss_bwd(const float* head_out, const float* target_tracks, const float* target_vis,
                   float* d_head_out, float l1_w, float bce_w, float inv_denom, int64_t rows,
                   int T, void* stream);

/* ---- elementwise helpers for the backward pass ---------------------------------------------
 * gelu_bwd: dx = dy * gelu'(pre)   (pre = pre-activation);  colsum: out[n] (+)= sum_r x[r,n]. */
int spa3d_gelu_bwd(const void* pre, int64_t ldp, int p_dtype, const void* dy, int64_t lddy,
                   int dy_dtype, void* dx, int64_t lddx, int dx_dtype, int64_t rows, int cols,
                   void* stream);
int spa3d_colsum(const void* x, int64_t ldx, int dtype, float* out, int accumulate,
                 int64_t rows, int cols, void* stream);
/* Compute-dtype shadows of one fp32 master matrix src [rows, cols] in one pass (after an optimiser step):
 * dst [rows, cols] = T(src) - the [out,in] operand of the forward GEMMs - and dst_t [cols, rows] = T(src)^T - the
 * [in,out] operand of dX = dY . W.  Either destination may be NULL.  dtype = SPA3D_BF16 | SPA3D_F32. */
int spa3d_shadow_weights(const float* src, int64_t lds, void* dst, int64_t ldd, void* dst_t, int64_t ldt, int dtype,
                         int rows, int cols, void* stream);
/* bytes of zeros (gradient buffers before the first accumulation). */
int spa3d_fill_zero(void* p, int64_t bytes, void* stream);
/* y = a + b elementwise over n floats (y may alias a). */
int spa3d_axpy(float* y, const float* x, float alpha, int64_t n, void* stream);

/* ---- a16: optimiser (train.py:41-57,239-243; optax adamw + clip_by_global_norm) ------------
 * sumsq[0] += sum g^2 ;  then adamw with clip factor computed on device from sumsq. */
int spa3d_sumsq(const float* g, int64_t n, float* sumsq, float* workspace, void* stream);
int spa3d_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                     float clip_norm, float lr, float b1, float b2, float eps, float wd,
                     int step, void* stream);

/* ---- host-side staging for the end-to-end path (csrc/host_pack.cc; no device work) ---------
 * dst[i] = bf16(src[i]) for n float32 values in HOST memory, round to nearest even (bit-identical to the
 * device-side conversion of the embedding producers), on `threads` host threads, streaming stores.
 * Replaces nothing in the reference: it halves the PCIe bytes of the float32 DINOv2 maps / features the
 * reference's pipeline hands over on the host (inference.py:523-590) before model.apply. */
int spa3d_host_pack_bf16(const float* src, void* dst, int64_t n, int threads);

#ifdef __cplusplus
}
#endif
#endif /* SPA3D_B200_H_ */
