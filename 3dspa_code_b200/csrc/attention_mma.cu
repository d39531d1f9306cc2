// Short-sequence attention core on tensor cores (bf16 in, fp32 softmax/accumulate):
//   o = softmax(q k^T [+ key mask]) v   per (sequence, head), whole sequence in one CTA.
//
// Used for the per-track temporal self-attention (L = T+1 = 151, track_autoencoder_3d.py:182-184),
// the decoder read-out attention (L = 129, :285) and the latent self-attentions (L = 128).
// One CTA owns one (sequence, head): Q, K, V ([L, Dh] each) are staged once in shared memory with
// cp.async (rows padded by 16 B so ldmatrix is bank-conflict free), each warp owns 16 query rows
// and walks the keys in tiles of 32 with an online softmax, so the L x L score matrix never
// leaves registers.  S = Q K^T and O = P V run on mma.sync.m16n8k16 (bf16 -> fp32).
//
// Why not tcgen05 here: L = 151/129 rows would pad to 256 (or 192 with M=64 tiles) in a UMMA
// tile, the core is 2-9 % of the layer FLOPs, and at Dh = 96 it is bound by the HBM traffic of
// q/k/v/o, not by the tensor pipe.  The dense projections around it use tcgen05.
//
// Mask semantics (attention.py:175, flax dot_product_attention): masked logits become
// finfo(float32).min, so an all-masked row is uniform; keys beyond Lk do not exist (-inf).
#include <cuda_pipeline.h>

#include "common.cuh"

namespace spa3d {

namespace am {

constexpr int KT = 32;  // keys per online-softmax step

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int DH>
__global__ void __launch_bounds__(512)
attn_fwd_mma_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
                    const bf16* __restrict__ v, int64_t ldv, bf16* __restrict__ o, int64_t ldo,
                    const uint8_t* __restrict__ mask, float* __restrict__ stats, int heads, int Lq,
                    int Lk, int QR, int KC, int q_tiles) {
  // QR = query rows per CTA (16 per warp), KC = keys staged per chunk (multiple of KT).
  // Lk <= KC: whole sequence in one tile (self-attention, L = 151/129/128).  Lk > KC: the keys are
  // streamed chunk by chunk under the same online softmax (latents <- tracks cross-attention,
  // 128 queries x N keys, track_autoencoder_3d.py:200-201).
  constexpr int LDS = DH + 8;    // padded smem row, elements (16 B pad)
  constexpr int CH = DH / 8;     // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + (size_t)QR * LDS;
  bf16* Vs = Ks + (size_t)KC * LDS;
  uint8_t* Ms = reinterpret_cast<uint8_t*>(Vs + (size_t)KC * LDS);  // 1 keep, 0 masked, 2 absent

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int qt = blockIdx.x % q_tiles;
  const int64_t bh = blockIdx.x / q_tiles;
  const int64_t b = bh / heads;
  const int h = (int)(bh % heads);
  const int q0 = qt * QR;
  const bf16* qg = q + (b * Lq + q0) * ldq + (int64_t)h * DH;
  const bf16* kg = k + (b * Lk) * ldk + (int64_t)h * DH;
  const bf16* vg = v + (b * Lk) * ldv + (int64_t)h * DH;

  for (int idx = tid; idx < QR * CH; idx += nthr) {
    int r = idx / CH, c = idx % CH;
    bf16* dst = Qs + r * LDS + c * 8;
    if (q0 + r < Lq) __pipeline_memcpy_async(dst, qg + (int64_t)r * ldq + c * 8, 16);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }

  const int r0 = warp * 16;
  const bool active = q0 + r0 < Lq;
  const int g = lane >> 2, tq = lane & 3;

  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float oacc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;

  // ldmatrix source coordinates (see fragment layouts of mma.m16n8k16)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;       // A (Q), x4
  const int kb_row = (lane & 7) + (lane >> 4) * 8, kb_col = ((lane >> 3) & 1) * 8;     // B from K
  const int vb_row = (lane & 7) + ((lane >> 3) & 1) * 8, vb_col = (lane >> 4) * 8;     // B from V (.trans)
  constexpr float LOG2E = 1.4426950408889634f;

  for (int kc0 = 0; kc0 < Lk; kc0 += KC) {
    if (kc0 > 0) __syncthreads();  // every warp is done with the previous chunk
    const int kn = min(KC, (Lk - kc0 + KT - 1) / KT * KT);  // staged rows of this chunk (padded to KT)
    for (int idx = tid; idx < kn * CH; idx += nthr) {
      int r = idx / CH, c = idx % CH;
      bf16* dk = Ks + r * LDS + c * 8;
      bf16* dv = Vs + r * LDS + c * 8;
      if (kc0 + r < Lk) {
        __pipeline_memcpy_async(dk, kg + (int64_t)(kc0 + r) * ldk + c * 8, 16);
        __pipeline_memcpy_async(dv, vg + (int64_t)(kc0 + r) * ldv + c * 8, 16);
      } else {
        *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
      }
    }
    for (int j = tid; j < kn; j += nthr)
      Ms[j] = kc0 + j < Lk ? (mask == nullptr ? 1 : (mask[b * Lk + kc0 + j] != 0 ? 1 : 0)) : 2;
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    if (!active) continue;

  for (int kt = 0; kt < kn; kt += KT) {
    float s[KT / 8][4];
#pragma unroll
    for (int i = 0; i < KT / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a[4];
      ldsm_x4(a, Qs + (r0 + a_row) * LDS + kk * 16 + a_col);
#pragma unroll
      for (int n2 = 0; n2 < KT / 16; ++n2) {
        uint32_t bb[4];
        ldsm_x4(bb, Ks + (kt + n2 * 16 + kb_row) * LDS + kk * 16 + kb_col);
        mma_bf16(s[n2 * 2], a, bb[0], bb[1]);
        mma_bf16(s[n2 * 2 + 1], a, bb[2], bb[3]);
      }
    }
    // mask + online softmax (rows g and g+8 of this warp's 16)
    float tmax0 = -INFINITY, tmax1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        uint8_t mk = Ms[kt + nt * 8 + tq * 2 + e];
        float lo = mk == 1 ? s[nt][e] : (mk == 0 ? -FLT_MAX : -INFINITY);
        float hi = mk == 1 ? s[nt][2 + e] : (mk == 0 ? -FLT_MAX : -INFINITY);
        s[nt][e] = lo;
        s[nt][2 + e] = hi;
        tmax0 = fmaxf(tmax0, lo);
        tmax1 = fmaxf(tmax1, hi);
      }
    }
    tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 1));
    tmax0 = fmaxf(tmax0, __shfl_xor_sync(0xffffffffu, tmax0, 2));
    tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 1));
    tmax1 = fmaxf(tmax1, __shfl_xor_sync(0xffffffffu, tmax1, 2));
    const float mn0 = fmaxf(m0, tmax0), mn1 = fmaxf(m1, tmax1);
    const float c0 = exp2f((m0 - mn0) * LOG2E), c1 = exp2f((m1 - mn1) * LOG2E);
    m0 = mn0;
    m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pa[KT / 16][4];
#pragma unroll
    for (int nt = 0; nt < KT / 8; ++nt) {
      float p00 = exp2f((s[nt][0] - mn0) * LOG2E), p01 = exp2f((s[nt][1] - mn0) * LOG2E);
      float p10 = exp2f((s[nt][2] - mn1) * LOG2E), p11 = exp2f((s[nt][3] - mn1) * LOG2E);
      ps0 += p00 + p01;
      ps1 += p10 + p11;
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(p00, p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p10, p11);
    }
    l0 = l0 * c0 + ps0;
    l1 = l1 * c1 + ps1;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      oacc[i][0] *= c0; oacc[i][1] *= c0;
      oacc[i][2] *= c1; oacc[i][3] *= c1;
    }
#pragma unroll
    for (int k16 = 0; k16 < KT / 16; ++k16) {
#pragma unroll
      for (int nd2 = 0; nd2 < DH / 16; ++nd2) {
        uint32_t bb[4];
        ldsm_x4_t(bb, Vs + (kt + k16 * 16 + vb_row) * LDS + nd2 * 16 + vb_col);
        mma_bf16(oacc[nd2 * 2], pa[k16], bb[0], bb[1]);
        mma_bf16(oacc[nd2 * 2 + 1], pa[k16], bb[2], bb[3]);
      }
    }
  }
  }  // key chunks
  if (!active) return;  // no block-wide sync below this point
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;

  // stage this warp's 16 output rows through its (now dead) Q rows, then 16-byte coalesced stores
  __syncwarp();
  bf16* Os = Qs + r0 * LDS;
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    *reinterpret_cast<uint32_t*>(Os + g * LDS + i * 8 + tq * 2) = pack_bf16(oacc[i][0] * i0, oacc[i][1] * i0);
    *reinterpret_cast<uint32_t*>(Os + (g + 8) * LDS + i * 8 + tq * 2) = pack_bf16(oacc[i][2] * i1, oacc[i][3] * i1);
  }
  __syncwarp();
  bf16* og = o + (b * Lq + q0) * ldo + (int64_t)h * DH;
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    int r = idx / CH, c = idx % CH;
    if (q0 + r0 + r < Lq)
      *reinterpret_cast<uint4*>(og + (int64_t)(r0 + r) * ldo + c * 8) = *reinterpret_cast<const uint4*>(Os + r * LDS + c * 8);
  }
  if (stats != nullptr && tq == 0) {
    int64_t base = (b * heads + h) * Lq + q0;
    if (q0 + r0 + g < Lq) { stats[(base + r0 + g) * 2] = m0; stats[(base + r0 + g) * 2 + 1] = i0; }
    if (q0 + r0 + g + 8 < Lq) { stats[(base + r0 + g + 8) * 2] = m1; stats[(base + r0 + g + 8) * 2 + 1] = i1; }
  }
}

template <int DH>
static int launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  void* o, int64_t ldo, const uint8_t* mask, float* stats, int64_t batch, int heads,
                  int Lq, int Lk, cudaStream_t st) {
  // whole-sequence tile when the keys fit (self-attention); otherwise 32 query rows per CTA and
  // the keys streamed in chunks of 128 (more CTAs, two or more resident per SM)
  const bool one_tile = Lk <= 256;
  const int LqPad = (Lq + 15) / 16 * 16;
  const int QR = one_tile ? LqPad : 32;
  const int KC = one_tile ? (Lk + KT - 1) / KT * KT : 128;
  const int q_tiles = (Lq + QR - 1) / QR;
  size_t smem = (size_t)(QR + 2 * KC) * (DH + 8) * 2 + KC;
  static size_t max_set = 0;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SPA3D_REQUIRE(e == cudaSuccess, "attention_mma: smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
    max_set = smem;
  }
  int64_t blocks = batch * heads * q_tiles;
  SPA3D_REQUIRE(blocks < (1ll << 31), "attention_mma: grid too large");
  attn_fwd_mma_kernel<DH><<<(unsigned)blocks, QR * 2, smem, st>>>(
      (const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)o, ldo, mask, stats, heads, Lq, Lk, QR, KC, q_tiles);
  return check_launch("attention_fwd_mma");
}

}  // namespace am

bool attention_fwd_mma_applicable(int dtype, int Lq, int Lk, int Dh, int64_t ldq, int64_t ldk,
                                  int64_t ldv, int64_t ldo) {
  if (dtype != SPA3D_BF16) return false;
  if (Dh != 64 && Dh != 96) return false;
  if (Lq > 256 && Lk <= 256) return false;  // one-tile mode keeps all queries of a sequence in one CTA
  return (ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (ldo % 8 == 0);
}

int attention_fwd_mma(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse_out,
                      int64_t batch, int heads, int Lq, int Lk, int Dh, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  SPA3D_REQUIRE(al(q) && al(k) && al(v) && al(o), "attention_mma: q/k/v/o must be 16-byte aligned");
  if (Dh == 96) return am::launch<96>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Lk, st);
  return am::launch<64>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse_out, batch, heads, Lq, Lk, st);
}

}  // namespace spa3d
