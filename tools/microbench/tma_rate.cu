// microbenchmark: TMA load rate for [rows x CH channels] boxes from a [B*L rows][ld] bf16 matrix, per SM, deep pipeline
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t d;
  do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(s32(b)), "r"(ph) : "memory"); } while (!d);
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"((uint64_t)m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// each "item" = (b, h): NOP operands x (96 / CH) boxes of [CH x 160]; STAGES items in flight
template <int CH, int STAGES>
__global__ void k(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2, int nop, int items, int heads) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[STAGES];
  constexpr int NB = 96 / CH, BOX = 160 * CH * 2;
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const CUtensorMap* maps[3] = {&m0, &m1, &m2};
  int issued = 0, waited = 0;
  for (int it = blockIdx.x; it < items || waited < issued; it += gridDim.x) {
    if (it < items) {
      if (issued - waited == STAGES) { mbar_wait(&full[waited % STAGES], (waited / STAGES) & 1); ++waited; }
      const int b = it / heads, h = it % heads, s = issued % STAGES;
      mbar_expect(&full[s], nop * NB * BOX);
      for (int o = 0; o < nop; ++o)
        for (int a = 0; a < NB; ++a) tma3(sm + (size_t)s * nop * NB * BOX + (o * NB + a) * BOX, maps[o % 3], h * 96 + a * CH, 0, b, &full[s]);
      ++issued;
    } else { mbar_wait(&full[waited % STAGES], (waited / STAGES) & 1); ++waited; }
  }
}
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fnp;
  const int B = 2048, L = 151, heads = 8, ld = 2304;   // qkv packed rows like the ITT
  void* buf; cudaMalloc(&buf, (size_t)B * L * ld * 2); cudaMemset(buf, 0, (size_t)B * L * ld * 2);
  void* flush; cudaMalloc(&flush, 256 << 20);
  auto mk = [&](CUtensorMap* m, int col0, int CH, CUtensorMapSwizzle sw) {
    cuuint64_t dims[3] = {768, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)CH, 160, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (char*)buf + col0 * 2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode failed %d\n", (int)r);
  };
  auto run = [&](auto kern, int CH, CUtensorMapSwizzle sw, int nop, int stages, const char* name) {
    CUtensorMap m[3];
    for (int o = 0; o < 3; ++o) mk(&m[o], o * 768, CH, sw);
    const int smem = stages * nop * 96 * 160 * 2 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemsetAsync(flush, rep, 256 << 20);
      cudaEventRecord(e0);
      kern<<<148, 32, smem>>>(m[0], m[1], m[2], nop, B * heads, heads);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)B * heads * nop * 96 * L * 2;
    printf("%-28s nop=%d stages=%d: %.3f ms  %.2f TB/s  (%s)\n", name, nop, stages, best, bytes / best / 1e9, cudaGetErrorString(err));
  };
  run(k<32, 1>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 3, 1, "64B rows, 1 item in flight");
  run(k<32, 2>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 3, 2, "64B rows, 2 items in flight");
  run(k<32, 1>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 4, 1, "64B rows, 4 ops, 1 in flight");
  run(k<32, 4>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 1, 4, "64B rows, 1 op x 4 in flight");
  run(k<96, 1>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 3, 1, "192B rows, 1 item in flight");
  run(k<96, 2>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 3, 2, "192B rows, 2 items in flight");
  run(k<96, 4>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 1, 4, "192B rows, 1 op x 4 in flight");
  return 0;
}
