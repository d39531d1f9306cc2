// microbenchmark: TMA store rate for [32 rows x CH channels] boxes into a [B*L rows][ld] bf16 matrix
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st3(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)m), "r"(s32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int CH, int INFLIGHT>
__global__ void k(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2, int items, int heads) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x & 31) return;
  const CUtensorMap* maps[3] = {&m0, &m1, &m2};
  constexpr int NB = 96 / CH;
  int cnt = 0;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it / heads, h = it % heads;
    for (int u = w; u < 3 * 5 * NB; u += nw) {
      const int o = u / (5 * NB), c = (u / NB) % 5, a = u % NB;
      st3(maps[o], sm + (w * 2 + (cnt & 1)) * (32 * CH * 2), h * 96 + a * CH, c * 32, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(INFLIGHT) : "memory");
      ++cnt;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fnp;
  const int B = 2048, L = 151, heads = 8, ld = 2304;
  void* buf; cudaMalloc(&buf, (size_t)B * L * ld * 2);
  void* flush; cudaMalloc(&flush, 256 << 20);
  auto mk = [&](CUtensorMap* m, int col0, int CH, CUtensorMapSwizzle sw) {
    cuuint64_t dims[3] = {768, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)CH, 32, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (char*)buf + col0 * 2, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode failed %d\n", (int)r);
  };
  auto run = [&](auto kern, int CH, CUtensorMapSwizzle sw, int warps, const char* name) {
    CUtensorMap m[3];
    for (int o = 0; o < 3; ++o) mk(&m[o], o * 768, CH, sw);
    const int smem = warps * 2 * 32 * CH * 2 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemsetAsync(flush, rep, 256 << 20);
      cudaEventRecord(e0);
      kern<<<148, warps * 32, smem>>>(m[0], m[1], m[2], B * heads, heads);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)B * heads * 3 * 96 * L * 2;
    printf("%-36s warps=%2d: %.3f ms  %.2f TB/s  (%s)\n", name, warps, best, bytes / best / 1e9, cudaGetErrorString(err));
  };
  run(k<32, 1>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 1, "64B rows, 1 issuer, 1 pending");
  run(k<32, 1>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 4, "64B rows, 4 issuers, 1 pending");
  run(k<32, 1>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 12, "64B rows, 12 issuers, 1 pending");
  run(k<32, 0>, 32, CU_TENSOR_MAP_SWIZZLE_64B, 12, "64B rows, 12 issuers, 0 pending");
  run(k<96, 1>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 1, "192B rows, 1 issuer, 1 pending");
  run(k<96, 1>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 4, "192B rows, 4 issuers, 1 pending");
  run(k<96, 0>, 96, CU_TENSOR_MAP_SWIZZLE_NONE, 4, "192B rows, 4 issuers, 0 pending");
  return 0;
}
