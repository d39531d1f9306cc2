// Host-side staging for the end-to-end path (model.apply_stream(host_pack="bf16")).
//
// The reference hands the model float32 DINOv2 patch maps / per-track features on the HOST (inference.py:523-590); the bf16 path
// rounds them to bfloat16 on the device anyway (embed_tcgen05.cu producers).  Rounding them on the host cores instead, into a pinned
// staging ring, moves 2 bytes per value over PCIe instead of 4 - the end-to-end step is PCIe-bound otherwise.
// Round to nearest even, NaN stays a quiet NaN: bit-identical to __float2bfloat16_rn on the device and to torch's .to(bfloat16).
//
// Plain C ABI (include/spa3d_b200.h); no CUDA in this file, it is compiled by the host compiler and linked into lib3dspa_b200.so.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace {

inline uint16_t bf16_rn_scalar(uint32_t x) {
    if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);
    return (uint16_t)((x + 0x7fffu + ((x >> 16) & 1u)) >> 16);
}

void pack_scalar(const float* src, uint16_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t x;
        memcpy(&x, src + i, 4);
        dst[i] = bf16_rn_scalar(x);
    }
}

__attribute__((target("avx2"))) void pack_avx2(const float* src, uint16_t* dst, int64_t n) {
    const __m256i one = _mm256_set1_epi32(1), bias = _mm256_set1_epi32(0x7fff), absmask = _mm256_set1_epi32(0x7fffffff),
                  inf = _mm256_set1_epi32(0x7f800000), quiet = _mm256_set1_epi32(0x40);
    const bool aligned = (((uintptr_t)dst) & 31) == 0;
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        __m256i r[2];
        for (int h = 0; h < 2; ++h) {
            __m256i x = _mm256_loadu_si256((const __m256i*)(src + i + 8 * h));
            __m256i hi = _mm256_srli_epi32(x, 16);
            __m256i rn = _mm256_srli_epi32(_mm256_add_epi32(x, _mm256_add_epi32(bias, _mm256_and_si256(hi, one))), 16);
            __m256i isnan = _mm256_cmpgt_epi32(_mm256_and_si256(x, absmask), inf);
            r[h] = _mm256_blendv_epi8(rn, _mm256_or_si256(hi, quiet), isnan);
        }
        __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(r[0], r[1]), 0xD8);
        if (aligned) _mm256_stream_si256((__m256i*)(dst + i), p);
        else _mm256_storeu_si256((__m256i*)(dst + i), p);
    }
    _mm_sfence();
    pack_scalar(src + i, dst + i, n - i);
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) void pack_avx512(const float* src, uint16_t* dst, int64_t n) {
    const __m512i one = _mm512_set1_epi32(1), bias = _mm512_set1_epi32(0x7fff), absmask = _mm512_set1_epi32(0x7fffffff),
                  inf = _mm512_set1_epi32(0x7f800000), quiet = _mm512_set1_epi32(0x40);
    const bool aligned = (((uintptr_t)dst) & 63) == 0;
    int64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i r[2];
        for (int h = 0; h < 2; ++h) {
            __m512i x = _mm512_loadu_si512((const void*)(src + i + 16 * h));
            __m512i hi = _mm512_srli_epi32(x, 16);
            __m512i rn = _mm512_srli_epi32(_mm512_add_epi32(x, _mm512_add_epi32(bias, _mm512_and_si512(hi, one))), 16);
            __mmask16 isnan = _mm512_cmpgt_epi32_mask(_mm512_and_si512(x, absmask), inf);
            r[h] = _mm512_cvtepi32_epi16(_mm512_mask_mov_epi32(rn, isnan, _mm512_or_si512(hi, quiet)));
        }
        __m512i p = _mm512_inserti64x4(_mm512_castsi256_si512(r[0]), r[1], 1);
        if (aligned) _mm512_stream_si512((__m512i*)(dst + i), p);
        else _mm512_storeu_si512((void*)(dst + i), p);
    }
    _mm_sfence();
    pack_scalar(src + i, dst + i, n - i);
}

typedef void (*pack_fn)(const float*, uint16_t*, int64_t);

pack_fn select_pack() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) return pack_avx512;
    if (__builtin_cpu_supports("avx2")) return pack_avx2;
    return pack_scalar;
}

}  // namespace

extern "C" int spa3d_host_pack_bf16(const float* src, void* dst, int64_t n, int threads) {
    if (n < 0 || (n > 0 && (src == nullptr || dst == nullptr))) return 1;
    static const pack_fn fn = select_pack();
    uint16_t* out = (uint16_t*)dst;
    const int64_t grain = 1 << 16;   // values per work unit: 256 KB read, a multiple of every vector width (keeps stores aligned)
    int64_t units = (n + grain - 1) / grain;
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 1, units));
    if (nt == 1) {
        fn(src, out, n);
        return 0;
    }
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    auto work = [&](int t) {
        int64_t u0 = units * t / nt, u1 = units * (t + 1) / nt;
        int64_t b = u0 * grain, e = std::min(n, u1 * grain);
        if (e > b) fn(src + b, out + b, e - b);
    };
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    return 0;
}
