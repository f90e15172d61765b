// Host memory throughput of the box for the three loops the end-to-end path runs on the host's cores (measurement tooling):
//   widen  uint32 -> int64 with non-temporal stores (4 B read + 8 B written per element)
//   sum    the position-sensitive 64-bit checksum of an int64 array (8 B read per entry)
//   copy   memcpy (8 B read + 8 B written per element, for scale)
// each on 1, 2, 4, 8, 16 threads over arrays far larger than the caches.
//
//   g++ -O3 -mavx2 -pthread tools/host_bw_probe.cpp -o /tmp/host_bw_probe && /tmp/host_bw_probe
#include <immintrin.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void widen(const uint32_t* src, int64_t* dst, int64_t n) {
    int64_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 4 <= n; i += 4) _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu32_epi64(_mm_loadu_si128((const __m128i*)(src + i))));
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static uint64_t sum(const int64_t* v, int64_t lo, int64_t hi) {
    const uint64_t golden = 0x9E3779B97F4A7C15ull;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t i = lo;
    for (; i + 4 <= hi; i += 4) {
        a0 += mix64((uint64_t)v[i] + golden * (uint64_t)(i + 1));
        a1 += mix64((uint64_t)v[i + 1] + golden * (uint64_t)(i + 2));
        a2 += mix64((uint64_t)v[i + 2] + golden * (uint64_t)(i + 3));
        a3 += mix64((uint64_t)v[i + 3] + golden * (uint64_t)(i + 4));
    }
    return a0 + a1 + a2 + a3;
}

template <class F>
static double run(int threads, F f) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> w;
    for (int t = 0; t < threads; ++t) w.emplace_back(f, t, threads);
    for (auto& x : w) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main() {
    const int64_t n = (int64_t)1 << 28;  // 1 GiB of uint32, 2 GiB of int64
    uint32_t* src = (uint32_t*)aligned_alloc(4096, (size_t)n * 4);
    int64_t* dst = (int64_t*)aligned_alloc(4096, (size_t)n * 8);
    int64_t* dst2 = (int64_t*)aligned_alloc(4096, (size_t)n * 8);
    for (int64_t i = 0; i < n; ++i) src[i] = (uint32_t)i;
    memset(dst, 1, (size_t)n * 8);
    memset(dst2, 1, (size_t)n * 8);
    volatile uint64_t sink = 0;
    for (int threads : {1, 2, 4, 8, 12, 16}) {
        double tw = 1e9, ts = 1e9, tc = 1e9;
        for (int rep = 0; rep < 2; ++rep) {
            tw = std::min(tw, run(threads, [&](int t, int nt) { widen(src + n * t / nt, dst + n * t / nt, n * (t + 1) / nt - n * t / nt); }));
            ts = std::min(ts, run(threads, [&](int t, int nt) { sink = sink + sum(dst, n * t / nt, n * (t + 1) / nt); }));
            tc = std::min(tc, run(threads, [&](int t, int nt) { memcpy(dst2 + n * t / nt, dst + n * t / nt, (size_t)(n * (t + 1) / nt - n * t / nt) * 8); }));
        }
        printf("threads %2d: widen %.1f GB/s of memory traffic (%.2f G elements/s) | sum %.1f GB/s | memcpy %.1f GB/s (read+write)\n", threads,
               n * 12.0 / tw / 1e9, n / tw / 1e9, n * 8.0 / ts / 1e9, n * 16.0 / tc / 1e9);
        fflush(stdout);
    }
    return 0;
}
