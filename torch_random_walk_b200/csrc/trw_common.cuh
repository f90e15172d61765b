// Shared device helpers for the sm_100a walk / window kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trw_b200.h"

namespace trw {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// RAII: make `device` current for the duration of one C-ABI call, then restore.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (device >= 0 && device != prev) ok = (cudaSetDevice(device) == cudaSuccess);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int sm_count(int device);

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al., SC'11).  Every draw in this library is a pure function of
// (seed, stream tag, item id, position), never of thread/block indices, so results do not
// depend on launch shape, sharding or GPU count.
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        uint32_t hi1 = mulhi32(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += kPhiloxW0;
        k.y += kPhiloxW1;
    }
    return c;
}

// Stream tags keep the entry points' random streams disjoint for equal seeds.
enum : uint32_t {
    kTagWalkCsr = 0x57414C4Bu,       // "WALK"
    kTagWalkEdgeList = 0x45444745u,  // "EDGE"
    kTagWalkTriples = 0x54524950u,   // "TRIP"
    kTagWindows = 0x57494E44u,       // "WIND"
};

__host__ __device__ __forceinline__ uint2 philox_key(int64_t seed, uint32_t tag) {
    uint64_t s = (uint64_t)seed;
    return make_uint2((uint32_t)s, (uint32_t)(s >> 32) ^ tag);
}

// Uniform integer in [0, n) from 32 random bits (n < 2^32) by multiply-shift.
__host__ __device__ __forceinline__ int64_t bounded32(uint32_t r, int64_t n) {
    return (int64_t)mulhi32(r, (uint32_t)n);
}
// General n (uses 64 random bits when n does not fit 32 bits).
__device__ __forceinline__ int64_t bounded(uint32_t r0, uint32_t r1, int64_t n) {
    if ((uint64_t)n <= 0xFFFFFFFFull) return (int64_t)__umulhi(r0, (uint32_t)n);
    return (int64_t)__umul64hi(((uint64_t)r0 << 32) | r1, (uint64_t)n);
}

// ---------------------------------------------------------------- memory helpers
__device__ __forceinline__ int64_t ldg64(const int64_t* p) { return __ldg(p); }

// Read-only 8-byte load that does not allocate in L1 (random gathers never re-hit it).
__device__ __forceinline__ int64_t ldg64_stream(const int64_t* p) {
    int64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// One full 32-byte sector in one instruction (LDG.E.256 on sm_100); p must be 32-byte aligned.
struct Sector64 { uint64_t a, b, c, d; };
__device__ __forceinline__ Sector64 ldg_sector(const void* p) {
    Sector64 s;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(s.a), "=l"(s.b), "=l"(s.c), "=l"(s.d) : "l"(p));
    return s;
}
__device__ __forceinline__ void stg_sector(void* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// L2 cache policies (createpolicy): evict_last for the small, re-read row index; evict_first
// for gathers and stores whose sectors are never touched again, so that the stream of random
// sectors does not displace the index from L2.
__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Store policy of the walk output by experiment mode (option store_mode): 0 evict_first (shipped),
// 1 evict_normal, 2 evict_last, 3 no stores at all (measures what the output costs; results are lost).
__device__ __forceinline__ uint64_t output_policy(int mode) {
    if (mode == 1) return make_policy_evict_normal();
    if (mode == 2) return make_policy_evict_last();
    if (mode == 3) return 0;
    return make_policy_evict_first();
}
__device__ __forceinline__ int64_t ldg32_keep(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return (int64_t)v;
}
__device__ __forceinline__ int64_t ldg64_keep(const int64_t* p, uint64_t pol) {
    int64_t v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int64_t ldg64_hint(const int64_t* p, uint64_t pol) {
    int64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ Sector64 ldg_sector_hint(const void* p, uint64_t pol) {
    Sector64 s;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=l"(s.a), "=l"(s.b), "=l"(s.c), "=l"(s.d) : "l"(p), "l"(pol));
    return s;
}
__device__ __forceinline__ void stg_sector_hint(void* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u64 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg64_hint(int64_t* p, int64_t v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

// Coherent 16-byte load of memory other threads are updating with atomics.
__device__ __forceinline__ uint4 ld_relaxed_u32x4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Murmur3 finaliser: node ids from R-MAT-like generators have very low-entropy bits.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// ---------------------------------------------------------------- staged row output
// Walk rows are [n, row_len] int64 with an odd row length in the common case (L+1 = 81), so a
// thread's consecutive 8-byte stores would each dirty a quarter of a 32-byte sector.  Each
// thread parks its elements in a 4-slot shared-memory ring laid out by the sector slot they
// will occupy in global memory and emits one 256-bit store per completed sector; only the
// ragged first/last sector of a row falls back to 8-byte stores.
template <int BLOCK>
struct RowStager {
    int64_t (*ring)[BLOCK];  // [4][BLOCK] in shared memory
    int64_t* row;            // this thread's output row
    uint64_t policy;         // L2 policy of the stores (evict_first: written once, never re-read)
    uint32_t phase;          // (address of row[0] / 8) & 3
    int tid;

    __device__ __forceinline__ void init(int64_t (*smem)[BLOCK], int64_t* row_, int tid_, uint64_t policy_) {
        ring = smem; row = row_; tid = tid_; policy = policy_;
        phase = (uint32_t)(((uintptr_t)row_ >> 3) & 3);
    }
    // Store element s (0-based, strictly increasing calls); `last` marks the final element.
    __device__ __forceinline__ void put(int s, int64_t v, bool last) {
        if (policy == 0) return;  // store_mode 3 (measurement only)
        uint32_t slot = (phase + (uint32_t)s) & 3u;
        ring[slot][tid] = v;
        if (slot == 3u || last) {
            int first = s - (int)slot;  // element index that sits in slot 0 of this sector
            if (slot == 3u && first >= 0) {
                stg_sector_hint(row + first, (uint64_t)ring[0][tid], (uint64_t)ring[1][tid],
                                (uint64_t)ring[2][tid], (uint64_t)ring[3][tid], policy);
            } else {
                int lo = first < 0 ? 0 : first;
                for (int e = lo; e <= s; ++e) stg64_hint(row + e, ring[(phase + (uint32_t)e) & 3u][tid], policy);
            }
        }
    }
};

}  // namespace trw
