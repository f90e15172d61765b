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
__device__ __forceinline__ uint2 ldg_u32x2_hint(const uint32_t* p, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
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
__device__ __forceinline__ void stg_u32x4_hint(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 ldg_u32x4_hint(const uint4* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg64_hint(int64_t* p, int64_t v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

// 4-byte load of L2-resident side data (the edge filter): kept in L2 (evict_last policy), not
// allocated in L1, where the in-flight gathers live.
__device__ __forceinline__ uint32_t ldg32_l2keep(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// Fire-and-forget OR into L2-resident side data (RED.OR, no return value), with an L2 policy.
__device__ __forceinline__ void red_or32_hint(uint32_t* p, uint32_t bits, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.or.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(bits), "l"(pol) : "memory");
}

// ---------------------------------------------------------------- CSR arrays of either width
// The reference's accessors are int64-only (csrc/cuda/rw_cuda.cu:206-209).  Here row_ptr and col_idx may be int64 or
// int32 arrays: IdxPtr carries the element size beside the address, every load widens to int64 (sign-extending), and
// the width test is uniform over the grid.  An int64 pointer converts implicitly, so int64 callers read as before.
struct IdxPtr {
    const char* base = nullptr;
    int shift = 3;  // log2 of the element size: 3 = int64, 2 = int32
    __host__ __device__ IdxPtr() {}
    __host__ __device__ IdxPtr(const int64_t* p) : base(reinterpret_cast<const char*>(p)), shift(3) {}
    __host__ __device__ IdxPtr(const int32_t* p) : base(reinterpret_cast<const char*>(p)), shift(2) {}
    __host__ __device__ IdxPtr(const void* p, int elem_bytes) : base(reinterpret_cast<const char*>(p)), shift(elem_bytes == 4 ? 2 : 3) {}
    __host__ __device__ IdxPtr operator+(int64_t i) const { IdxPtr r = *this; r.base += i * ((int64_t)1 << shift); return r; }
    __host__ __device__ IdxPtr operator-(int64_t i) const { return *this + (-i); }
    __host__ __device__ explicit operator bool() const { return base != nullptr; }
    __host__ __device__ bool wide() const { return shift == 3; }
    __host__ __device__ const int64_t* as64() const { return reinterpret_cast<const int64_t*>(base); }
    __host__ __device__ const int32_t* as32() const { return reinterpret_cast<const int32_t*>(base); }
};
// The same with the width fixed at compile time: what the streaming kernels use inside their loops (a run-time
// width test in front of every load keeps the compiler from issuing a thread's loads together).
template <bool WIDE>
struct IdxPtrT {
    const char* base;
    __device__ explicit IdxPtrT(IdxPtr p) : base(p.base) {}
    __device__ IdxPtrT(const char* b) : base(b) {}
    __device__ IdxPtrT operator+(int64_t i) const { return IdxPtrT(base + i * (WIDE ? 8 : 4)); }
    __device__ IdxPtrT operator-(int64_t i) const { return IdxPtrT(base - i * (WIDE ? 8 : 4)); }
    __device__ operator IdxPtr() const { return IdxPtr(base, WIDE ? 8 : 4); }
};
template <bool WIDE>
__device__ __forceinline__ int64_t ldg_idx(IdxPtrT<WIDE> p) {
    if (WIDE) return __ldg(reinterpret_cast<const int64_t*>(p.base));
    return (int64_t)__ldg(reinterpret_cast<const int32_t*>(p.base));
}
template <bool WIDE>
__device__ __forceinline__ int64_t ldg64_stream(IdxPtrT<WIDE> p) {
    if (WIDE) return ldg64_stream(reinterpret_cast<const int64_t*>(p.base));
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p.base));
    return (int64_t)v;
}
template <bool WIDE>
__device__ __forceinline__ int64_t ldg64_hint(IdxPtrT<WIDE> p, uint64_t pol) {
    if (WIDE) return ldg64_hint(reinterpret_cast<const int64_t*>(p.base), pol);
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p.base), "l"(pol));
    return (int64_t)v;
}

__device__ __forceinline__ int64_t ldg_idx(IdxPtr p) { return p.wide() ? __ldg(p.as64()) : (int64_t)__ldg(p.as32()); }
__device__ __forceinline__ int64_t ldg64_stream(IdxPtr p) {
    if (p.wide()) return ldg64_stream(p.as64());
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p.base));
    return (int64_t)v;
}
__device__ __forceinline__ int64_t ldg64_hint(IdxPtr p, uint64_t pol) {
    if (p.wide()) return ldg64_hint(p.as64(), pol);
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p.base), "l"(pol));
    return (int64_t)v;
}
__device__ __forceinline__ int64_t ldg64_keep(IdxPtr p, uint64_t pol) {
    if (p.wide()) return ldg64_keep(p.as64(), pol);
    int32_t v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p.base), "l"(pol));
    return (int64_t)v;
}

// Content checksum of CSR arrays (csr_checksum.cu): adds the position-sensitive sum of elements [base, base + n) of
// `arr` to *out_device; col_idx and row_ptr use different multipliers.  The host twin lives in walk_host.cu.
constexpr uint64_t kChecksumColGolden = 0x9E3779B97F4A7C15ull, kChecksumRowGolden = 0xD6E8FEB86659FD93ull;
int csr_checksum_part(IdxPtr arr, int64_t n, int64_t base, bool is_col_idx, uint64_t* out_device, int device, cudaStream_t st);

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


// ---------------------------------------------------------------- staged row output, 128-byte lines
// What bounds a random-gather kernel on B200 is the number of *requests* that leave L2 for memory
// (about 48 G/s: profiles/), and a store request counts like a missed load whatever its size.  The
// sector stager above issues one request per 32 bytes of output; this one parks 16 elements per
// thread (one 128-byte line of its row) in shared memory and lets the warp write finished lines
// cooperatively: 16 lanes x 8 bytes cover one whole line, so a store instruction carries two full
// lines in two requests, a quarter of the requests per byte.  put() is a warp collective: all 32
// lanes call it in convergence, `have` says whether this lane appends an element.
template <int BLOCK, int SLOTS = 16>
struct LineStager {
    static_assert(SLOTS == 4 || SLOTS == 8 || SLOTS == 16, "a piece is 32, 64 or 128 bytes");
    static constexpr int kPitch = SLOTS + 1;  // + 1: keeps the lanes' puts off a common bank for any row stride
    static constexpr int kOwners = 32 / SLOTS;  // pieces written per store instruction
    int64_t* ring;    // [BLOCK][kPitch] in shared memory
    int64_t* row;     // this lane's output row
    uint64_t policy;  // L2 policy of the stores
    uint32_t phase;   // (address of row[0] / 8) & (SLOTS - 1)
    int tid;

    __device__ __forceinline__ void init(int64_t* smem, int64_t* row_, int tid_, uint64_t policy_) {
        ring = smem; row = row_; tid = tid_; policy = policy_;
        phase = (uint32_t)(((uintptr_t)row_ >> 3) & (SLOTS - 1));
    }
    // Element s of the row (calls with have == true come with strictly increasing s); `last` marks the final one.
    __device__ __forceinline__ void put(bool have, int s, int64_t v, bool last) {
        if (policy == 0) return;  // store_mode 3 (measurement only)
        bool pending = false;
        uint32_t slot = 0;
        if (have) {
            slot = (phase + (uint32_t)s) & (uint32_t)(SLOTS - 1);
            ring[tid * kPitch + slot] = v;
            pending = (slot == (uint32_t)(SLOTS - 1)) || last;
        }
        flush(pending, s, slot);
    }
    // Four consecutive elements s4 .. s4+3 at once, for rows whose base and s4 are multiples of four elements (a group
    // then never straddles a line): one collective per group instead of four.
    __device__ __forceinline__ void put4(bool have, int s4, int64_t v0, int64_t v1, int64_t v2, int64_t v3, bool last) {
        if (policy == 0) return;
        bool pending = false;
        uint32_t slot = 0;
        if (have) {
            slot = (phase + (uint32_t)s4) & (uint32_t)(SLOTS - 1);
            int64_t* r = ring + tid * kPitch + slot;
            r[0] = v0; r[1] = v1; r[2] = v2; r[3] = v3;
            slot += 3u;
            pending = (slot == (uint32_t)(SLOTS - 1)) || last;
        }
        flush(pending, s4 + 3, slot);
    }
    // Warp collective: the lanes whose piece is complete (`pending`; `s` = index of the element in `slot`, the last one
    // filled) have it written by the warp, SLOTS lanes per piece.
    __device__ __forceinline__ void flush(bool pending, int s, uint32_t slot) {
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, pending);
        if (mask == 0) return;
        __syncwarp();
        // this lane's finished piece: slot 0 sits at row[first]; slots [lo, slot] hold elements of this row
        const int first = s - (int)slot;
        const unsigned long long piece_ptr = (unsigned long long)(uintptr_t)(row + first);
        const uint32_t range = (uint32_t)(first < 0 ? -first : 0) | (slot << 8);
        const int lane = tid & 31, sub = lane & (SLOTS - 1), group = lane / SLOTS;
        while (mask) {
            // the group-th pending lane serves as this lane's owner (-1: none left for this group)
            int owner = -1;
#pragma unroll
            for (int k = 0; k < kOwners; ++k) {
                if (k == group && mask) owner = __ffs(mask) - 1;
                mask &= mask - 1;  // no-op once zero
            }
            const int src = owner < 0 ? 0 : owner;
            const unsigned long long pp = __shfl_sync(0xFFFFFFFFu, piece_ptr, src);
            const uint32_t rg = __shfl_sync(0xFFFFFFFFu, range, src);
            if (owner >= 0 && (uint32_t)sub >= (rg & 0xFFu) && (uint32_t)sub <= (rg >> 8)) {
                const int64_t val = ring[((tid & ~31) + owner) * kPitch + sub];
                stg64_hint(reinterpret_cast<int64_t*>((uintptr_t)pp) + sub, val, policy);
            }
        }
        __syncwarp();
    }
};

}  // namespace trw
