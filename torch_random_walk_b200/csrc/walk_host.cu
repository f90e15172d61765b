// trw_walk_csr_host / trw_walk_csr_to_host: the CSR walk for callers whose tensors live in host memory.
//
// The reference has no such path (a CPU tensor simply ran csrc/cpu); a drop-in user who holds
// CPU tensors still has to get them to the GPU and the walks back, and that round trip is what
// an end-to-end measurement pays for.  So it is engineered rather than left to three blocking
// copies:
//   * the start nodes are walked in chunks on one stream and each finished chunk is copied back on a
//     second stream while the next chunk is being walked (PCIe is full duplex, the copy engines run
//     beside the SMs);
//   * node ids fit 32 bits for every graph the fast paths take and the host can narrow/widen with its
//     cores (14 G entries/s on 16 threads), so with enough threads col_idx goes up as uint32, and the
//     walks come back as uint32 for as many chunks as the host team keeps up with (HostTeam below; only
//     while the rank's own PCIe link is the limit: download_mode), converted while the neighbouring
//     chunks are in flight; any id that does not fit falls back to plain copies;
//   * the graph does not cross PCIe again when the caller comes back with the same arrays: the device
//     replica and its preparation are kept per device, keyed by the host pointers and sizes and
//     VALIDATED BY CONTENT on every call -- row_ptr and col_idx are summed again (the same 64-bit
//     position-sensitive sum the device computed over the replica when it was uploaded,
//     csr_checksum.cu; part by the copy engine re-reading pinned arrays, part by the host threads)
//     while the walk already runs on the kept replica; a mismatch throws that work away, uploads
//     afresh and walks again.  Like the device-side graph cache the preparation grows
//     with use (per-call needs, then the full kept preparation, then the triangle Blooms).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "trw_common.cuh"
#include "trw_options.h"
#include "walk_csr.h"

namespace trw {

// Device buffers, streams and events of the host path are kept between calls (per device,
// grow-only): cudaMalloc/cudaFree of ~15 GB cost more than the walk itself.  Released by
// trw_release_cached_buffers() or with option host_cache_buffers = 0.
constexpr int kRing = 3;  // chunks in flight on the way down (walk buffers, packed staging)
enum { kBufRowPtr = 0, kBufColIdx, kBufTargets, kBufWorkspace, kBufOut0, kBufOut1, kBufOut2, kBufUp0, kBufUp1, kBufDown0, kBufDown1, kBufDown2,
       kBufCheck, kNumBufs };
enum { kPinUp0 = 0, kPinUp1, kPinDown0, kPinDown1, kPinDown2, kNumPinned };
struct HostWalkCache {
    void* ptr[kNumBufs] = {};
    size_t cap[kNumBufs] = {};
    void* pinned[kNumPinned] = {};  // host staging of the compressed transfers
    size_t pinned_cap[kNumPinned] = {};
    cudaStream_t compute = nullptr, copy = nullptr, check = nullptr;
    cudaEvent_t walked[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr}, uploaded[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> landed;  // download pipeline: one event per chunk (the host waits on them in its own order)
    std::mutex mu;  // one host-path call at a time per device
    // the kept replica: which host arrays it mirrors, its device-side checksum, and what has been prepared on it
    const void* key_row_ptr = nullptr;
    const void* key_col_idx = nullptr;
    int64_t key_n_nodes = -1, key_nnz = -1;
    uint64_t replica_checksum = 0;
    bool have_replica = false;
    int64_t up_bytes = 0, down_bytes = 0;  // what the last call moved over PCIe in each direction (trw_host_replica_info)
    int last_call = 0;    // how the last call through this cache got its graph: 1 kept replica validated, 2 fresh upload, 3 kept replica
                          // found changed (then uploaded afresh)
    int level = -1;       // -1 nothing prepared, 0 what one call's (p, q) needed, 1 the full kept preparation, 2 with triangle Blooms
    int hits = 0;
    CsrGraph graph;
    void forget_replica() { have_replica = false; level = -1; hits = 0; key_row_ptr = key_col_idx = nullptr; key_n_nodes = key_nnz = -1; }
    void release() {
        forget_replica();
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
        if (check) cudaStreamSynchronize(check);
        for (int k = 0; k < kNumBufs; ++k) {
            if (ptr[k]) cudaFree(ptr[k]);
            ptr[k] = nullptr;
            cap[k] = 0;
        }
        for (int k = 0; k < kNumPinned; ++k) {
            if (pinned[k]) cudaFreeHost(pinned[k]);
            pinned[k] = nullptr;
            pinned_cap[k] = 0;
        }
        for (int k = 0; k < 2; ++k) {
            if (walked[k]) cudaEventDestroy(walked[k]);
            if (copied[k]) cudaEventDestroy(copied[k]);
            if (uploaded[k]) cudaEventDestroy(uploaded[k]);
            walked[k] = copied[k] = uploaded[k] = nullptr;
        }
        for (cudaEvent_t e : landed) cudaEventDestroy(e);
        landed.clear();
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        if (check) cudaStreamDestroy(check);
        compute = copy = check = nullptr;
        cudaGetLastError();
    }
    int reserve(int slot, size_t bytes, const char* what) {
        if (bytes == 0) bytes = 8;
        if (cap[slot] >= bytes) return TRW_OK;
        if (ptr[slot]) cudaFree(ptr[slot]);
        ptr[slot] = nullptr;
        cap[slot] = 0;
        int rc = check_cuda(cudaMalloc(&ptr[slot], bytes), what);
        if (rc == TRW_OK) cap[slot] = bytes;
        return rc;
    }
    int reserve_pinned(int slot, size_t bytes, const char* what) {
        if (bytes == 0) bytes = 8;
        if (pinned_cap[slot] >= bytes) return TRW_OK;
        if (pinned[slot]) cudaFreeHost(pinned[slot]);
        pinned[slot] = nullptr;
        pinned_cap[slot] = 0;
        int rc = check_cuda(cudaHostAlloc(&pinned[slot], bytes, cudaHostAllocDefault), what);
        if (rc == TRW_OK) pinned_cap[slot] = bytes;
        return rc;
    }
};

// f(tid, n_threads) on n_threads host threads (the caller is thread 0).  Threads are created per call: a chunk
// is milliseconds of work, thread creation tens of microseconds, and no pool outlives the call.
template <class F>
static void parallel_for(int n_threads, F f) {
    std::vector<std::thread> workers;
    workers.reserve(n_threads > 1 ? n_threads - 1 : 0);
    int started = 1;  // slice 0 is the caller's
    try {
        for (; started < n_threads; ++started) workers.emplace_back(f, started, n_threads);
    } catch (...) {
        // the system refused a thread: the caller works through the slices that got none
    }
    f(0, n_threads);
    for (int t = started; t < n_threads; ++t) f(t, n_threads);
    for (auto& w : workers) w.join();
}

// The two host-side conversions.  The destination is written once and never read here, so the AVX2 forms use
// non-temporal stores: no read-for-ownership of the destination lines (a third of the memory traffic of the
// widening).  Both return true when every source value fits 32 bits unsigned.
static bool narrow_scalar(const int64_t* src, uint32_t* dst, int64_t n) {
    uint64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { acc |= (uint64_t)src[i]; dst[i] = (uint32_t)src[i]; }
    return (acc >> 32) == 0;
}
static void widen_scalar(const uint32_t* src, int64_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (int64_t)src[i];
}
#if defined(__x86_64__)
__attribute__((target("avx2"))) static bool narrow_avx2(const int64_t* src, uint32_t* dst, int64_t n) {
    int64_t i = 0;
    uint64_t acc = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { acc |= (uint64_t)src[i]; dst[i] = (uint32_t)src[i]; ++i; }
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
    __m256i any = _mm256_setzero_si256();
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 4));
        any = _mm256_or_si256(any, _mm256_or_si256(a, b));
        const __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, pick));
        const __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, pick));
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_set_m128i(hi, lo));
    }
    alignas(32) uint64_t lanes[4];
    _mm256_store_si256((__m256i*)lanes, any);
    acc |= lanes[0] | lanes[1] | lanes[2] | lanes[3];
    for (; i < n; ++i) { acc |= (uint64_t)src[i]; dst[i] = (uint32_t)src[i]; }
    _mm_sfence();
    return (acc >> 32) == 0;
}
__attribute__((target("avx2"))) static void widen_avx2(const uint32_t* src, int64_t* dst, int64_t n) {
    int64_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = (int64_t)src[i]; ++i; }
    for (; i + 4 <= n; i += 4)
        _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu32_epi64(_mm_loadu_si128((const __m128i*)(src + i))));
    for (; i < n; ++i) dst[i] = (int64_t)src[i];
    _mm_sfence();
}
#endif
static bool narrow_to_u32(const int64_t* src, uint32_t* dst, int64_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return narrow_avx2(src, dst, n);
#endif
    return narrow_scalar(src, dst, n);
}
static void widen_to_i64(const uint32_t* src, int64_t* dst, int64_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return widen_avx2(src, dst, n);
#endif
    widen_scalar(src, dst, n);
}

// Threads a rank needs before the UPLOAD of a fresh graph narrows col_idx on the host (round 1, 16 cores: 16 threads per
// rank turn 190 ms into 144 ms per C3 call; with 8 the narrowing does not keep up with the upload).  The download format
// is decided separately (download_mode).
constexpr int kMinCompressThreads = 12;

// Host threads available to this process for the wire compression: the machine's, shared among the ranks of
// a torchrun launch on this node (LOCAL_WORLD_SIZE), or option host_threads.
// Ranks of a torchrun launch that share this host (LOCAL_WORLD_SIZE), 1 otherwise.
static int local_rank_count() {
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    const int ranks = lws ? atoi(lws) : 1;
    return ranks > 1 ? ranks : 1;
}
static int host_thread_count() {
    if (options().host_threads > 0) return (int)options().host_threads;
    int n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    n /= local_rank_count();
    return n > 0 ? n : 1;
}

__global__ void __launch_bounds__(256) widen_u32_kernel(const uint32_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n) {
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gsz) dst[i] = (int64_t)src[i];
}
// Narrows a chunk of walks; raises *overflow when an entry does not fit (it cannot for ids that were checked on
// the way up, but the flag is what the host trusts).
__global__ void __launch_bounds__(256) narrow_i64_kernel(const int64_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n,
                                                         int* __restrict__ overflow) {
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gsz) {
        const int64_t v = src[i];
        bad |= (uint64_t)v > 0xFFFFFFFFull;
        dst[i] = (uint32_t)v;
    }
    if (bad) *overflow = 1;
}
static HostWalkCache g_host_cache[64];

#define TRW_TRY(expr, what)                       \
    do {                                          \
        int rc__ = check_cuda((expr), what);      \
        if (rc__) return rc__;                    \
    } while (0)

// The host twin of csr_checksum_kernel (csr_checksum.cu): same sum, so a host array can be compared with its device replica.
static inline uint64_t mix64_host(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static uint64_t checksum_range_scalar(const int64_t* v, int64_t lo, int64_t hi, uint64_t golden) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // four independent chains keep the multipliers busy
    int64_t i = lo;
    for (; i + 4 <= hi; i += 4) {
        a0 += mix64_host((uint64_t)v[i] + golden * (uint64_t)(i + 1));
        a1 += mix64_host((uint64_t)v[i + 1] + golden * (uint64_t)(i + 2));
        a2 += mix64_host((uint64_t)v[i + 2] + golden * (uint64_t)(i + 3));
        a3 += mix64_host((uint64_t)v[i + 3] + golden * (uint64_t)(i + 4));
    }
    for (; i < hi; ++i) a0 += mix64_host((uint64_t)v[i] + golden * (uint64_t)(i + 1));
    return a0 + a1 + a2 + a3;
}
#if defined(__x86_64__)
// The scalar loop is bound by its two 64-bit multiplies per entry (6.4 GB/s per core measured on the GPU box, a sixth of
// what a core streams); AVX-512DQ has the 64-bit low multiply, eight entries at a time.  Same sum: all arithmetic is mod 2^64.
__attribute__((target("avx512f,avx512dq"))) static uint64_t checksum_range_avx512(const int64_t* v, int64_t lo, int64_t hi, uint64_t golden) {
    const __m512i c1 = _mm512_set1_epi64((long long)0xBF58476D1CE4E5B9ull), c2 = _mm512_set1_epi64((long long)0x94D049BB133111EBull);
    const __m512i g = _mm512_set1_epi64((long long)golden), step = _mm512_set1_epi64((long long)(golden * 16u));
    __m512i pos0 = _mm512_mullo_epi64(_mm512_add_epi64(_mm512_set1_epi64(lo + 1), _mm512_setr_epi64(0, 1, 2, 3, 4, 5, 6, 7)), g);  // golden * (i + 1 + lane)
    __m512i pos1 = _mm512_add_epi64(pos0, _mm512_set1_epi64((long long)(golden * 8u)));
    __m512i a0 = _mm512_setzero_si512(), a1 = _mm512_setzero_si512();
    int64_t i = lo;
    for (; i + 16 <= hi; i += 16) {
        __m512i z0 = _mm512_add_epi64(_mm512_loadu_si512(v + i), pos0);
        __m512i z1 = _mm512_add_epi64(_mm512_loadu_si512(v + i + 8), pos1);
        z0 = _mm512_mullo_epi64(_mm512_xor_si512(z0, _mm512_srli_epi64(z0, 30)), c1);
        z1 = _mm512_mullo_epi64(_mm512_xor_si512(z1, _mm512_srli_epi64(z1, 30)), c1);
        z0 = _mm512_mullo_epi64(_mm512_xor_si512(z0, _mm512_srli_epi64(z0, 27)), c2);
        z1 = _mm512_mullo_epi64(_mm512_xor_si512(z1, _mm512_srli_epi64(z1, 27)), c2);
        a0 = _mm512_add_epi64(a0, _mm512_xor_si512(z0, _mm512_srli_epi64(z0, 31)));
        a1 = _mm512_add_epi64(a1, _mm512_xor_si512(z1, _mm512_srli_epi64(z1, 31)));
        pos0 = _mm512_add_epi64(pos0, step);
        pos1 = _mm512_add_epi64(pos1, step);
    }
    return (uint64_t)_mm512_reduce_add_epi64(_mm512_add_epi64(a0, a1)) + checksum_range_scalar(v, i, hi, golden);
}
#endif
static uint64_t checksum_range(const int64_t* v, int64_t lo, int64_t hi, uint64_t golden, bool simd = true) {
#if defined(__x86_64__)
    static const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
    if (simd && avx512) return checksum_range_avx512(v, lo, hi, golden);
#endif
    return checksum_range_scalar(v, lo, hi, golden);
}

constexpr int kRetryPlain = 1;  // host_pipeline: a walk entry did not fit the uint32 wire format

// Is this host pointer page-locked (cudaHostAlloc / cudaHostRegister / a pinned torch tensor)?  Only then can the copy
// engine read it asynchronously.
static bool host_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

struct HostCallShape {
    int64_t n_walks, walk_id_offset, id_block, id_stride;
    int walk_length;
};

// ---- the host's share of one call ----------------------------------------------------------------------------------
// Two kinds of work want the host's cores while the chunks of a call are in flight: widening packed chunks that have
// landed and summing the host CSR for the kept replica's content check.  One team of threads serves both.  Neither
// split is fixed in advance, because what the host can carry differs from box to box and from minute to minute (its
// memory bandwidth is shared with the DMA traffic and with whatever else runs on the machine):
//   * the pieces of the checksum sit in one list that the copy engine eats from the front (re-reading pinned arrays over
//     the upload direction, summed on the device) and the team from the back, until they meet;
//   * a worker sums before it widens whenever the checksum lags behind the chunks (so the sums end with the download,
//     not after it), and the pipeline sends a chunk packed only while the team keeps up with the widening (below).
// The calling thread runs the pipeline and lends a hand with widening whenever it has to wait.
constexpr int64_t kWidenPiece = (int64_t)1 << 18;  // elements per piece: 1 MB of staging in, 2 MB out
constexpr int64_t kMaxSumPiece = (int64_t)1 << 22;  // entries per piece of the checksum list: at most 32 MB (option host_sum_piece)

struct HostSum {  // entries [lo, hi) of an array of the host CSR, summed as csr_checksum_part sums its replica
    const int64_t* v;
    int64_t lo, hi;
    uint64_t golden;
};
struct WidenJob {
    const uint32_t* src = nullptr;
    int64_t* dst = nullptr;
    int64_t n = 0, pieces = 0;
    std::atomic<int64_t> next{0}, done{0};
    bool finished() const { return done.load(std::memory_order_acquire) >= pieces; }
};

class HostTeam {
  public:
    HostTeam(int n_threads, int max_jobs, const HostSum* sums, int n_sums, int64_t sum_piece) : jobs_((size_t)std::max(max_jobs, 0)) {
        for (int k = 0; k < n_sums; ++k)
            for (int64_t lo = sums[k].lo; lo < sums[k].hi; lo += sum_piece)
                sum_pieces_.push_back(HostSum{sums[k].v, lo, std::min(lo + sum_piece, sums[k].hi), sums[k].golden});
        back_ = (int64_t)sum_pieces_.size();
        if (jobs_.empty() && sum_pieces_.empty()) return;
        try {
            for (int t = 0; t < n_threads; ++t) workers_.emplace_back([this] { run(); });
        } catch (...) {
            // the system refused a thread: the calling thread works through what the others leave (drain())
        }
    }
    ~HostTeam() {
        abort_.store(true, std::memory_order_release);
        join();
    }
    // A packed chunk is on its way into `src`: job j = jobs added so far.  Widened once publish() has passed it.
    int add_job(const uint32_t* src, int64_t* dst, int64_t n) {
        WidenJob& w = jobs_[(size_t)added_];
        w.src = src; w.dst = dst; w.n = n;
        w.pieces = (n + kWidenPiece - 1) / kWidenPiece;
        return added_++;
    }
    int jobs_added() const { return added_; }
    bool job_finished(int j) const { return j < 0 || jobs_[(size_t)j].finished(); }
    void publish(int n) { published_.store(n, std::memory_order_release); }  // jobs [0, n) have landed in their staging
    // landed jobs the team has not finished widening: how far the host is behind
    int backlog() {
        const int pub = published_.load(std::memory_order_relaxed);
        while (caller_done_ < pub && jobs_[(size_t)caller_done_].finished()) ++caller_done_;
        int n = 0;
        for (int j = caller_done_; j < pub; ++j) n += jobs_[(size_t)j].finished() ? 0 : 1;
        return n;
    }
    // How far the download has come, in 1/1024: the sums are kept level with it.
    void set_progress(int64_t done, int64_t total) { progress_.store(total > 0 ? done * 1024 / total : 1024, std::memory_order_relaxed); }
    // The copy engine's side of the checksum list: the next piece from the front, while it may take one.
    bool claim_front(HostSum* out, int64_t max_front) {
        std::lock_guard<std::mutex> lock(mu_);
        if (front_ >= back_ || front_ >= max_front) return false;
        *out = sum_pieces_[(size_t)front_++];
        return true;
    }
    bool help_widen() { return widen_once(caller_first_); }  // one piece of widening on the calling thread, if there is any
    // No more jobs will be added: everything added is widened and every sum piece summed on return (the caller works
    // along; the copy engine's pieces are the caller's to wait for).  Returns the team's sum.
    uint64_t drain() {
        final_jobs_.store(added_, std::memory_order_release);
        publish(added_);
        for (;;) {
            if (widen_once(caller_first_) || sum_once(total_caller_)) continue;
            bool all = sums_taken_.load(std::memory_order_acquire) == sums_done_.load(std::memory_order_acquire) && !sums_left();
            for (int j = 0; all && j < added_; ++j) all = jobs_[(size_t)j].finished();
            if (all) break;
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        join();
        return total_.load(std::memory_order_acquire) + total_caller_;
    }

  private:
    bool sums_left() {
        std::lock_guard<std::mutex> lock(mu_);
        return front_ < back_;
    }
    bool widen_once(int& first) {
        const int pub = published_.load(std::memory_order_acquire);
        for (int j = first; j < pub; ++j) {
            WidenJob& w = jobs_[(size_t)j];
            const int64_t pc = w.next.load(std::memory_order_relaxed) < w.pieces ? w.next.fetch_add(1, std::memory_order_relaxed) : w.pieces;
            if (pc >= w.pieces) {
                if (j == first) ++first;
                continue;
            }
            const int64_t lo = pc * kWidenPiece;
            widen_to_i64(w.src + lo, w.dst + lo, std::min(kWidenPiece, w.n - lo));
            w.done.fetch_add(1, std::memory_order_release);
            return true;
        }
        return false;
    }
    bool sum_once(uint64_t& acc) {
        HostSum s;
        {
            std::lock_guard<std::mutex> lock(mu_);
            if (front_ >= back_) return false;
            s = sum_pieces_[(size_t)--back_];
            sums_taken_.fetch_add(1, std::memory_order_relaxed);
        }
        acc += checksum_range(s.v, s.lo, s.hi, s.golden);
        sums_done_.fetch_add(1, std::memory_order_release);
        return true;
    }
    bool sums_lag() {  // fewer pieces are gone (either way) than the download's progress asks for
        if (sum_pieces_.empty()) return false;
        std::lock_guard<std::mutex> lock(mu_);
        const int64_t n = (int64_t)sum_pieces_.size(), gone = front_ + (n - back_);
        return front_ < back_ && gone * 1024 < progress_.load(std::memory_order_relaxed) * n;
    }
    void run() {
        uint64_t acc = 0;
        int first = 0;
        while (!abort_.load(std::memory_order_acquire)) {
            if (sums_lag() && sum_once(acc)) continue;
            if (widen_once(first) || sum_once(acc)) continue;
            if (first >= final_jobs_.load(std::memory_order_acquire)) break;  // every job widened or taken, no sum piece left
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        total_.fetch_add(acc, std::memory_order_acq_rel);
    }
    void join() {
        for (auto& w : workers_)
            if (w.joinable()) w.join();
    }
    std::vector<WidenJob> jobs_;
    std::vector<HostSum> sum_pieces_;
    std::vector<std::thread> workers_;
    std::mutex mu_;            // the two ends of the checksum list
    int64_t front_ = 0, back_ = 0;
    std::atomic<int> published_{0}, final_jobs_{1 << 30};
    std::atomic<int64_t> sums_taken_{0}, sums_done_{0}, progress_{0};
    std::atomic<uint64_t> total_{0};
    std::atomic<bool> abort_{false};
    uint64_t total_caller_ = 0;
    int added_ = 0, caller_first_ = 0, caller_done_ = 0;
};

// What the kept replica's content check asks of a pipeline run: the arrays to sum and who may sum them.
struct ContentCheck {
    HostSum sums[2];
    int n_sums = 0;
    int dma_eighths = 0;  // the copy engine may take up to this many eighths of the pieces (0: the arrays are pageable, or option host_check_dma = 0)
    uint64_t sum = 0;     // out: the checksum of the host arrays
};

// Walks `shape.n_walks` start nodes (already on the device) in chunks and streams the chunks into host memory, two
// copies queued at a time.  A packed chunk is narrowed on the device, copied as uint32 into pinned staging and widened into
// `out` by the host team; a plain chunk is int64 rows copied straight into `out`.  `mode` 0..8: that many of every 8
// chunks packed (0 plain copies, 8 everything packed).  `mode` < 0: decided chunk by chunk -- packed while the team keeps
// up (at most one landed chunk still being widened), plain otherwise, so the copy engine and the host's cores each carry
// what they can.  With `check`, the team and the copy engine also sum the host CSR (ContentCheck).  Waits for everything
// before it returns.
static int host_pipeline(HostWalkCache& r, int d, const CsrWalkPlan& plan, const int64_t* d_targets, const HostCallShape& sh,
                         int64_t* out, int mode, int n_threads, ContentCheck* check = nullptr) {
    const int64_t row_len = (int64_t)sh.walk_length + 1;
    // Chunks of host_chunk_walks, but at least eight of them where the list allows (a rank's shard of a sharded list is
    // short: with two chunks the first walk and the last copy are not overlapped with anything; 8 ranks on c3: 67 -> 62 ms).
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(options().host_chunk_walks, sh.n_walks));
    {
        const int64_t unit = sh.id_block > 0 ? sh.id_block : 4096, eighth = ((sh.n_walks + 7) / 8 + unit - 1) / unit * unit;
        if (eighth < chunk && eighth >= 16 * unit && chunk % unit == 0) chunk = eighth;
    }
    if (sh.id_block > 0 && chunk % sh.id_block != 0 && sh.n_walks > chunk) {
        set_error("host walk: host_chunk_walks must be a multiple of the walk-id block (%lld)", (long long)sh.id_block);
        return TRW_ERR_ARG;
    }
    const int n_chunks = (int)((sh.n_walks + chunk - 1) / chunk);
    const int n_buf = std::min(n_chunks, kRing);
    const bool adaptive = mode < 0;
    auto packed_by_pattern = [mode](int c) { return ((c + 1) * mode) / 8 != (c * mode) / 8; };  // `mode` of every 8 chunks, evenly spread
    int rc = TRW_OK;
    for (int k = 0; k < n_buf && !rc; ++k) rc = r.reserve(kBufOut0 + k, (size_t)chunk * row_len * 8, "cudaMalloc walks");
    const size_t down_bytes = (size_t)chunk * row_len * 4;
    for (int k = 0; k < n_buf && mode != 0 && !rc; ++k) {
        rc = r.reserve(kBufDown0 + k, down_bytes + 256, "cudaMalloc download staging");
        if (!rc) rc = r.reserve_pinned(kPinDown0 + k, down_bytes, "cudaHostAlloc download staging");
    }
    const int dma_eighths = check ? check->dma_eighths : 0;
    const int64_t sum_piece = std::min(std::max<int64_t>(options().host_sum_piece, 16), kMaxSumPiece);
    if (!rc && dma_eighths > 0) rc = r.reserve(kBufCheck, 2 * (size_t)kMaxSumPiece * 8 + 256, "cudaMalloc check buffer");
    if (!rc && dma_eighths > 0) rc = r.reserve_pinned(kPinUp0, 256, "cudaHostAlloc checksum cell");
    if (rc) return rc;
    while ((int)r.landed.size() < n_chunks + 2) {
        cudaEvent_t e;
        TRW_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create");
        r.landed.push_back(e);
    }
    cudaEvent_t* const summed = r.landed.data() + n_chunks;  // the two side buffers of the copy engine's checksum pieces
    int* d_overflow = mode != 0 ? (int*)((char*)r.ptr[kBufDown0] + down_bytes) : nullptr;  // the 256 spare bytes
    if (mode != 0) TRW_TRY(cudaMemsetAsync(d_overflow, 0, sizeof(int), r.compute), "overflow flag memset");
    uint64_t* d_sum = dma_eighths > 0 ? (uint64_t*)((char*)r.ptr[kBufCheck] + 2 * (size_t)kMaxSumPiece * 8) : nullptr;
    if (d_sum) TRW_TRY(cudaMemsetAsync(d_sum, 0, sizeof(uint64_t), r.check), "check sum memset");

    HostTeam team(n_threads, mode != 0 ? n_chunks : 0, check ? check->sums : nullptr, check ? check->n_sums : 0, sum_piece);
    std::vector<int> job_of_chunk((size_t)n_chunks, -1), chunk_of_job;
    int64_t n_sum_pieces = 0;
    if (check)
        for (int k = 0; k < check->n_sums; ++k) n_sum_pieces += (check->sums[k].hi - check->sums[k].lo + sum_piece - 1) / sum_piece;
    const int64_t dma_max = n_sum_pieces * dma_eighths / 8;
    int dma_issued = 0;
    // the copy engine's share of the checksum: keep its two side buffers busy
    auto feed_check = [&]() -> int {
        while (dma_issued < dma_max) {
            const int slot = dma_issued & 1;
            if (dma_issued >= 2) {
                const cudaError_t st = cudaEventQuery(summed[slot]);
                if (st == cudaErrorNotReady) return TRW_OK;
                if (st != cudaSuccess) return check_cuda(st, "query check piece");
            }
            HostSum pc;
            if (!team.claim_front(&pc, dma_max)) { dma_issued = (int)dma_max; return TRW_OK; }
            int64_t* side = (int64_t*)r.ptr[kBufCheck] + (size_t)slot * kMaxSumPiece;
            TRW_TRY(cudaMemcpyAsync(side, pc.v + pc.lo, (size_t)(pc.hi - pc.lo) * 8, cudaMemcpyHostToDevice, r.check), "H2D check piece");
            r.up_bytes += (pc.hi - pc.lo) * 8;
            const int rc2 = csr_checksum_part(IdxPtr((const int64_t*)side), pc.hi - pc.lo, pc.lo, pc.golden == kChecksumColGolden, d_sum, d, r.check);
            if (rc2) return rc2;
            TRW_TRY(cudaEventRecord(summed[slot], r.check), "record check piece");
            ++dma_issued;
        }
        return TRW_OK;
    };
    auto enqueue = [&](int c, bool packed) -> int {
        const int b = c % kRing;
        const int64_t done = (int64_t)c * chunk, m = std::min(chunk, sh.n_walks - done);
        if (c >= kRing) TRW_TRY(cudaStreamWaitEvent(r.compute, r.landed[(size_t)(c - kRing)], 0), "wait landed");  // its walk buffer is free
        // ids of this chunk: contiguous shards advance the offset; block-cyclic ones advance it by whole strides
        const int64_t off = sh.id_block > 0 ? sh.walk_id_offset + (done / sh.id_block) * sh.id_stride : sh.walk_id_offset + done;
        int rc2 = csr_walk_launch(plan, d_targets + done, m, off, (int64_t*)r.ptr[kBufOut0 + b], row_len, r.compute, sh.id_block, sh.id_stride);
        if (rc2) return rc2;
        int j = -1;
        if (packed) {  // (its staging is free: the job that used it last has been widened, so its copy is long done)
            j = team.add_job((const uint32_t*)r.pinned[kPinDown0 + team.jobs_added() % kRing], out + done * row_len, m * row_len);
            job_of_chunk[(size_t)c] = j;
            chunk_of_job.push_back(c);
            narrow_i64_kernel<<<sm_count(d) * 8, 256, 0, r.compute>>>((const int64_t*)r.ptr[kBufOut0 + b], (uint32_t*)r.ptr[kBufDown0 + j % kRing],
                                                                     m * row_len, d_overflow);
            count_launch(1);
        }
        TRW_TRY(cudaEventRecord(r.walked[c & 1], r.compute), "record walked");
        TRW_TRY(cudaStreamWaitEvent(r.copy, r.walked[c & 1], 0), "wait walked");
        if (packed)
            TRW_TRY(cudaMemcpyAsync(r.pinned[kPinDown0 + j % kRing], r.ptr[kBufDown0 + j % kRing], (size_t)m * row_len * 4, cudaMemcpyDeviceToHost,
                                    r.copy), "D2H walks (uint32)");
        else
            TRW_TRY(cudaMemcpyAsync(out + done * row_len, r.ptr[kBufOut0 + b], (size_t)m * row_len * 8, cudaMemcpyDeviceToHost, r.copy), "D2H walks");
        r.down_bytes += m * row_len * (packed ? 4 : 8);
        TRW_TRY(cudaEventRecord(r.landed[(size_t)c], r.copy), "record landed");
        return TRW_OK;
    };
    // The calling thread keeps two copies queued (a chunk's format is decided as late as that allows), announces packed
    // chunks to the team as they land and feeds the copy engine's checksum pieces.  It never blocks inside the driver:
    // while it waits it widens.
    int enq = 0, pub = 0, landed_upto = 0, n_packed = 0;
    const bool timing = getenv("TRW_HOST_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double t_enqueued = 0, t_landed = 0;
    while (enq < n_chunks || pub < team.jobs_added()) {
        bool moved = false;
        rc = feed_check();
        if (rc) return rc;
        while (landed_upto < enq) {  // chunks land in order
            const cudaError_t st = cudaEventQuery(r.landed[(size_t)landed_upto]);
            if (st == cudaErrorNotReady) break;
            if (st != cudaSuccess) return check_cuda(st, "query landed");
            if (job_of_chunk[(size_t)landed_upto] >= 0) team.publish(++pub);
            if (++landed_upto == n_chunks) t_landed = ms_now();
            team.set_progress(landed_upto, n_chunks);
            moved = true;
        }
        if (enq < n_chunks && enq - landed_upto < 2) {
            const int j_next = team.jobs_added();
            const bool staging_free = team.job_finished(j_next - kRing);
            bool packed = false, can = true;
            if (adaptive) packed = staging_free && team.backlog() <= 1;
            else if (packed_by_pattern(enq)) { packed = true; can = staging_free; }
            if (can) {
                rc = enqueue(enq, packed);
                if (rc) return rc;
                n_packed += packed ? 1 : 0;
                if (++enq == n_chunks) t_enqueued = ms_now();
                moved = true;
            }
        }
        if (!moved && !team.help_widen()) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    team.set_progress(1, 1);
    while (dma_issued < dma_max) {  // (only if the download ended first: the engine's pieces run on, the team takes the rest)
        rc = feed_check();
        if (rc) return rc;
        if (dma_issued < dma_max && !team.help_widen()) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    const uint64_t sum = team.drain();
    const double t_drained = ms_now();
    TRW_TRY(cudaStreamSynchronize(r.compute), "sync compute");
    TRW_TRY(cudaStreamSynchronize(r.copy), "sync copy");
    if (check) {
        check->sum = sum;
        if (d_sum) {
            uint64_t* cell = (uint64_t*)r.pinned[kPinUp0] + 1;
            TRW_TRY(cudaMemcpyAsync(cell, d_sum, sizeof(uint64_t), cudaMemcpyDeviceToHost, r.check), "D2H check sum");
            TRW_TRY(cudaStreamSynchronize(r.check), "sync content check");
            check->sum += *cell;
        }
    }
    if (timing)
        fprintf(stderr, "[host_pipeline] %d chunks (%d packed%s), %d host threads, %lld sum pieces (%d by the copy engine) | last chunk enqueued %.1f ms, "
                "landed %.1f, host work done %.1f, all done %.1f\n", n_chunks, n_packed, adaptive ? ", adaptive" : "", n_threads,
                (long long)n_sum_pieces, dma_issued, t_enqueued, t_landed, t_drained, ms_now());
    if (n_packed) {
        int overflow = 0;
        TRW_TRY(cudaMemcpy(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost), "read overflow flag");
        if (overflow) return kRetryPlain;  // an id beyond 32 bits came out of the graph: the caller walks again with plain copies
    }
    return TRW_OK;
}
// The content check is complete after the first pass whatever its outcome, so a retry carries none.
static int host_pipeline_any(HostWalkCache& r, int d, const CsrWalkPlan& plan, const int64_t* d_targets, const HostCallShape& sh,
                             int64_t* out, int mode, int n_threads, ContentCheck* check = nullptr) {
    int rc = host_pipeline(r, d, plan, d_targets, sh, out, mode, n_threads, check);
    if (rc == kRetryPlain) rc = host_pipeline(r, d, plan, d_targets, sh, out, 0, n_threads);
    return rc;
}

// On any failure nothing may still be writing into the caller's buffers when the call returns.
struct StreamDrain {
    HostWalkCache* c;
    ~StreamDrain() {
        if (c->compute) cudaStreamSynchronize(c->compute);
        if (c->copy) cudaStreamSynchronize(c->copy);
        if (c->check) cudaStreamSynchronize(c->check);
        cudaGetLastError();
    }
};

static int ensure_streams(HostWalkCache& r) {
    if (r.compute) return TRW_OK;
    TRW_TRY(cudaStreamCreateWithFlags(&r.compute, cudaStreamNonBlocking), "stream create");
    TRW_TRY(cudaStreamCreateWithFlags(&r.copy, cudaStreamNonBlocking), "stream create");
    TRW_TRY(cudaStreamCreateWithFlags(&r.check, cudaStreamNonBlocking), "stream create");
    for (int k = 0; k < 2; ++k) {
        TRW_TRY(cudaEventCreateWithFlags(&r.walked[k], cudaEventDisableTiming), "event create");
        TRW_TRY(cudaEventCreateWithFlags(&r.copied[k], cudaEventDisableTiming), "event create");
        TRW_TRY(cudaEventCreateWithFlags(&r.uploaded[k], cudaEventDisableTiming), "event create");
    }
    return TRW_OK;
}

// Download mode for this call: how many of every 8 chunks travel packed, or -1 for "decided chunk by chunk by how the
// host keeps up" (host_pipeline).  Option host_compress 0: plain copies; ids must fit 32 bits.  Option host_packed_share
// fixes the share (measurements, tests); the default is the adaptive mode whenever the host path has threads to widen
// with.  (c3, 16 host threads, kept replica with its content check, profiles/r02_host_path.md: fixed shares measured
// between 84 and 122 ms per call depending on share AND box; no fixed share is best on every box.)
static int download_mode(int n_threads, int64_t n_nodes, bool ids_fit, bool checking = false) {
    (void)checking;
    const int64_t want = options().host_compress;
    if (want == 0 || !ids_fit || (uint64_t)n_nodes >= 0xFFFFFFFFull) return 0;
    const int share = (int)options().host_packed_share;  // of 8 chunks, how many travel packed (-1: adaptive)
    if (share >= 0) return share > 8 ? 8 : share;
    if (want == 2) return n_threads >= 4 ? 4 : 0;
    // The packed form trades PCIe bytes for host memory traffic (16 bytes per entry through the host's memory system against
    // 8).  That pays while a rank's own PCIe link is what limits it: one or two ranks on a host (c3: 111 -> 85-105 ms at one
    // rank, 84.8 -> 69.3 ms at two).  With four or eight ranks the host's aggregate device->host ceiling and its memory
    // bandwidth are the limit, every rank's team hammers the same DIMMs, and plain copies win (four ranks: 45 ms plain, 83 ms
    // adaptive; eight: 62 vs 65 ms) -- and no rank can see that from its own backlog.
    if (local_rank_count() > 2) return 0;
    return n_threads >= 2 ? -1 : 0;
}

}  // namespace trw

using namespace trw;

extern "C" int trw_walk_csr_host(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                                 const int64_t* targets, int64_t n_walks, int64_t walk_id_offset, double p, double q,
                                 int walk_length, int64_t seed, int64_t* out, int device) {
    if (n_walks < 0 || n_nodes < 0 || nnz < 0 || walk_length < 0) {
        set_error("trw_walk_csr_host: negative size");
        return TRW_ERR_ARG;
    }
    if (n_walks > 0 && (!row_ptr || !targets || !out || (nnz > 0 && !col_idx))) {
        set_error("trw_walk_csr_host: null pointer");
        return TRW_ERR_ARG;
    }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_csr_host: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_csr_host: p and q must be positive"); return TRW_ERR_ARG; }

    // TRW_HOST_TIMING=1 prints where an end-to-end call spends its time.
    const bool timing = getenv("TRW_HOST_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    const auto t_start = now();

    HostWalkCache local;  // used (and released on return) when caching is off or the ordinal is unusual
    const bool cached = options().host_cache_buffers != 0 && d < 64;
    HostWalkCache& r = cached ? g_host_cache[d] : local;
    std::lock_guard<std::mutex> lock(r.mu);
    struct Releaser {
        HostWalkCache* c;
        ~Releaser() { if (c) c->release(); }
    } releaser{cached ? nullptr : &local};
    int rc = ensure_streams(r);
    if (rc) return rc;
    StreamDrain drain{&r};

    const int n_threads = host_thread_count();
    const HostCallShape shape{n_walks, walk_id_offset, 0, 0, walk_length};
    bool uniform, want_table, want_strict, want_records;
    csr_one_shot_needs(p, q, nnz, n_walks, walk_length, &uniform, &want_table, &want_strict, &want_records);

    rc = r.reserve(kBufTargets, (size_t)n_walks * 8, "cudaMalloc targets");
    if (rc) return rc;
    TRW_TRY(cudaMemcpyAsync(r.ptr[kBufTargets], targets, (size_t)n_walks * 8, cudaMemcpyHostToDevice, r.compute), "H2D targets");
    r.up_bytes = n_walks * 8;
    r.down_bytes = 0;
    std::atomic<int> wide_targets{0};  // start nodes travel as int64, but their values come back inside the walks
    if (options().host_compress != 0 && n_threads >= 2)  // (a wide id that slips through is caught on the device: kRetryPlain)
        parallel_for(n_threads, [&](int tid, int nt) {
            bool bad = false;
            for (int64_t i = n_walks * tid / nt, e = n_walks * (tid + 1) / nt; i < e; ++i) bad |= (uint64_t)targets[i] > 0xFFFFFFFFull;
            if (bad) wide_targets.store(1, std::memory_order_relaxed);
        });

    // ---- the kept replica: same host arrays as last time?  Then walk on it at once and check its content meanwhile.
    const bool keep = cached && options().host_keep_graph != 0;
    if (keep && r.have_replica && r.key_row_ptr == row_ptr && r.key_col_idx == col_idx && r.key_n_nodes == n_nodes && r.key_nnz == nnz) {
        const auto t_hit = now();
        ++r.hits;
        // grow the preparation with use: the full kept form on the first hit, the triangle Blooms two hits later
        if (r.level < 1) {
            const size_t ws_bytes = csr_workspace_layout(n_nodes, nnz, false, options().records != 0).total;
            rc = r.reserve(kBufWorkspace, ws_bytes, "cudaMalloc workspace");
            if (!rc) rc = csr_graph_prepare(&r.graph, (const int64_t*)r.ptr[kBufRowPtr], (const int64_t*)r.ptr[kBufColIdx], n_nodes, nnz,
                                            /*uniform=*/false, /*want_table=*/true, /*want_strict=*/true, options().records != 0,
                                            r.ptr[kBufWorkspace], ws_bytes, d, r.compute, /*bloom_cap=*/0);
            if (rc) { r.forget_replica(); return rc; }
            r.level = 1;
        } else if (r.level == 1 && r.hits >= 3 && !uniform && options().edge_bloom_cap > 0) {
            rc = csr_add_blooms(&r.graph.prepared, (const int64_t*)r.ptr[kBufColIdx], n_nodes, nnz, options().edge_bloom_cap, d, r.compute);
            if (rc) { r.forget_replica(); return rc; }
            r.level = 2;
        }
        CsrWalkPlan plan;
        rc = csr_walk_plan(&plan, r.graph, p, q, walk_length, seed);
        if (rc) return rc;
        // The content check runs beside the pipeline, shared between the copy engine (pinned arrays only: it re-reads
        // pieces over the upload direction and the device sums them) and the host team -- see HostTeam.
        ContentCheck check;
        if (nnz > 0) check.sums[check.n_sums++] = HostSum{col_idx, 0, nnz, kChecksumColGolden};
        check.sums[check.n_sums++] = HostSum{row_ptr, 0, n_nodes + 1, kChecksumRowGolden};
        if (host_pinned(row_ptr) && host_pinned(col_idx)) check.dma_eighths = (int)std::min<int64_t>(std::max<int64_t>(options().host_check_dma, 0), 8);
        const int mode = download_mode(n_threads, n_nodes, wide_targets.load() == 0, /*checking=*/true);
        rc = host_pipeline_any(r, d, plan, (const int64_t*)r.ptr[kBufTargets], shape, out, mode, n_threads, &check);
        if (rc) return rc;
        const uint64_t host_sum = check.sum;
        if (host_sum == r.replica_checksum) {
            if (timing)
                fprintf(stderr, "[trw_walk_csr_host] kept replica (level %d, hit %d), content check by %s, %d of 8 chunks packed, %d host threads | total %.1f ms\n",
                        r.level, r.hits, check.dma_eighths > 0 ? "copy engine and host threads" : "host threads", mode, n_threads, ms_since(t_hit));
            r.last_call = 1;
            return TRW_OK;
        }
        r.forget_replica();  // the arrays changed under the same pointers: upload afresh and walk again
        r.last_call = 3;
    } else {
        r.last_call = 2;
    }

    // ---- fresh upload
    r.forget_replica();
    const size_t ws_bytes = csr_workspace_layout(n_nodes, nnz, uniform, want_records).total;
    // Wire compression: ids and CSR entries must fit 32 bits (checked value by value on the way up) and the
    // host needs threads to spare (kMinCompressThreads).
    bool compress = options().host_compress != 0 && n_threads >= kMinCompressThreads && (uint64_t)n_nodes < 0xFFFFFFFFull && nnz > 0;
    const int64_t up_chunk = std::max<int64_t>(1 << 16, options().host_up_chunk);
    rc = r.reserve(kBufRowPtr, (size_t)(n_nodes + 1) * 8, "cudaMalloc row_ptr");
    if (!rc) rc = r.reserve(kBufColIdx, (size_t)nnz * 8, "cudaMalloc col_idx");
    if (!rc && ws_bytes) rc = r.reserve(kBufWorkspace, ws_bytes, "cudaMalloc workspace");
    if (compress) {
        const size_t up_bytes = (size_t)std::min<int64_t>(up_chunk, nnz) * 4;
        for (int k = 0; k < 2 && !rc; ++k) {
            rc = r.reserve(kBufUp0 + k, up_bytes, "cudaMalloc upload staging");
            if (!rc) rc = r.reserve_pinned(kPinUp0 + k, up_bytes, "cudaHostAlloc upload staging");
        }
    }
    if (rc) return rc;
    void* const d_row_ptr = r.ptr[kBufRowPtr];
    void* const d_col_idx = r.ptr[kBufColIdx];
    void* const d_workspace = ws_bytes ? r.ptr[kBufWorkspace] : nullptr;

    const double ms_alloc = ms_since(t_start);
    const auto t_up = now();
    TRW_TRY(cudaMemcpyAsync(d_row_ptr, row_ptr, (size_t)(n_nodes + 1) * 8, cudaMemcpyHostToDevice, r.compute), "H2D row_ptr");
    r.up_bytes += (n_nodes + 1) * 8 + nnz * (compress ? 4 : 8);
    if (compress) {
        std::atomic<int> wide{0};
        int64_t sent = 0;
        for (int k = 0; sent < nnz; ++k) {
            const int b = k & 1;
            const int64_t m = std::min(up_chunk, nnz - sent);
            if (k >= 2) TRW_TRY(cudaEventSynchronize(r.uploaded[b]), "wait upload staging");
            uint32_t* stage = (uint32_t*)r.pinned[kPinUp0 + b];
            const int64_t* src = col_idx + sent;
            parallel_for(n_threads, [&](int tid, int nt) {
                const int64_t lo = m * tid / nt, hi = m * (tid + 1) / nt;
                if (!narrow_to_u32(src + lo, stage + lo, hi - lo)) wide.store(1, std::memory_order_relaxed);
            });
            if (wide.load()) break;
            TRW_TRY(cudaMemcpyAsync(r.ptr[kBufUp0 + b], stage, (size_t)m * 4, cudaMemcpyHostToDevice, r.compute), "H2D col_idx (uint32)");
            TRW_TRY(cudaEventRecord(r.uploaded[b], r.compute), "record uploaded");
            widen_u32_kernel<<<sm_count(d) * 8, 256, 0, r.compute>>>((const uint32_t*)r.ptr[kBufUp0 + b], (int64_t*)d_col_idx + sent, m);
            count_launch(1);
            sent += m;
        }
        TRW_TRY(cudaGetLastError(), "widen launch");
        if (wide.load()) {  // an id that needs more than 32 bits: plain copies for the whole call
            compress = false;
            r.up_bytes += nnz * 8;  // (col_idx goes up a second time, as int64)
            TRW_TRY(cudaStreamSynchronize(r.compute), "sync before the uncompressed upload");
        }
    }
    if (!compress && nnz) TRW_TRY(cudaMemcpyAsync(d_col_idx, col_idx, (size_t)nnz * 8, cudaMemcpyHostToDevice, r.compute), "H2D col_idx");

    double ms_upload = 0.0;
    if (timing) {
        TRW_TRY(cudaStreamSynchronize(r.compute), "sync uploads");
        ms_upload = ms_since(t_up);
    }
    const auto t_walk = now();
    rc = csr_graph_prepare(&r.graph, (const int64_t*)d_row_ptr, (const int64_t*)d_col_idx, n_nodes, nnz, uniform, want_table,
                           want_strict, want_records, d_workspace, ws_bytes, d, r.compute);
    if (rc) return rc;
    if (keep) {
        // the replica's own checksum (device side: it is what the host arrays will be compared with next time)
        rc = r.reserve(kBufUp0, 256, "cudaMalloc checksum cell");
        if (!rc) rc = r.reserve_pinned(kPinUp0, 256, "cudaHostAlloc checksum cell");
        if (rc) return rc;
        rc = trw_csr_checksum((const int64_t*)d_row_ptr, (const int64_t*)d_col_idx, n_nodes, nnz, (uint64_t*)r.ptr[kBufUp0], d, r.compute);
        if (rc) return rc;
        TRW_TRY(cudaMemcpyAsync(r.pinned[kPinUp0], r.ptr[kBufUp0], 8, cudaMemcpyDeviceToHost, r.compute), "D2H checksum");
    }
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, r.graph, p, q, walk_length, seed);
    if (rc) return rc;
    const int mode = download_mode(n_threads, n_nodes, compress && wide_targets.load() == 0);
    rc = host_pipeline_any(r, d, plan, (const int64_t*)r.ptr[kBufTargets], shape, out, mode, n_threads);
    if (rc) return rc;
    if (keep) {
        r.replica_checksum = *(const uint64_t*)r.pinned[kPinUp0];  // the pipeline has synchronised the stream
        r.key_row_ptr = row_ptr; r.key_col_idx = col_idx; r.key_n_nodes = n_nodes; r.key_nnz = nnz;
        r.have_replica = true;
        r.level = 0;
        r.hits = 0;
    }
    if (timing) {
        const double h2d_gb = ((double)(n_nodes + 1) + (double)nnz * (compress ? 0.5 : 1.0) + (double)n_walks) * 8 / 1e9;
        fprintf(stderr, "[trw_walk_csr_host] fresh upload, %s up, download mode %d, %d host threads | alloc %.1f ms | upload %.1f ms "
                        "(%.2f GB, %.1f GB/s) | prepare+walk+download %.1f ms | total %.1f ms\n",
                compress ? "uint32" : "int64", mode, n_threads, ms_alloc, ms_upload, h2d_gb,
                ms_upload > 0 ? h2d_gb / (ms_upload / 1e3) : 0.0, ms_since(t_walk), ms_since(t_start));
    }
    return TRW_OK;
}

// The download half on its own: the graph is already on the device and prepared (trw_csr_graph_prepare), the start
// nodes and the walks live in host memory.  This is what each rank of a multi-GPU job runs on its shard.
extern "C" int trw_walk_csr_to_host(const trw_csr_graph_view* view, const int64_t* targets, int64_t n_walks,
                                    int64_t walk_id_offset, int64_t walk_id_block, int64_t walk_id_stride, double p, double q,
                                    int walk_length, int64_t seed, int64_t* out) {
    if (!view || !view->graph) { set_error("trw_walk_csr_to_host: null graph"); return TRW_ERR_ARG; }
    if (n_walks < 0 || walk_length < 0) { set_error("trw_walk_csr_to_host: negative size"); return TRW_ERR_ARG; }
    if (n_walks > 0 && (!targets || !out)) { set_error("trw_walk_csr_to_host: null pointer"); return TRW_ERR_ARG; }
    if (n_walks == 0) return TRW_OK;
    CsrGraph g;
    int rc = csr_graph_of_handle(view->graph, &g);
    if (rc) return rc;
    if (view->row_ptr) g.row_ptr.base = reinterpret_cast<const char*>(view->row_ptr);
    if (view->col_idx) g.col_idx.base = reinterpret_cast<const char*>(view->col_idx);
    const int d = g.device;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_csr_to_host: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    HostWalkCache local;
    const bool cached = options().host_cache_buffers != 0 && d < 64;
    HostWalkCache& r = cached ? g_host_cache[d] : local;
    std::lock_guard<std::mutex> lock(r.mu);
    struct Releaser {
        HostWalkCache* c;
        ~Releaser() { if (c) c->release(); }
    } releaser{cached ? nullptr : &local};
    rc = ensure_streams(r);
    if (rc) return rc;
    StreamDrain drain{&r};
    if (view->ready_stream) {  // order after whatever prepared the graph
        TRW_TRY(cudaEventRecord(r.uploaded[0], (cudaStream_t)view->ready_stream), "record ready");
        TRW_TRY(cudaStreamWaitEvent(r.compute, r.uploaded[0], 0), "wait ready");
    }
    rc = r.reserve(kBufTargets, (size_t)n_walks * 8, "cudaMalloc targets");
    if (rc) return rc;
    TRW_TRY(cudaMemcpyAsync(r.ptr[kBufTargets], targets, (size_t)n_walks * 8, cudaMemcpyHostToDevice, r.compute), "H2D targets");
    r.up_bytes = n_walks * 8;
    r.down_bytes = 0;
    const int n_threads = host_thread_count();
    bool ids_fit = true;
    if (options().host_compress != 0 && n_threads >= 2) {
        std::atomic<int> wide{0};
        parallel_for(n_threads, [&](int tid, int nt) {
            bool bad = false;
            for (int64_t i = n_walks * tid / nt, e = n_walks * (tid + 1) / nt; i < e; ++i) bad |= (uint64_t)targets[i] > 0xFFFFFFFFull;
            if (bad) wide.store(1, std::memory_order_relaxed);
        });
        ids_fit = wide.load() == 0;
    }
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, g, p, q, walk_length, seed);
    if (rc) return rc;
    const HostCallShape shape{n_walks, walk_id_offset, walk_id_block, walk_id_stride, walk_length};
    return host_pipeline_any(r, d, plan, (const int64_t*)r.ptr[kBufTargets], shape, out, download_mode(n_threads, g.n_nodes, ids_fit), n_threads);
}

extern "C" void trw_release_cached_buffers(void) {
    int prev = -1;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; ++d) {
        HostWalkCache& c = g_host_cache[d];
        std::lock_guard<std::mutex> lock(c.mu);
        bool used = c.compute != nullptr;
        for (int k = 0; k < kNumBufs; ++k) used |= c.ptr[k] != nullptr;
        if (!used) continue;
        cudaSetDevice(d);
        c.release();
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

// Host twin of trw_csr_checksum: the same 64-bit sum over arrays in host memory, on `n_threads` threads (<= 0: the
// host path's thread count).  `simd` = 0 keeps to the scalar loop (what the tests compare the AVX-512 form against).
extern "C" int trw_csr_checksum_host(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz, int n_threads, int simd,
                                     uint64_t* out) {
    if (n_nodes < 0 || nnz < 0 || !out || (!row_ptr && n_nodes + 1 > 0) || (nnz > 0 && !col_idx)) {
        set_error("trw_csr_checksum_host: bad argument");
        return TRW_ERR_ARG;
    }
    const int nt = std::max(1, n_threads > 0 ? n_threads : host_thread_count());
    std::vector<uint64_t> part((size_t)nt, 0);
    parallel_for(nt, [&](int tid, int n) {
        const int64_t n_row = n_nodes + 1;
        part[(size_t)tid] = checksum_range(col_idx, nnz * tid / n, nnz * (tid + 1) / n, kChecksumColGolden, simd != 0) +
                            checksum_range(row_ptr, n_row * tid / n, n_row * (tid + 1) / n, kChecksumRowGolden, simd != 0);
    });
    uint64_t total = 0;
    for (uint64_t x : part) total += x;
    *out = total;
    return TRW_OK;
}

// The kept replica of trw_walk_csr_host on `device`: out[0] a replica is held, out[1] its preparation level (-1 none, 0
// one call's needs, 1 the full kept preparation, 2 with triangle Blooms), out[2] validated hits so far, out[3] how the
// last call got its graph (0 no call yet, 1 kept replica validated, 2 fresh upload, 3 kept replica found changed); with
// n_out >= 6 also out[4], out[5]: the bytes the last host-path call on the device (either entry) moved host->device and
// device->host.
extern "C" int trw_host_replica_info(int device, int64_t* out, int n_out) {
    const int d = resolve_device(device);
    if (d < 0 || d >= 64 || !out || n_out < 4) {
        set_error("trw_host_replica_info: bad argument");
        return TRW_ERR_ARG;
    }
    HostWalkCache& c = g_host_cache[d];
    std::lock_guard<std::mutex> lock(c.mu);
    out[0] = c.have_replica ? 1 : 0;
    out[1] = c.level;
    out[2] = c.hits;
    out[3] = c.last_call;
    if (n_out >= 6) { out[4] = c.up_bytes; out[5] = c.down_bytes; }
    return TRW_OK;
}
