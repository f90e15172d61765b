// trw_walk_csr_host: the CSR walk for callers whose tensors live in host memory.
//
// The reference has no such path (a CPU tensor simply ran csrc/cpu); a drop-in user who holds
// CPU tensors still has to get them to the GPU and the walks back, and that round trip is what
// an end-to-end measurement pays for.  So it is engineered rather than left to three blocking
// copies: the graph goes up once, the start nodes are walked in chunks on one stream, and each
// finished chunk is copied back on a second stream while the next chunk is being walked
// (PCIe is full duplex and the copy engines run beside the SMs).
//
// What is left is the wire: 4.4 GB up and 5.7 GB back per C3 call at ~55 GB/s.  Node ids and CSR
// entries fit 32 bits for every graph the fast paths take, and the host can narrow/widen at ~100 GB/s
// with its cores (tools/host_bandwidth_probe.py), so with enough threads col_idx crosses PCIe as uint32
// (narrowed chunk by chunk into pinned staging while the previous chunk is in flight, widened on the
// device) and the walks come back as uint32 (narrowed on the device, widened by the host threads into
// the caller's buffer while the next chunk is in flight).  Any id that does not fit falls back to the
// plain copies.
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
enum { kBufRowPtr = 0, kBufColIdx, kBufTargets, kBufWorkspace, kBufOut0, kBufOut1, kBufUp0, kBufUp1, kBufDown0, kBufDown1, kNumBufs };
enum { kPinUp0 = 0, kPinUp1, kPinDown0, kPinDown1, kNumPinned };
struct HostWalkCache {
    void* ptr[kNumBufs] = {};
    size_t cap[kNumBufs] = {};
    void* pinned[kNumPinned] = {};  // host staging of the compressed transfers
    size_t pinned_cap[kNumPinned] = {};
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t walked[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr}, uploaded[2] = {nullptr, nullptr};
    void release() {
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
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
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        compute = copy = nullptr;
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

// Measured on the GPU box (16 cores): 16 threads per rank turn 190 ms into 144 ms per C3 call; two ranks with 8
// threads each share the host's memory bandwidth and come out 3 % behind the plain copies.
constexpr int kMinCompressThreads = 12;

// Host threads available to this process for the wire compression: the machine's, shared among the ranks of
// a torchrun launch on this node (LOCAL_WORLD_SIZE), or option host_threads.
static int host_thread_count() {
    if (options().host_threads > 0) return (int)options().host_threads;
    int n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    const int ranks = lws ? atoi(lws) : 1;
    if (ranks > 1) n /= ranks;
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
static std::mutex g_host_mutex;

#define TRW_TRY(expr, what)                       \
    do {                                          \
        int rc__ = check_cuda((expr), what);      \
        if (rc__) return rc__;                    \
    } while (0)

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

    const int64_t row_len = (int64_t)walk_length + 1;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(options().host_chunk_walks, n_walks));
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_csr_host: p and q must be positive"); return TRW_ERR_ARG; }
    bool uniform, want_table, want_strict, want_records;
    csr_one_shot_needs(p, q, nnz, n_walks, walk_length, &uniform, &want_table, &want_strict, &want_records);
    const size_t ws_bytes = csr_workspace_layout(n_nodes, nnz, uniform, want_records).total;

    // TRW_HOST_TIMING=1 prints where an end-to-end call spends its time (adds one stream sync after the uploads).
    const bool timing = getenv("TRW_HOST_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    const auto t_start = now();

    std::lock_guard<std::mutex> lock(g_host_mutex);
    HostWalkCache local;  // used (and released on return) when caching is off or the ordinal is unusual
    const bool cached = options().host_cache_buffers != 0 && d < 64;
    HostWalkCache& r = cached ? g_host_cache[d] : local;
    struct Releaser {
        HostWalkCache* c;
        ~Releaser() { if (c) c->release(); }
    } releaser{cached ? nullptr : &local};
    if (!r.compute) {
        TRW_TRY(cudaStreamCreateWithFlags(&r.compute, cudaStreamNonBlocking), "stream create");
        TRW_TRY(cudaStreamCreateWithFlags(&r.copy, cudaStreamNonBlocking), "stream create");
        for (int k = 0; k < 2; ++k) {
            TRW_TRY(cudaEventCreateWithFlags(&r.walked[k], cudaEventDisableTiming), "event create");
            TRW_TRY(cudaEventCreateWithFlags(&r.copied[k], cudaEventDisableTiming), "event create");
            TRW_TRY(cudaEventCreateWithFlags(&r.uploaded[k], cudaEventDisableTiming), "event create");
        }
    }
    // Wire compression: ids and CSR entries must fit 32 bits (checked value by value on the way up) and the
    // host needs threads to spare (kMinCompressThreads).
    const int n_threads = host_thread_count();
    bool compress = options().host_compress != 0 && n_threads >= kMinCompressThreads && (uint64_t)n_nodes < 0xFFFFFFFFull && nnz > 0;
    const int64_t up_chunk = std::max<int64_t>(1 << 16, options().host_up_chunk);

    const int n_buf = n_walks > chunk ? 2 : 1;
    int rc = r.reserve(kBufRowPtr, (size_t)(n_nodes + 1) * 8, "cudaMalloc row_ptr");
    if (!rc) rc = r.reserve(kBufColIdx, (size_t)nnz * 8, "cudaMalloc col_idx");
    if (!rc) rc = r.reserve(kBufTargets, (size_t)n_walks * 8, "cudaMalloc targets");
    if (!rc && ws_bytes) rc = r.reserve(kBufWorkspace, ws_bytes, "cudaMalloc workspace");
    for (int k = 0; k < n_buf && !rc; ++k) rc = r.reserve(kBufOut0 + k, (size_t)chunk * row_len * 8, "cudaMalloc walks");
    if (compress) {
        const size_t up_bytes = (size_t)std::min<int64_t>(up_chunk, nnz) * 4, down_bytes = (size_t)chunk * row_len * 4;
        for (int k = 0; k < 2 && !rc; ++k) {
            rc = r.reserve(kBufUp0 + k, up_bytes, "cudaMalloc upload staging");
            if (!rc) rc = r.reserve_pinned(kPinUp0 + k, up_bytes, "cudaHostAlloc upload staging");
        }
        for (int k = 0; k < n_buf && !rc; ++k) {
            rc = r.reserve(kBufDown0 + k, down_bytes + 256, "cudaMalloc download staging");
            if (!rc) rc = r.reserve_pinned(kPinDown0 + k, down_bytes, "cudaHostAlloc download staging");
        }
    }
    if (rc) return rc;
    void* const d_row_ptr = r.ptr[kBufRowPtr];
    void* const d_col_idx = r.ptr[kBufColIdx];
    void* const d_targets = r.ptr[kBufTargets];
    void* const d_workspace = ws_bytes ? r.ptr[kBufWorkspace] : nullptr;
    void* const d_out[2] = {r.ptr[kBufOut0], r.ptr[kBufOut1]};

    const double ms_alloc = ms_since(t_start);
    const auto t_up = now();
    TRW_TRY(cudaMemcpyAsync(d_row_ptr, row_ptr, (size_t)(n_nodes + 1) * 8, cudaMemcpyHostToDevice, r.compute), "H2D row_ptr");
    TRW_TRY(cudaMemcpyAsync(d_targets, targets, (size_t)n_walks * 8, cudaMemcpyHostToDevice, r.compute), "H2D targets");
    if (compress) {
        // start nodes travel as int64, but their values come back inside the uint32 walks
        std::atomic<int> wide{0};
        parallel_for(n_threads, [&](int tid, int nt) {
            bool bad = false;
            for (int64_t i = n_walks * tid / nt, e = n_walks * (tid + 1) / nt; i < e; ++i) bad |= (uint64_t)targets[i] > 0xFFFFFFFFull;
            if (bad) wide.store(1, std::memory_order_relaxed);
        });
        int64_t sent = 0;
        for (int k = 0; sent < nnz && !wide.load(); ++k) {
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
    CsrGraph graph;
    rc = csr_graph_prepare(&graph, (const int64_t*)d_row_ptr, (const int64_t*)d_col_idx, n_nodes, nnz, uniform, want_table,
                           want_strict, want_records, d_workspace, ws_bytes, d, r.compute);
    if (rc) return rc;
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, graph, p, q, walk_length, seed);
    if (rc) return rc;

    int* d_overflow = compress ? (int*)((char*)r.ptr[kBufDown0] + (size_t)chunk * row_len * 4) : nullptr;  // the 256 spare bytes
    if (compress) TRW_TRY(cudaMemsetAsync(d_overflow, 0, sizeof(int), r.compute), "overflow flag memset");
    // widen chunk c (already in pinned staging) into the caller's buffer with the host threads
    auto widen_to_caller = [&](int b, int64_t first_walk, int64_t m) {
        const uint32_t* stage = (const uint32_t*)r.pinned[kPinDown0 + b];
        int64_t* dst = out + first_walk * row_len;
        const int64_t n_el = m * row_len;
        parallel_for(n_threads, [&](int tid, int nt) {
            const int64_t lo = n_el * tid / nt, hi = n_el * (tid + 1) / nt;
            widen_to_i64(stage + lo, dst + lo, hi - lo);
        });
    };
    int64_t done = 0, prev_first = 0, prev_m = 0;
    int prev_b = -1;
    for (int c = 0; done < n_walks; ++c) {
        const int b = c & 1;
        const int64_t m = std::min(chunk, n_walks - done);
        if (c >= 2) TRW_TRY(cudaStreamWaitEvent(r.compute, r.copied[b], 0), "wait copied");
        rc = csr_walk_launch(plan, (const int64_t*)d_targets + done, m, walk_id_offset + done, (int64_t*)d_out[b],
                             row_len, r.compute);
        if (rc) return rc;
        if (compress) {
            narrow_i64_kernel<<<sm_count(d) * 8, 256, 0, r.compute>>>((const int64_t*)d_out[b], (uint32_t*)r.ptr[kBufDown0 + b],
                                                                     m * row_len, d_overflow);
            count_launch(1);
        }
        TRW_TRY(cudaEventRecord(r.walked[b], r.compute), "record walked");
        TRW_TRY(cudaStreamWaitEvent(r.copy, r.walked[b], 0), "wait walked");
        if (compress)
            TRW_TRY(cudaMemcpyAsync(r.pinned[kPinDown0 + b], r.ptr[kBufDown0 + b], (size_t)m * row_len * 4, cudaMemcpyDeviceToHost,
                                    r.copy), "D2H walks (uint32)");
        else
            TRW_TRY(cudaMemcpyAsync(out + done * row_len, d_out[b], (size_t)m * row_len * 8, cudaMemcpyDeviceToHost, r.copy),
                    "D2H walks");
        TRW_TRY(cudaEventRecord(r.copied[b], r.copy), "record copied");
        if (compress && prev_b >= 0) {  // the previous chunk has landed (or lands now): widen it while this one walks and copies
            TRW_TRY(cudaEventSynchronize(r.copied[prev_b]), "wait previous chunk");
            widen_to_caller(prev_b, prev_first, prev_m);
        }
        prev_b = b; prev_first = done; prev_m = m;
        done += m;
    }
    TRW_TRY(cudaStreamSynchronize(r.compute), "sync compute");
    TRW_TRY(cudaStreamSynchronize(r.copy), "sync copy");
    if (compress) {
        if (prev_b >= 0) widen_to_caller(prev_b, prev_first, prev_m);
        int overflow = 0;
        TRW_TRY(cudaMemcpy(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost), "read overflow flag");
        if (overflow) { set_error("trw_walk_csr_host: a walk entry did not fit the 32-bit wire format (internal error)"); return TRW_ERR_CUDA; }
    }
    if (timing) {
        const double h2d_gb = ((double)(n_nodes + 1) + (double)nnz * (compress ? 0.5 : 1.0) + (double)n_walks) * 8 / 1e9;
        const double d2h_gb = (double)n_walks * row_len * (compress ? 4 : 8) / 1e9;
        fprintf(stderr, "[trw_walk_csr_host] %s, %d host threads | alloc %.1f ms | upload %.1f ms (%.2f GB, %.1f GB/s) | "
                        "walk+download %.1f ms (%.2f GB back) | total %.1f ms\n",
                compress ? "uint32 wire format" : "int64 copies", n_threads, ms_alloc, ms_upload, h2d_gb,
                ms_upload > 0 ? h2d_gb / (ms_upload / 1e3) : 0.0, ms_since(t_walk), d2h_gb, ms_since(t_start));
    }
    return TRW_OK;
}

extern "C" void trw_release_cached_buffers(void) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    int prev = -1;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; ++d) {
        HostWalkCache& c = g_host_cache[d];
        bool used = c.compute != nullptr;
        for (int k = 0; k < kNumBufs; ++k) used |= c.ptr[k] != nullptr;
        if (!used) continue;
        cudaSetDevice(d);
        c.release();
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}
