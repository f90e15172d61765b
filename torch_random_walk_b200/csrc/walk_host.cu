// trw_walk_csr_host: the CSR walk for callers whose tensors live in host memory.
//
// The reference has no such path (a CPU tensor simply ran csrc/cpu); a drop-in user who holds
// CPU tensors still has to get them to the GPU and the walks back, and that round trip is what
// an end-to-end measurement pays for.  So it is engineered rather than left to three blocking
// copies: the graph goes up once, the start nodes are walked in chunks on one stream, and each
// finished chunk is copied back on a second stream while the next chunk is being walked
// (PCIe is full duplex and the copy engines run beside the SMs).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "trw_common.cuh"
#include "trw_options.h"
#include "walk_csr.h"

namespace trw {

// Device buffers, streams and events of the host path are kept between calls (per device,
// grow-only): cudaMalloc/cudaFree of ~15 GB cost more than the walk itself.  Released by
// trw_release_cached_buffers() or with option host_cache_buffers = 0.
enum { kBufRowPtr = 0, kBufColIdx, kBufTargets, kBufWorkspace, kBufOut0, kBufOut1, kNumBufs };
struct HostWalkCache {
    void* ptr[kNumBufs] = {};
    size_t cap[kNumBufs] = {};
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t walked[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    void release() {
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
        for (int k = 0; k < kNumBufs; ++k) {
            if (ptr[k]) cudaFree(ptr[k]);
            ptr[k] = nullptr;
            cap[k] = 0;
        }
        for (int k = 0; k < 2; ++k) {
            if (walked[k]) cudaEventDestroy(walked[k]);
            if (copied[k]) cudaEventDestroy(copied[k]);
            walked[k] = copied[k] = nullptr;
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
};
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
        }
    }
    const int n_buf = n_walks > chunk ? 2 : 1;
    int rc = r.reserve(kBufRowPtr, (size_t)(n_nodes + 1) * 8, "cudaMalloc row_ptr");
    if (!rc) rc = r.reserve(kBufColIdx, (size_t)nnz * 8, "cudaMalloc col_idx");
    if (!rc) rc = r.reserve(kBufTargets, (size_t)n_walks * 8, "cudaMalloc targets");
    if (!rc && ws_bytes) rc = r.reserve(kBufWorkspace, ws_bytes, "cudaMalloc workspace");
    for (int k = 0; k < n_buf && !rc; ++k) rc = r.reserve(kBufOut0 + k, (size_t)chunk * row_len * 8, "cudaMalloc walks");
    if (rc) return rc;
    void* const d_row_ptr = r.ptr[kBufRowPtr];
    void* const d_col_idx = r.ptr[kBufColIdx];
    void* const d_targets = r.ptr[kBufTargets];
    void* const d_workspace = ws_bytes ? r.ptr[kBufWorkspace] : nullptr;
    void* const d_out[2] = {r.ptr[kBufOut0], r.ptr[kBufOut1]};

    const double ms_alloc = ms_since(t_start);
    const auto t_up = now();
    TRW_TRY(cudaMemcpyAsync(d_row_ptr, row_ptr, (size_t)(n_nodes + 1) * 8, cudaMemcpyHostToDevice, r.compute), "H2D row_ptr");
    if (nnz) TRW_TRY(cudaMemcpyAsync(d_col_idx, col_idx, (size_t)nnz * 8, cudaMemcpyHostToDevice, r.compute), "H2D col_idx");
    TRW_TRY(cudaMemcpyAsync(d_targets, targets, (size_t)n_walks * 8, cudaMemcpyHostToDevice, r.compute), "H2D targets");

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

    int64_t done = 0;
    for (int c = 0; done < n_walks; ++c) {
        const int b = c & 1;
        const int64_t m = std::min(chunk, n_walks - done);
        if (c >= 2) TRW_TRY(cudaStreamWaitEvent(r.compute, r.copied[b], 0), "wait copied");
        rc = csr_walk_launch(plan, (const int64_t*)d_targets + done, m, walk_id_offset + done, (int64_t*)d_out[b],
                             row_len, r.compute);
        if (rc) return rc;
        TRW_TRY(cudaEventRecord(r.walked[b], r.compute), "record walked");
        TRW_TRY(cudaStreamWaitEvent(r.copy, r.walked[b], 0), "wait walked");
        TRW_TRY(cudaMemcpyAsync(out + done * row_len, d_out[b], (size_t)m * row_len * 8, cudaMemcpyDeviceToHost, r.copy),
                "D2H walks");
        TRW_TRY(cudaEventRecord(r.copied[b], r.copy), "record copied");
        done += m;
    }
    TRW_TRY(cudaStreamSynchronize(r.compute), "sync compute");
    TRW_TRY(cudaStreamSynchronize(r.copy), "sync copy");
    if (timing) {
        const double h2d_gb = ((double)(n_nodes + 1) + (double)nnz + (double)n_walks) * 8 / 1e9;
        const double d2h_gb = (double)n_walks * row_len * 8 / 1e9;
        fprintf(stderr, "[trw_walk_csr_host] alloc %.1f ms | upload %.1f ms (%.2f GB, %.1f GB/s) | walk+download %.1f ms "
                        "(%.2f GB back) | total %.1f ms\n",
                ms_alloc, ms_upload, h2d_gb, ms_upload > 0 ? h2d_gb / (ms_upload / 1e3) : 0.0, ms_since(t_walk), d2h_gb,
                ms_since(t_start));
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
