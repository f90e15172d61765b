// CSR random walks for sm_100a: first-order (uniform) and second-order (node2vec) kernels.
//
// Replaces the reference's walk_gpu / uniform_walk_gpu / biased_walk_gpu
// (csrc/cuda/rw_cuda.cu:186-248 / 59-98 / 100-184).  Semantics kept:
//   * out[i,0] = targets[i]; every later entry is drawn from adj(previous entry);
//   * a node without out-edges keeps the walk where it is (rw_cuda.cu:25-30);
//   * node2vec: step 1 uniform, then propose x ~ U(adj(v)), accept with probability
//     prob_0 (x == t), prob_1 (x in adj(t)), prob_2 (otherwise), prob_k = {1/p,1,1/q}/max
//     (rw_cuda.cu:119-123, 146-179).
// Design: both kernels are latency-bound dependent gathers, so one
// thread owns one walk and the SM is kept full of them.  The node2vec loop is flattened to
// one *trial* per iteration so that lanes whose proposal was accepted move on to their next
// step instead of idling until the slowest lane of the warp is accepted.  "x in adj(t)" is
// answered with one 32-byte sector from a hashed copy of the adjacency built per call into
// caller-provided workspace (build_member_table), not by scanning or searching adj(t).
#include "trw_common.cuh"
#include "trw_options.h"
#include "walk_csr.h"

namespace trw {

// ------------------------------------------------------------------------------------------
// Membership table.  Row t of the CSR owns the bytes [8*row_ptr[t], 8*row_ptr[t+1]) of a
// table as large as col_idx; the 32-byte buckets wholly inside that span hold the row's
// neighbour ids as uint32 (8 slots per bucket, EMPTY = 0xFFFFFFFF), open addressing over
// buckets.  A row of degree d >= kMinTableDeg owns >= (d-6)/4 buckets = 2d-12 >= d slots, so
// inserts always find room; shorter rows are scanned directly (<= 15 ids, <= 5 sectors).
// No per-row pointer is needed: the bucket range follows from the row_ptr pair the walk
// already holds.  Lookup: hash -> bucket -> one sector; a hit, or any EMPTY slot (the bucket
// never overflowed), ends the probe.
// ------------------------------------------------------------------------------------------
constexpr int64_t kMinTableDeg = 16;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ void table_span(int64_t b, int64_t e, int64_t& first, int64_t& nb) {
    first = (b + 3) >> 2;
    nb = (e >> 2) - first;
}

constexpr int kBuildThreads = 256;
constexpr int kBuildPerThread = 8;
constexpr int kBuildTile = kBuildThreads * kBuildPerThread;

// Largest r with row_ptr[r] <= e (the non-empty row that holds CSR entry e).
__device__ __forceinline__ int64_t row_of_entry(const int64_t* __restrict__ row_ptr, int64_t n_nodes, int64_t e) {
    int64_t lo = 0, hi = n_nodes;  // invariant: row_ptr[lo] <= e < row_ptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kBuildThreads)
build_member_table_kernel(const int64_t* __restrict__ row_ptr, const int64_t* __restrict__ col_idx,
                          int64_t n_nodes, int64_t nnz, uint32_t* __restrict__ table) {
    __shared__ uint32_t head[kBuildTile];  // row offset (relative to r0) that starts at this entry
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    __shared__ int64_t s_r0, s_r1;

    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * kBuildTile;
    const int64_t e1 = min(e0 + (int64_t)kBuildTile, nnz);
    if (tid == 0) s_r0 = row_of_entry(row_ptr, n_nodes, e0);
    if (tid == 32) s_r1 = row_of_entry(row_ptr, n_nodes, e1 - 1);
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) head[k * kBuildThreads + tid] = 0;
    __syncthreads();
    const int64_t r0 = s_r0, r1 = s_r1;
    // Mark the first entry of every non-empty row that starts inside the tile.
    for (int64_t r = r0 + 1 + tid; r <= r1; r += kBuildThreads) {
        int64_t b = __ldg(row_ptr + r), e = __ldg(row_ptr + r + 1);
        if (e > b) head[b - e0] = (uint32_t)(r - r0);
    }
    __syncthreads();
    // Inclusive max-scan: entry j belongs to row r0 + max(head[0..j]).
    uint32_t own[kBuildPerThread];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        run = max(run, head[tid * kBuildPerThread + k]);
        own[k] = run;
    }
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((tid & 31) >= d) incl = max(incl, o);
    }
    if ((tid & 31) == 31) warp_max[tid >> 5] = incl;
    __syncthreads();
    uint32_t before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if ((tid & 31) == 0) before = 0;
    for (int w = 0; w < (tid >> 5); ++w) before = max(before, warp_max[w]);

    int64_t cur_row = -1, first = 0, nb = 0;
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        const int64_t e = e0 + tid * kBuildPerThread + k;
        if (e >= e1) break;
        const int64_t r = r0 + max(before, own[k]);
        if (r != cur_row) {
            cur_row = r;
            int64_t b = __ldg(row_ptr + r), en = __ldg(row_ptr + r + 1);
            if (en - b >= kMinTableDeg) table_span(b, en, first, nb); else nb = 0;
        }
        if (nb <= 0) continue;
        const uint32_t x = (uint32_t)ldg64_stream(col_idx + e);
        const uint32_t h = mix32(x);
        int64_t bkt = (int64_t)__umul64hi((uint64_t)h << 32, (uint64_t)nb);
        bool done = false;
        while (!done) {
            uint32_t* slots = table + (first + bkt) * 8;
            // Snapshot the bucket, then claim the first EMPTY slot seen; a lost race just moves on.
            uint4 lo4 = ld_relaxed_u32x4(slots);
            uint4 hi4 = ld_relaxed_u32x4(slots + 4);
            uint32_t snap[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
            for (int j = 0; j < 8 && !done; ++j) {
                if (snap[j] == x) done = true;
                else if (snap[j] == kEmpty) {
                    uint32_t old = atomicCAS(slots + j, kEmpty, x);
                    if (old == kEmpty || old == x) done = true;
                }
            }
            if (++bkt == nb) bkt = 0;
        }
    }
}

// x in adj(t)?  (b,e) = row_ptr[t], row_ptr[t+1].
template <bool TABLE>
__device__ __forceinline__ bool is_member(int64_t x, int64_t b, int64_t e, const int64_t* __restrict__ col_idx,
                                          const uint32_t* __restrict__ table) {
    if (TABLE && e - b >= kMinTableDeg) {
        int64_t first, nb;
        table_span(b, e, first, nb);
        const uint32_t x32 = (uint32_t)x;
        int64_t bkt = (int64_t)__umul64hi((uint64_t)mix32(x32) << 32, (uint64_t)nb);
        for (int64_t probes = 0; probes < nb; ++probes) {
            Sector64 s = ldg_sector(table + (first + bkt) * 8);
            uint32_t w[8] = {(uint32_t)s.a, (uint32_t)(s.a >> 32), (uint32_t)s.b, (uint32_t)(s.b >> 32),
                             (uint32_t)s.c, (uint32_t)(s.c >> 32), (uint32_t)s.d, (uint32_t)(s.d >> 32)};
            bool hit = false, open = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) { hit |= (w[j] == x32); open |= (w[j] == kEmpty); }
            if (hit) return true;
            if (open) return false;
            if (++bkt == nb) bkt = 0;
        }
        return false;
    }
    // Short (or table-less) row: the reference's scan, csrc/cuda/rw_cuda.cu:48-53.
    // Eight independent loads per round so the scan is not one dependent chain.
    for (int64_t i = b; i < e; i += 8) {
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i + j < e) found |= (ldg64_stream(col_idx + i + j) == x);
        if (found) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void load_row(const WalkArgs& a, int64_t v, int64_t& b, int64_t& e) {
    if ((uint64_t)v < (uint64_t)a.n_nodes) {
        b = __ldg(a.row_ptr + v);
        e = __ldg(a.row_ptr + v + 1);
    } else {
        b = e = 0;  // id outside the graph: treated as a node without out-edges
    }
}

// Neighbour of v at a uniformly random position, or v itself when it has none (rw_cuda.cu:8-31).
__device__ __forceinline__ int64_t pick_neighbor(const WalkArgs& a, int64_t v, int64_t b, int64_t e, uint32_t r0,
                                                 uint32_t r1) {
    const int64_t deg = e - b;
    if (deg <= 0) return v;
    const int64_t idx = b + bounded(r0, r1, deg);
    if ((uint64_t)idx >= (uint64_t)a.nnz) return v;
    return ldg64_stream(a.col_idx + idx);
}

template <int BLOCK, bool STAGE>
struct RowOut {
    RowStager<BLOCK> st;
    int64_t* row;
    __device__ __forceinline__ void init(int64_t (*ring)[BLOCK], int64_t* r, int tid) {
        row = r;
        if (STAGE) st.init(ring, r, tid);
    }
    __device__ __forceinline__ void put(int s, int64_t v, bool last) {
        if (STAGE) st.put(s, v, last); else row[s] = v;
    }
};

// First-order walk: one thread per walk, two dependent gathers per step (row_ptr pair, then
// the chosen col_idx entry), one Philox block per four steps.
template <int BLOCK, bool STAGE>
__global__ void __launch_bounds__(BLOCK) uniform_walk_kernel(const WalkArgs a) {
    __shared__ int64_t ring[STAGE ? 4 : 1][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowOut<BLOCK, STAGE> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x);

    int64_t v = __ldg(a.targets + i);
    const int L = a.walk_length;
    o.put(0, v, L == 0);
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int s = 1; s <= L; ++s) {
        const int k = (s - 1) & 3;
        if (k == 0)
            rnd = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)((s - 1) >> 2), 0x80000000u), a.key);
        const uint32_t r = rnd.x;
        rnd.x = rnd.y; rnd.y = rnd.z; rnd.z = rnd.w;
        int64_t b, e;
        load_row(a, v, b, e);
        v = pick_neighbor(a, v, b, e, r, r * 0x9E3779B1u + (uint32_t)s);
        o.put(s, v, s == L);
    }
}

// Second-order walk.  One iteration of the loop = one rejection trial of this thread's walk.
template <int BLOCK, bool STAGE, bool TABLE, bool SPECULATE>
__global__ void __launch_bounds__(BLOCK) node2vec_walk_kernel(const WalkArgs a) {
    __shared__ int64_t ring[STAGE ? 4 : 1][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowOut<BLOCK, STAGE> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x);
    const int L = a.walk_length;
    const uint32_t wlo = (uint32_t)wid, whi = (uint32_t)(wid >> 32);

    int64_t t = __ldg(a.targets + i);
    o.put(0, t, L == 0);
    if (L == 0) return;
    int64_t tb, te;
    load_row(a, t, tb, te);
    uint4 rnd = philox4x32_10(make_uint4(wlo, whi, 1u, 0u), a.key);
    int64_t v = pick_neighbor(a, t, tb, te, rnd.x, rnd.z);  // first step is uniform (rw_cuda.cu:138)
    o.put(1, v, L == 1);
    if (L == 1) return;
    int64_t vb, ve;
    load_row(a, v, vb, ve);

    const uint64_t thr_any = min(a.thr0, min(a.thr1, a.thr2));  // below this every class accepts
    const uint64_t thr_far = max(a.thr1, a.thr2);               // at or above this only x == t can accept
    int s = 2;
    uint32_t trial = 0;
    while (s <= L) {
        uint32_t r, u, r_hi;
        if ((trial & 1u) == 0u) {
            rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, trial >> 1), a.key);
            r = rnd.x; u = rnd.y; r_hi = rnd.z;
        } else {
            r = rnd.z; u = rnd.w; r_hi = rnd.x;
        }
        const int64_t x = pick_neighbor(a, v, vb, ve, r, r_hi);
        const bool back = (x == t);
        const bool possible = back ? (u < a.thr0) : (u < thr_far);
        int64_t xb = 0, xe = 0;
        if (SPECULATE && possible && s < L) load_row(a, x, xb, xe);
        bool accept;
        if (u < thr_any) accept = true;
        else if (back) accept = u < a.thr0;
        else if (!possible) accept = false;
        else if (a.thr1 == a.thr2) accept = true;  // q == 1: membership cannot change the answer
        else accept = u < (is_member<TABLE>(x, tb, te, a.col_idx, a.table) ? a.thr1 : a.thr2);
        if (accept) {
            o.put(s, x, s == L);
            if (!SPECULATE && s < L) load_row(a, x, xb, xe);
            t = v; tb = vb; te = ve;
            v = x; vb = xb; ve = xe;
            ++s;
            trial = 0;
        } else {
            ++trial;
        }
    }
}

template <int BLOCK, bool STAGE>
static void launch_uniform(const WalkArgs& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    uniform_walk_kernel<BLOCK, STAGE><<<grid, BLOCK, 0, st>>>(a);
}

template <int BLOCK, bool STAGE, bool TABLE>
static void launch_n2v(const WalkArgs& a, bool speculate, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    if (speculate) node2vec_walk_kernel<BLOCK, STAGE, TABLE, true><<<grid, BLOCK, 0, st>>>(a);
    else node2vec_walk_kernel<BLOCK, STAGE, TABLE, false><<<grid, BLOCK, 0, st>>>(a);
}

static uint64_t threshold(double prob) {
    double t = prob * 4294967296.0;
    if (!(t > 0.0)) return 0;
    if (t >= 4294967296.0) return 4294967296ull;
    return (uint64_t)t;
}

static size_t table_bytes(int64_t nnz) { return (size_t)((nnz + 3) / 4) * 32 + 256; }

// Optional persisting-L2 window over row_ptr (the small, degree-skewed array every step reads).
static void set_row_ptr_window(cudaStream_t st, const void* base, size_t bytes, int device, bool on) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (on) {
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        size_t carve = min((size_t)options().persist_l2_mb << 20, (size_t)max_persist);
        if (carve == 0 || max_win == 0) return;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        size_t win = min(bytes, (size_t)max_win);
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = (float)min(1.0, (double)carve / (double)win);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

}  // namespace trw

using namespace trw;

extern "C" size_t trw_walk_csr_workspace_bytes(int64_t n_nodes, int64_t nnz, double p, double q) {
    if (p == 1.0 && q == 1.0) return 0;
    if (n_nodes < 0 || nnz <= 0 || (uint64_t)n_nodes >= 0xFFFFFFFFull) return 0;  // ids must fit uint32 slots
    return table_bytes(nnz);
}

namespace trw {

// Validates the graph-side arguments, derives the acceptance thresholds and (node2vec only)
// builds the membership table into `workspace`.  After this the plan can launch any number of
// shards of start nodes (trw_walk_csr launches one; trw_walk_csr_host one per chunk).
int csr_walk_prepare(CsrWalkPlan* plan, const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                     double p, double q, int walk_length, int64_t seed, void* workspace, size_t workspace_bytes,
                     int device, cudaStream_t st) {
    if (n_nodes < 0 || nnz < 0 || walk_length < 0) { set_error("trw_walk_csr: negative size"); return TRW_ERR_ARG; }
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_csr: p and q must be positive"); return TRW_ERR_ARG; }
    if (!row_ptr || (nnz > 0 && !col_idx)) { set_error("trw_walk_csr: null pointer"); return TRW_ERR_ARG; }
    const Options& opt = options();
    WalkArgs& a = plan->a;
    a.row_ptr = row_ptr; a.col_idx = col_idx; a.n_nodes = n_nodes; a.nnz = nnz;
    a.targets = nullptr; a.n_walks = 0; a.walk_id_offset = 0;
    a.walk_length = walk_length; a.key = philox_key(seed, kTagWalkCsr);
    a.out = nullptr; a.out_row_stride = 0; a.table = nullptr;
    a.thr0 = a.thr1 = a.thr2 = 0;
    plan->device = device;
    plan->uniform = (p == 1.0 && q == 1.0);  // rw_cuda.cu:226
    plan->table = false;
    plan->speculate = opt.n2v_speculate != 0;
    plan->stage = opt.stage_output != 0;
    plan->persist = opt.persist_row_ptr != 0;
    if (plan->uniform) return TRW_OK;
    const double mx = fmax(fmax(1.0 / p, 1.0), 1.0 / q);  // rw_cuda.cu:119-123
    a.thr0 = threshold(1.0 / p / mx);
    a.thr1 = threshold(1.0 / mx);
    a.thr2 = threshold(1.0 / q / mx);
    const size_t need = trw_walk_csr_workspace_bytes(n_nodes, nnz, p, q);
    const bool table = opt.n2v_table != 0 && workspace != nullptr && need > 0 && a.thr1 != a.thr2;
    if (!table) return TRW_OK;
    if (workspace_bytes < need || ((uintptr_t)workspace & 255)) {
        set_error("trw_walk_csr: workspace needs %zu bytes at 256-byte alignment (got %zu)", need, workspace_bytes);
        return TRW_ERR_WORKSPACE;
    }
    timing_begin(0, st);
    int rc = check_cuda(cudaMemsetAsync(workspace, 0xFF, need, st), "table memset");
    if (rc) return rc;
    const unsigned grid = (unsigned)((nnz + kBuildTile - 1) / kBuildTile);
    build_member_table_kernel<<<grid, kBuildThreads, 0, st>>>(row_ptr, col_idx, n_nodes, nnz, (uint32_t*)workspace);
    timing_end(0, st);
    count_launch(1);
    rc = check_cuda(cudaGetLastError(), "build_member_table launch");
    if (rc) return rc;
    a.table = (const uint32_t*)workspace;
    plan->table = true;
    return TRW_OK;
}

int csr_walk_launch(const CsrWalkPlan& plan, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                    int64_t* out, int64_t out_row_stride, cudaStream_t st) {
    if (n_walks <= 0) return TRW_OK;
    WalkArgs a = plan.a;
    a.targets = targets; a.n_walks = n_walks; a.walk_id_offset = walk_id_offset;
    a.out = out; a.out_row_stride = out_row_stride;
    const bool stage = plan.stage && (((uintptr_t)out & 7) == 0);
    if (plan.persist) set_row_ptr_window(st, a.row_ptr, (size_t)(a.n_nodes + 1) * 8, plan.device, true);
    constexpr int BLOCK = 256;
    timing_begin(1, st);
    if (plan.uniform) {
        if (stage) launch_uniform<BLOCK, true>(a, st); else launch_uniform<BLOCK, false>(a, st);
    } else if (plan.table) {
        if (stage) launch_n2v<BLOCK, true, true>(a, plan.speculate, st); else launch_n2v<BLOCK, false, true>(a, plan.speculate, st);
    } else {
        if (stage) launch_n2v<BLOCK, true, false>(a, plan.speculate, st); else launch_n2v<BLOCK, false, false>(a, plan.speculate, st);
    }
    timing_end(1, st);
    count_launch(1);
    const int rc = check_cuda(cudaGetLastError(), "walk kernel launch");
    if (plan.persist) set_row_ptr_window(st, nullptr, 0, plan.device, false);
    return rc;
}

}  // namespace trw

extern "C" int trw_walk_csr(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                            const int64_t* targets, int64_t n_walks, int64_t walk_id_offset, double p, double q,
                            int walk_length, int64_t seed, int64_t* out, int64_t out_row_stride, void* workspace,
                            size_t workspace_bytes, int device, void* stream) {
    if (n_walks < 0 || walk_length < 0 || out_row_stride < (int64_t)walk_length + 1) {
        set_error("trw_walk_csr: negative size or out_row_stride < walk_length+1");
        return TRW_ERR_ARG;
    }
    if (n_walks > 0 && (!targets || !out)) { set_error("trw_walk_csr: null pointer"); return TRW_ERR_ARG; }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_csr: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    cudaStream_t st = (cudaStream_t)stream;
    CsrWalkPlan plan;
    int rc = csr_walk_prepare(&plan, row_ptr, col_idx, n_nodes, nnz, p, q, walk_length, seed, workspace, workspace_bytes, d, st);
    if (rc) return rc;
    return csr_walk_launch(plan, targets, n_walks, walk_id_offset, out, out_row_stride, st);
}
