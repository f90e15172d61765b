// Edge-list and knowledge-graph triple walks for sm_100a (one thread per walk).
//
// Replaces walk_edge_list_gpu (csrc/cuda/rw_cuda_edge_list.cu:243-308; kernels :42-96 and
// :126-240) and triples::walk_triples_gpu (csrc/cuda/rw_cuda_triples.cu:103-169; kernel :49-96).
// Both keep the reference's transition rules exactly -- including, for the second-order
// edge-list walk, the acceptance chain of rw_cuda_edge_list.cu:203-230 whose empirical
// distribution differs from textbook node2vec: parity for that path is
// defined against the reference, so its rule is reproduced rather than "fixed".
//
// The reference tests "x is a neighbour of t" by scanning t's out-edges (rw_cuda_edge_list.cu:98-123),
// O(deg(t)) per rejection trial: 1-4 M steps/s on a power-law graph.  With a workspace
// (trw_walk_edge_list_ws) the edge list is viewed as a CSR -- tails copied contiguously, row starts
// found by bisection of the sorted heads -- and the CSR walk's hashed membership table
// (member_table.cuh) answers the test with one sector, reference quirk included (see
// el_is_neighbor_table).  Same draws, same decisions: the walks are bit-identical to the scan path.
#include "member_table.cuh"
#include "trw_common.cuh"
#include "trw_options.h"

namespace trw {

struct IndexedWalkArgs {
    const int64_t* rows;   // edge_list[n_rows,2] or triples[n_rows,3], sorted by head
    int64_t n_rows;
    const int64_t* index;  // [n_index_rows,2] inclusive [first,last] or [-1,-1]
    int64_t n_index_rows;
    const int64_t* targets;
    int64_t n_walks, walk_id_offset;
    int walk_length;
    uint2 key;
    int64_t pad;
    int restart;
    int64_t* out;
    int64_t out_row_stride;
    uint64_t thr0, thr1, thr2;
    // membership table over the CSR view of the edge list (all null: scan, as the reference does)
    const int64_t* col;            // tails, contiguous: col[k] = rows[2k+1]
    const uint32_t* table;
    const int* table_failed;       // a hub segment overflowed during the build
    const int* view_mismatch;      // node_edge_index disagrees with the bisection of the heads (or heads unsorted)
};

// Inclusive row range of node v; false when v has no rows (or lies outside the index, which the
// reference would read out of bounds -- e.g. row `padding_idx` when padding_idx == N).
__device__ __forceinline__ bool row_range(const IndexedWalkArgs& a, int64_t v, int64_t& first, int64_t& last) {
    if ((uint64_t)v >= (uint64_t)a.n_index_rows) return false;
    first = __ldg(a.index + 2 * v);
    last = __ldg(a.index + 2 * v + 1);
    return !(first == -1 || last == -1);
}

// rw_cuda_edge_list.cu:13-40.
__device__ __forceinline__ int64_t el_sample(const IndexedWalkArgs& a, int64_t v, int64_t jump, uint32_t r0, uint32_t r1) {
    if (v == a.pad) return jump;
    int64_t first, last;
    if (!row_range(a, v, first, last)) return a.pad;
    const int64_t k = first + bounded(r0, r1, last + 1 - first);
    if ((uint64_t)k >= (uint64_t)a.n_rows) return a.pad;
    return ldg64_stream(a.rows + 2 * k + 1);
}

// rw_cuda_edge_list.cu:98-123: scans rows [first,last) -- the last out-edge is not looked at.
__device__ __forceinline__ bool el_is_neighbor(const IndexedWalkArgs& a, int64_t x, int64_t t) {
    int64_t first, last;
    if (!row_range(a, t, first, last)) return false;
    for (int64_t i = first; i < last; i += 4) {
        bool found = false;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i + j < last) found |= (ldg64_stream(a.rows + 2 * (i + j) + 1) == x);
        if (found) return true;
    }
    return false;
}

// The same answer from the hashed table.  The reference's scan stops BEFORE the last out-edge of t
// (half-open [first, last) over an inclusive range: rw_cuda_edge_list.cu:112).  The table holds the
// whole row, every entry of it, so when x is that last tail it counts only if it is stored twice.
__device__ __forceinline__ bool el_is_neighbor_table(const IndexedWalkArgs& a, int64_t x, int64_t t, uint64_t pol_stream) {
    int64_t first, last;
    if (!row_range(a, t, first, last)) return false;
    if (last <= first) return false;  // one out-edge: the scanned range is empty
    if ((uint64_t)x >= (uint64_t)kEmpty) return el_is_neighbor(a, x, t);  // an id the uint32 table never holds
    if (!is_member<true>(x, first, last + 1, a.col, a.table, pol_stream)) return false;
    if (ldg64_hint(a.col + last, pol_stream) != x) return true;
    return member_twice(x, first, last + 1, a.table, pol_stream);
}

// CSR view of an edge list sorted by head: contiguous tails, row starts by bisection of the heads,
// and a check that node_edge_index says the same (it is what the walk samples from).
__global__ void __launch_bounds__(256) edge_list_view_kernel(const int64_t* __restrict__ rows, int64_t n_rows,
                                                             const int64_t* __restrict__ index, int64_t n_index_rows,
                                                             int64_t* __restrict__ col, int64_t* __restrict__ row_ptr,
                                                             int* __restrict__ mismatch) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = gtid; k < n_rows; k += gsz) {
        col[k] = ldg64_stream(rows + 2 * k + 1);
        if (k + 1 < n_rows && __ldg(rows + 2 * k) > __ldg(rows + 2 * k + 2)) *mismatch = 1;  // heads not sorted
    }
    auto lower_bound = [&](int64_t v) {  // first edge whose head is >= v
        int64_t lo = 0, hi = n_rows;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(rows + 2 * mid) < v) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    for (int64_t v = gtid; v <= n_index_rows; v += gsz) {
        const int64_t b = lower_bound(v);
        row_ptr[v] = v == n_index_rows ? n_rows : b;
        // edges whose head lies outside the index belong to no row of it: the CSR view would hand them to
        // the first or last row, so such a list takes the reference's scan instead
        if (v == 0 && b != 0) *mismatch = 1;
        if (v == n_index_rows) {
            if (b != n_rows) *mismatch = 1;
            break;
        }
        const int64_t e = lower_bound(v + 1);
        const int64_t first = __ldg(index + 2 * v), last = __ldg(index + 2 * v + 1);
        const bool has = !(first == -1 || last == -1);
        if (has ? (first != b || last + 1 != e) : (b != e)) *mismatch = 1;
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) edge_list_uniform_kernel(const IndexedWalkArgs a) {
    __shared__ int64_t ring[4][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowStager<BLOCK> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x, make_policy_evict_first());
    const int64_t start = __ldg(a.targets + i);
    const int64_t jump = a.restart ? start : a.pad;
    const int L = a.walk_length;
    int64_t v = start;
    o.put(0, v, L == 0);
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int s = 1; s <= L; ++s) {
        if (((s - 1) & 1) == 0)
            rnd = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)((s - 1) >> 1), 0x80000000u), a.key);
        v = el_sample(a, v, jump, rnd.x, rnd.y);
        rnd.x = rnd.z; rnd.y = rnd.w;
        o.put(s, v, s == L);
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) edge_list_biased_kernel(const IndexedWalkArgs a) {
    __shared__ int64_t ring[4][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    // the table is used only when its build succeeded and the CSR view agrees with node_edge_index;
    // any doubt selects the reference's scan (same results, slower)
    const bool use_table = a.table != nullptr && *a.table_failed == 0 && *a.view_mismatch == 0;
    const uint64_t pol_stream = make_policy_evict_first();
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    const uint32_t wlo = (uint32_t)wid, whi = (uint32_t)(wid >> 32);
    RowStager<BLOCK> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x, make_policy_evict_first());
    const int64_t start = __ldg(a.targets + i);
    const int64_t jump = a.restart ? start : a.pad;
    const int L = a.walk_length;
    o.put(0, start, L == 0);
    if (L == 0) return;
    uint4 rnd = philox4x32_10(make_uint4(wlo, whi, 1u, 0u), a.key);
    int64_t t = start;
    int64_t v = el_sample(a, start, jump, rnd.x, rnd.z);
    o.put(1, v, L == 1);
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
        const int64_t x = el_sample(a, v, jump, r, r_hi);
        bool accept = false;
        int64_t sel = x;
        if (x == t && u < a.thr0) {
            accept = true;                      // return edge (rw_cuda_edge_list.cu:203-208)
        } else if (x == a.pad) {                // separate `if` in the source: reached after a rejected return too
            if (u < a.thr0) { accept = true; sel = jump; }
        } else if (use_table ? el_is_neighbor_table(a, x, t, pol_stream) : el_is_neighbor(a, x, t)) {
            accept = u < a.thr1;
        } else {
            accept = u < a.thr2;
        }
        if (accept) {
            o.put(s, sel, s == L);
            t = v;
            v = sel;
            ++s;
            trial = 0;
        } else {
            ++trial;
        }
    }
}

// rw_cuda_triples.cu:49-96 with sample_neighbor_gpu :13-47.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) triples_walk_kernel(const IndexedWalkArgs a) {
    __shared__ int64_t ring[4][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowStager<BLOCK> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x, make_policy_evict_first());
    const int L = a.walk_length;
    int64_t v = __ldg(a.targets + i);
    o.put(0, v, L == 0);
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int hop = 0; hop < L; ++hop) {
        if ((hop & 1) == 0)
            rnd = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)(hop >> 1), 0u), a.key);
        int64_t rel = a.pad, tail = a.pad, first, last;
        if (v != a.pad && row_range(a, v, first, last)) {
            const int64_t k = first + bounded(rnd.x, rnd.y, last + 1 - first);
            if ((uint64_t)k < (uint64_t)a.n_rows) {
                rel = ldg64_stream(a.rows + 3 * k + 1);
                tail = ldg64_stream(a.rows + 3 * k + 2);
            }
        }
        rnd.x = rnd.z; rnd.y = rnd.w;
        o.put(2 * hop + 1, rel, false);
        o.put(2 * hop + 2, tail, hop == L - 1);
        v = tail;
    }
}

static uint64_t threshold(double prob) {
    double t = prob * 4294967296.0;
    if (!(t > 0.0)) return 0;
    if (t >= 4294967296.0) return 4294967296ull;
    return (uint64_t)t;
}

}  // namespace trw

using namespace trw;

namespace trw {
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
struct EdgeListWorkspace {
    size_t col, row_ptr, flags, csr, total;
    CsrWorkspace w;
};
static EdgeListWorkspace edge_list_workspace_layout(int64_t n_edges, int64_t n_index_rows) {
    EdgeListWorkspace l{};
    l.w = csr_workspace_layout(n_index_rows, n_edges, /*uniform=*/false, /*records=*/false, /*filter=*/false);
    size_t off = 0;
    l.col = off; off += align256((size_t)n_edges * 8);
    l.row_ptr = off; off += align256((size_t)(n_index_rows + 1) * 8);
    l.flags = off; off += 256;
    l.csr = off; off += l.w.total;
    l.total = (l.w.has_table && n_edges > 0 && n_index_rows > 0) ? off : 0;
    return l;
}
}  // namespace trw

extern "C" size_t trw_walk_edge_list_workspace_bytes(int64_t n_edges, int64_t n_index_rows, double p, double q) {
    if (n_edges < 0 || n_index_rows < 0 || (p == 1.0 && q == 1.0) || options().el_table == 0) return 0;
    return edge_list_workspace_layout(n_edges, n_index_rows).total;
}

extern "C" int trw_walk_edge_list_ws(const int64_t* edge_list, int64_t n_edges, const int64_t* node_edge_index,
                                     int64_t n_index_rows, const int64_t* targets, int64_t n_walks,
                                     int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                                     int64_t padding_idx, int restart, int64_t* out, int64_t out_row_stride,
                                     void* workspace, size_t workspace_bytes, int device, void* stream) {
    if (n_walks < 0 || n_edges < 0 || n_index_rows < 0 || walk_length < 0 || out_row_stride < (int64_t)walk_length + 1) {
        set_error("trw_walk_edge_list: negative size or out_row_stride < walk_length+1");
        return TRW_ERR_ARG;
    }
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_edge_list: p and q must be positive"); return TRW_ERR_ARG; }
    if (n_walks > 0 && (!targets || !out || (n_index_rows > 0 && !node_edge_index) || (n_edges > 0 && !edge_list))) {
        set_error("trw_walk_edge_list: null pointer");
        return TRW_ERR_ARG;
    }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_edge_list: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    IndexedWalkArgs a{};
    a.rows = edge_list; a.n_rows = n_edges; a.index = node_edge_index; a.n_index_rows = n_index_rows;
    a.targets = targets; a.n_walks = n_walks; a.walk_id_offset = walk_id_offset; a.walk_length = walk_length;
    a.key = philox_key(seed, kTagWalkEdgeList); a.pad = padding_idx; a.restart = restart ? 1 : 0;
    a.out = out; a.out_row_stride = out_row_stride;
    const double mx = fmax(fmax(1.0 / p, 1.0), 1.0 / q);  // rw_cuda_edge_list.cu:148-152
    a.thr0 = threshold(1.0 / p / mx); a.thr1 = threshold(1.0 / mx); a.thr2 = threshold(1.0 / q / mx);
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((n_walks + BLOCK - 1) / BLOCK);
    cudaStream_t st = (cudaStream_t)stream;
    const bool uniform = (p == 1.0 && q == 1.0);  // rw_cuda_edge_list.cu:281
    // (the table must hold every entry of a row, duplicates included: only the shared-memory build does)
    if (!uniform && workspace != nullptr && options().el_table != 0 && options().build_mode != 0) {
        const EdgeListWorkspace l = edge_list_workspace_layout(n_edges, n_index_rows);
        if (l.total > 0) {
            if (workspace_bytes < l.total || ((uintptr_t)workspace & 255)) {
                set_error("trw_walk_edge_list_ws: workspace needs %zu bytes at 256-byte alignment (got %zu)", l.total, workspace_bytes);
                return TRW_ERR_WORKSPACE;
            }
            char* ws = (char*)workspace;
            int64_t* col = (int64_t*)(ws + l.col);
            int64_t* row_ptr = (int64_t*)(ws + l.row_ptr);
            int* mismatch = (int*)(ws + l.flags);
            int rc = check_cuda(cudaMemsetAsync(mismatch, 0, 256, st), "edge-list flags memset");
            if (rc) return rc;
            edge_list_view_kernel<<<sm_count(d) * 8, 256, 0, st>>>(edge_list, n_edges, node_edge_index, n_index_rows, col, row_ptr, mismatch);
            count_launch(1);
            rc = check_cuda(cudaGetLastError(), "edge-list view launch");
            if (rc) return rc;
            CsrPrepared prepared;
            rc = csr_prepare_device(row_ptr, col, n_index_rows, n_edges, ws + l.csr, l.w, /*want_table=*/true, /*want_row32=*/false,
                                    /*want_strict=*/false, /*want_records=*/false, (int)options().build_mode, d, st, &prepared);
            if (rc) return rc;
            if (prepared.table) {
                a.col = col; a.table = prepared.table; a.table_failed = prepared.table_failed;
                a.view_mismatch = mismatch;
            }
        }
    }
    if (uniform) edge_list_uniform_kernel<BLOCK><<<grid, BLOCK, 0, st>>>(a);
    else edge_list_biased_kernel<BLOCK><<<grid, BLOCK, 0, st>>>(a);
    count_launch(1);
    return check_cuda(cudaGetLastError(), "edge-list walk launch");
}

extern "C" int trw_walk_edge_list(const int64_t* edge_list, int64_t n_edges, const int64_t* node_edge_index,
                                  int64_t n_index_rows, const int64_t* targets, int64_t n_walks,
                                  int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                                  int64_t padding_idx, int restart, int64_t* out, int64_t out_row_stride, int device,
                                  void* stream) {
    return trw_walk_edge_list_ws(edge_list, n_edges, node_edge_index, n_index_rows, targets, n_walks, walk_id_offset, p, q,
                                 walk_length, seed, padding_idx, restart, out, out_row_stride, nullptr, 0, device, stream);
}

extern "C" int trw_walk_triples(const int64_t* triples, int64_t n_triples, const int64_t* relation_tail_index,
                                int64_t n_index_rows, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                                int walk_length, int64_t padding_idx, int restart, int64_t seed, int64_t* out,
                                int64_t out_row_stride, int device, void* stream) {
    (void)restart;  // ignored by the reference as well (rw_cuda_triples.cu never reads it)
    if (n_walks < 0 || n_triples < 0 || n_index_rows < 0 || walk_length < 0 ||
        out_row_stride < 2 * (int64_t)walk_length + 1) {
        set_error("trw_walk_triples: negative size or out_row_stride < 2*walk_length+1");
        return TRW_ERR_ARG;
    }
    if (n_walks > 0 && (!targets || !out || (n_index_rows > 0 && !relation_tail_index) || (n_triples > 0 && !triples))) {
        set_error("trw_walk_triples: null pointer");
        return TRW_ERR_ARG;
    }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_triples: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    IndexedWalkArgs a{};
    a.rows = triples; a.n_rows = n_triples; a.index = relation_tail_index; a.n_index_rows = n_index_rows;
    a.targets = targets; a.n_walks = n_walks; a.walk_id_offset = walk_id_offset; a.walk_length = walk_length;
    a.key = philox_key(seed, kTagWalkTriples); a.pad = padding_idx; a.restart = 0;
    a.out = out; a.out_row_stride = out_row_stride; a.thr0 = a.thr1 = a.thr2 = 0;
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((n_walks + BLOCK - 1) / BLOCK);
    triples_walk_kernel<BLOCK><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(a);
    count_launch(1);
    return check_cuda(cudaGetLastError(), "triple walk launch");
}
