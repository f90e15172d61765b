// Hashed adjacency-membership table of the node2vec walk: layout, device-side lookup, and the
// host entry that builds it (member_table.cu).
//
// The reference answers "is x a neighbour of t?" by scanning adj(t) (csrc/cuda/rw_cuda.cu:33-57),
// O(deg) per rejection trial.  Here it is one 32-byte sector.  Row t of the CSR owns the bytes
// [8*row_ptr[t], 8*row_ptr[t+1]) of a table as large as col_idx; the 32-byte buckets wholly inside
// that span hold the row's neighbour ids as uint32 (8 slots per bucket, EMPTY = 0xFFFFFFFF), with
// open addressing over buckets.  No per-row pointer is needed: the bucket range follows from the
// row span the walk already holds.  A row of degree d >= kMinTableDeg owns nb >= (d-6)/4 buckets,
// i.e. at least 2d-12 >= d slots.  A shorter row keeps a packed uint32 copy of its d ids at the start
// of its own bytes (words [2b, 2b+d), EMPTY up to 2e): at most six 8-byte loads instead of a scan of
// the int64 adjacency, whose predicated 64-bit loads and compares were 8 % of the walk kernel's
// instructions (the kernel is partly issue-bound: profiles/).
//
// Probing stays inside a *segment* of the row's buckets: rows with fewer than 2*kSegBuckets
// buckets are one segment (so the capacity bound above is a guarantee); longer rows (hubs) are cut
// into segments of kSegBuckets buckets (the last one up to 2*kSegBuckets-1), which is what lets a
// CTA assemble one segment in shared memory.  A hub segment holds on average <= 0.67 * capacity
// entries with a standard deviation of a few dozen, so it cannot fill up short of an adversarial
// hash collision set; if it ever does, the build raises a flag and the walk falls back to the scan.
// Lookup: hash -> home bucket -> one sector; a hit, or any EMPTY slot (that bucket never
// overflowed), ends the probe.
#pragma once

#include "trw_common.cuh"

namespace trw {

constexpr int64_t kMinTableDeg = 12;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kSegShift = 11;
constexpr int64_t kSegBuckets = 1ll << kSegShift;

__host__ __device__ __forceinline__ void table_span(int64_t b, int64_t e, int64_t& first, int64_t& nb) {
    first = (b + 3) >> 2;
    nb = (e >> 2) - first;
}

// Home bucket of x in a row of nb buckets, and the segment [lo, hi) probing is confined to.
__device__ __forceinline__ int64_t home_bucket(uint32_t x, int64_t nb) {
    return (int64_t)__umul64hi((uint64_t)mix32(x) << 32, (uint64_t)nb);
}
__host__ __device__ __forceinline__ int64_t segment_count(int64_t nb) {
    const int64_t n = nb >> kSegShift;
    return n > 1 ? n : 1;
}
__device__ __forceinline__ void probe_segment(int64_t nb, int64_t home, int64_t& lo, int64_t& hi) {
    const int64_t nseg = segment_count(nb);
    int64_t j = home >> kSegShift;
    if (j > nseg - 1) j = nseg - 1;
    lo = j << kSegShift;
    hi = (j == nseg - 1) ? nb : lo + kSegBuckets;
}

// x in adj(t)?  (b,e) = row span of t.  table == nullptr (or TABLE == false) selects the scan.
template <bool TABLE>
__device__ __forceinline__ bool is_member(int64_t x, int64_t b, int64_t e, IdxPtr col_idx,
                                          const uint32_t* __restrict__ table, uint64_t pol_stream) {
    // Table slots hold uint32 ids below kEmpty; an id that does not fit (out-of-graph, possible only in a
    // hand-built col_idx) is never stored there, so the table cannot speak for it: such an x is scanned.
    const bool fits = (uint64_t)x < (uint64_t)kEmpty;
    if (TABLE && table != nullptr && fits && e - b < kMinTableDeg) {
        const uint32_t x32 = (uint32_t)x;
        const uint32_t* words = table + 2 * b;
        const int d = (int)(e - b);
        // all (at most six) loads are issued before the first compare: one memory latency, not one per pair
        uint2 w[(kMinTableDeg) / 2];
#pragma unroll
        for (int k = 0; k < (int)kMinTableDeg / 2; ++k)  // the word after an odd-length row's last id is EMPTY
            w[k] = 2 * k < d ? ldg_u32x2_hint(words + 2 * k, pol_stream) : make_uint2(kEmpty, kEmpty);
        bool found = false;
#pragma unroll
        for (int k = 0; k < (int)kMinTableDeg / 2; ++k) found |= (w[k].x == x32) | (w[k].y == x32);
        return found;
    }
    if (TABLE && table != nullptr && fits) {
        int64_t first, nb, lo, hi;
        table_span(b, e, first, nb);
        const uint32_t x32 = (uint32_t)x;
        int64_t bkt = home_bucket(x32, nb);
        probe_segment(nb, bkt, lo, hi);
        for (int64_t probes = lo; probes < hi; ++probes) {
            const Sector64 s = ldg_sector_hint(table + (first + bkt) * 8, pol_stream);
            const uint32_t w[8] = {(uint32_t)s.a, (uint32_t)(s.a >> 32), (uint32_t)s.b, (uint32_t)(s.b >> 32),
                                   (uint32_t)s.c, (uint32_t)(s.c >> 32), (uint32_t)s.d, (uint32_t)(s.d >> 32)};
            bool hit = false, open = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) { hit |= (w[j] == x32); open |= (w[j] == kEmpty); }
            if (hit) return true;
            if (open) return false;
            if (++bkt == hi) bkt = lo;
        }
        return false;
    }
    // Short (or table-less) row: the reference's scan, csrc/cuda/rw_cuda.cu:48-53, eight
    // independent loads per round so that it is not one dependent chain.
    for (int64_t i = b; i < e; i += 8) {
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i + j < e) found |= (ldg64_hint(col_idx + i + j, pol_stream) == x);
        if (found) return true;
    }
    return false;
}

// ---------------------------------------------------------------- hub-pair filter (L2-resident)
// The triangle Blooms (below) answer most membership questions of a walk for free, but a 32-bit word cannot describe the
// common neighbourhood of two hubs: edges whose shorter row exceeds the Bloom cap keep a saturated word, and on an R-MAT
// graph those are 40 % of the edges a walk crosses -- a fifth of the questions still reach the table (one random
// 128-byte line of HBM each), nearly all of them to hear "no".  They all concern pairs of hubs, and the hub-hub edges are
// few enough for a bitmap that stays in L2: one bit per hashed unordered pair {row, id}, set by the Bloom pass for every
// stored entry whose two endpoints both have more than `min_deg` out-entries (evict_last; option edge_filter_mb).  For
// such a pair a clear bit proves that neither (row, id) nor (id, row) is stored, so "not a member" is exact; a set bit
// says "maybe" and the caller asks the table as before; any other pair is never asked.  Same answers, fewer lines.
// Ids that do not fit 32 bits are never inserted and never asked.
struct EdgeFilter {
    const uint32_t* bits = nullptr;
    uint32_t n_bits = 0;
    uint32_t min_deg = 0;  // the filter knows every stored pair whose endpoints both have more out-entries than this
};
__host__ __device__ __forceinline__ uint32_t pair_slot(uint32_t a, uint32_t b, uint32_t n_bits) {
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    uint64_t z = (((uint64_t)hi << 32) | lo) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 32;
    z *= 0xD6E8FEB86659FD93ull;
    return mulhi32((uint32_t)(z >> 32), n_bits);
}
// false: `x` is certainly not stored in the row of node `row`; true: ask the table.  d_row, d_x: their out-degrees.
__device__ __forceinline__ bool filter_maybe(const EdgeFilter& f, int64_t row, int64_t x, int64_t d_row, int64_t d_x, uint64_t pol_keep) {
    if (f.bits == nullptr || d_row <= (int64_t)f.min_deg || d_x <= (int64_t)f.min_deg || ((uint64_t)row | (uint64_t)x) >= (uint64_t)kEmpty) return true;
    const uint32_t slot = pair_slot((uint32_t)row, (uint32_t)x, f.n_bits);
    return (ldg32_l2keep(f.bits + (slot >> 5), pol_keep) >> (slot & 31)) & 1u;
}

// Does x occur at least TWICE in the row?  Only meaningful for tables assembled in shared memory
// (build_mode 2), which store every entry of a row, duplicates included.  Used by the edge-list walk,
// whose reference semantics ignore the last out-edge of a row (walk_other.cu).
__device__ __forceinline__ bool member_twice(int64_t x, int64_t b, int64_t e, const uint32_t* __restrict__ table,
                                             uint64_t pol_stream) {
    const uint32_t x32 = (uint32_t)x;
    int count = 0;
    if (e - b < kMinTableDeg) {
        const uint32_t* words = table + 2 * b;
        const int d = (int)(e - b);
#pragma unroll
        for (int k = 0; k < (int)kMinTableDeg / 2; ++k) {
            const uint2 w = 2 * k < d ? ldg_u32x2_hint(words + 2 * k, pol_stream) : make_uint2(kEmpty, kEmpty);
            count += (w.x == x32) + (w.y == x32);
        }
        return count >= 2;
    }
    int64_t first, nb, lo, hi;
    table_span(b, e, first, nb);
    int64_t bkt = home_bucket(x32, nb);
    probe_segment(nb, bkt, lo, hi);
    for (int64_t probes = lo; probes < hi; ++probes) {
        const Sector64 s = ldg_sector_hint(table + (first + bkt) * 8, pol_stream);
        const uint32_t w[8] = {(uint32_t)s.a, (uint32_t)(s.a >> 32), (uint32_t)s.b, (uint32_t)(s.b >> 32),
                               (uint32_t)s.c, (uint32_t)(s.c >> 32), (uint32_t)s.d, (uint32_t)(s.d >> 32)};
        bool open = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) { count += (w[j] == x32); open |= (w[j] == kEmpty); }
        if (count >= 2) return true;
        if (open) return false;
        if (++bkt == hi) bkt = lo;
    }
    return false;
}

// ---------------------------------------------------------------- host side (member_table.cu)
// Offsets of the per-call scratch inside the caller's workspace (all 256-byte aligned).
struct CsrWorkspace {
    size_t table, tile_row0, hub_list, seg_work, cells, row32, records, filter, total;
    int64_t n_tiles, n_buckets, max_hubs, max_segs;
    uint32_t filter_bits;
    bool has_table, has_row32, has_records;
};
CsrWorkspace csr_workspace_layout(int64_t n_nodes, int64_t nnz, bool uniform, bool records, bool filter = true);

// Edge record k of the CSR (option `records`): the neighbour id col_idx[k] together with that
// neighbour's own row span, so that the gather which proposes the next node also returns where
// its adjacency lives and the walk never touches the row index again (16 bytes, one LDG.128).
//   .x = id, low 32 bits   .y = degree of id   .z = row start of id
//   .w = degree 0: id, high 32 bits (an id outside [0, n_nodes) gets degree 0, i.e. a node without
//        out-edges, as load_row() treats it -- only such an id can need a high word);
//        degree > 0: the TRIANGLE BLOOM of the edge, see below.
//
// Triangle Bloom of entry (t -> v): 32 bits, bit bloom_bit(w) set for every w in adj(t) & adj(v)
// (kBloomAll when it was not computed: one-shot calls, pairs of two hubs).  A node2vec step at v
// coming from t asks "x in adj(t)?" for x drawn from adj(v); x can only be in adj(t) if it is a common
// neighbour of t and v, so a clear bit bloom_bit(x) in the word of (t -> v) -- fetched for free when v
// was proposed -- answers "no" without touching memory.  On a symmetric graph the word of (v -> x),
// fetched with the proposal itself, gives a second, independent look at the same triangle through
// bloom_bit(t).  Both are one-sided (a set bit means "ask the table"), so the walk is bit-identical.
constexpr uint32_t kBloomAll = 0xFFFFFFFFu;
__host__ __device__ __forceinline__ uint32_t bloom_bit(int64_t w) { return 1u << (((uint32_t)w * 0x9E3779B1u) >> 27); }
__host__ __device__ __forceinline__ uint32_t record_last_word(uint64_t id, uint32_t deg, uint32_t bloom) {
    return deg != 0 ? bloom : (uint32_t)(id >> 32);
}

struct CsrPrepared {
    const uint32_t* table = nullptr;     // membership table, or nullptr
    const uint32_t* row32 = nullptr;     // uint32 copy of row_ptr, or nullptr
    const int* table_failed = nullptr;   // device flag: non-zero when a hub segment overflowed
    const unsigned long long* strict_counts = nullptr;  // [descents in col_idx, descents at row boundaries]
    const uint4* records = nullptr;      // edge records, or nullptr
    EdgeFilter filter;                   // L2-resident hub-pair filter, filled by the Bloom pass (bits == nullptr: none, or not filled yet)
    uint32_t* filter_space = nullptr;    // where it goes: workspace reserved by csr_prepare_device
    uint32_t filter_space_bits = 0;
    const int* asymmetric = nullptr;     // device flag of the triangle-Bloom pass: non-zero when some (t -> v) has no (v -> t); nullptr: not checked
    int* bloom_flag = nullptr;           // where that flag lives once the pass has run (workspace cell)
};

// Enqueues on `st` the per-call preparation of a CSR graph: the uint32 row index and (node2vec)
// the membership table, and (want_strict) the two counters that tell whether every row is strictly
// increasing, i.e. sorted without duplicate edges.  build_mode: 2 = shared memory (default), 0 = global CAS.
int csr_prepare_device(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, void* workspace,
                       const CsrWorkspace& w, bool want_table, bool want_row32, bool want_strict, bool want_records,
                       int build_mode, int device, cudaStream_t st, CsrPrepared* out, int64_t bloom_cap = 0);

// Adds the triangle Blooms to the edge records of a prepared graph (no-op when they are there, or when the
// graph has no table / records to hang them on).
int csr_add_blooms(CsrPrepared* pr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, int64_t cap, int device, cudaStream_t st);

}  // namespace trw
