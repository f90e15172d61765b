// Per-call preparation of a CSR graph for the walk kernels: the uint32 row index and the hashed
// membership table (layout and lookup in member_table.cuh).
//
// Global atomics turned out to be the bound of a straightforward build (about 37-45 G CAS/s on
// B200 whether or not the lines are L2-resident; profiles/), so the shipped build has none on the
// data path.  Every bucket is assembled in shared memory by exactly one CTA and written once with
// coalesced 16-byte stores, which also removes the need to clear the table first:
//   * short rows (< kHubDeg neighbours): the CTA that owns the tile of kBuildTile CSR entries in
//     which a row STARTS builds all of that row's buckets (build_tiled_kernel);
//   * hub rows: one CTA per segment of kSegBuckets buckets scans the whole row and keeps the ids
//     whose home bucket falls into its segment (build_hub_kernel).  The row is re-read once per
//     segment, from L2.
// A pre-pass finds the first row of every tile (edge-balanced tiling of a power-law graph), lists
// the hub rows and their segments, and writes the uint32 row index.
#include <algorithm>

#include "member_table.cuh"
#include "trw_options.h"

namespace trw {

constexpr int kBuildThreads = 256;
constexpr int kBuildPerThread = 8;
constexpr int kBuildTile = kBuildThreads * kBuildPerThread;  // CSR entries per tile
constexpr int kHubDeg = kBuildTile;                          // rows at least this long are hubs
constexpr int kBuildRange = kBuildTile + kHubDeg;            // entries a tile's CTA may touch
constexpr int kRangePerThread = kBuildRange / kBuildThreads;
constexpr uint32_t kNotOurs = 0xFFFFFFFFu;
constexpr int kHeadWords = kBuildRange + kBuildRange / 32;  // row codes, skewed by one word per 32 (bank-conflict-free both ways)
constexpr size_t kTiledSmemBytes = (size_t)kHeadWords * 4 + (size_t)kBuildRange * 2 * 4 + (size_t)kBuildRange;  // codes + bucket image + counters
__device__ __forceinline__ int head_at(int i) { return i + (i >> 5); }
constexpr size_t kHubSmemBytes = (size_t)(2 * kSegBuckets) * 32 + (size_t)(2 * kSegBuckets) * 4;                  // one segment image + counters
constexpr int kHubThreads = 1024;

struct HubEntry {
    int64_t row;
    uint32_t first_segment;  // index of its first segment in the global segment list
    uint32_t n_segments;
};
struct SegmentWork {
    int64_t row;
    int64_t segment;
};

struct BuildArgs {
    IdxPtr row_ptr;
    IdxPtr col_idx;
    int64_t n_nodes, nnz;
    uint32_t* table;
    int64_t* tile_row0;              // [n_tiles + 1]: row holding the first entry of each tile
    HubEntry* hubs;                  // [max_hubs]
    SegmentWork* segments;           // [max_segs]
    unsigned long long* hub_counter; // (number of hubs << 32) | number of hub segments
    unsigned long long* next_segment; // work counter of build_hub_kernel
    int* failed;                     // set when a segment had no room (see member_table.cuh)
    uint32_t* row32;                 // optional uint32 copy of row_ptr
    int64_t n_tiles, n_buckets, max_hubs, max_segs;
};

// A neighbour id as the uint32 structures store it; kEmpty for one that does not fit (never stored:
// is_member scans for such ids).
__device__ __forceinline__ uint32_t slot_id(int64_t x) { return (uint64_t)x < (uint64_t)kEmpty ? (uint32_t)x : kEmpty; }

// Hub-pair filter (member_table.cuh): set the bit of the unordered pair {row, x}.  On a symmetric graph the mirror entry
// (x, row) sets the same bit, so the larger-row half looks first and usually finds it set: half the
// atomics.  A stale look only costs a redundant RED.
__device__ __forceinline__ void filter_insert(uint32_t* __restrict__ filter, uint32_t n_bits, uint32_t row, uint32_t x,
                                              uint64_t pol_keep) {
    const uint32_t slot = pair_slot(row, x, n_bits);
    uint32_t* word = filter + (slot >> 5);
    const uint32_t bit = 1u << (slot & 31);
    if (row > x) {
        uint32_t seen;
        asm volatile("ld.relaxed.gpu.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(seen) : "l"(word), "l"(pol_keep) : "memory");
        if (seen & bit) return;
    }
    red_or32_hint(word, bit, pol_keep);
}

// Largest r with row_ptr[r] <= e (the non-empty row that holds CSR entry e).
__device__ __forceinline__ int64_t row_of_entry(IdxPtr row_ptr, int64_t n_nodes, int64_t e) {
    int64_t lo = 0, hi = n_nodes;  // invariant: row_ptr[lo] <= e < row_ptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (ldg_idx(row_ptr + mid) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

// Pre-pass: first row of every tile, the uint32 row index, and the list of hub rows.
__global__ void __launch_bounds__(256) csr_prepass_kernel(const BuildArgs a) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    if (a.tile_row0) {
        for (int64_t j = gtid; j <= a.n_tiles; j += gsz) {
            const int64_t e = j * kBuildTile;
            a.tile_row0[j] = e < a.nnz ? row_of_entry(a.row_ptr, a.n_nodes, e) : a.n_nodes;
        }
    }
    for (int64_t r = gtid; r <= a.n_nodes; r += gsz) {
        const int64_t b = ldg_idx(a.row_ptr + r);
        if (a.row32) a.row32[r] = (uint32_t)b;
        if (a.hubs && r < a.n_nodes) {
            const int64_t e = ldg_idx(a.row_ptr + r + 1);
            if (e - b >= kHubDeg) {
                int64_t first, nb;
                table_span(b, e, first, nb);
                const unsigned long long nseg = (unsigned long long)segment_count(nb);
                const unsigned long long got = atomicAdd(a.hub_counter, (1ull << 32) + nseg);
                const uint64_t slot = got >> 32;
                if (slot < (uint64_t)a.max_hubs) a.hubs[slot] = HubEntry{r, (uint32_t)got, (uint32_t)nseg};
                else *a.failed = 1;
            }
        }
    }
}

// Expands the hub list into one work item per segment.
__global__ void __launch_bounds__(256) hub_segments_kernel(const BuildArgs a) {
    const unsigned long long counter = *a.hub_counter;
    const int64_t n_hubs = min((int64_t)(counter >> 32), a.max_hubs);
    if ((int64_t)(counter & 0xFFFFFFFFull) > a.max_segs) { *a.failed = 1; return; }
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n_hubs; h += (int64_t)gridDim.x * blockDim.x) {
        const HubEntry hub = a.hubs[h];
        for (uint32_t j = 0; j < hub.n_segments; ++j) a.segments[hub.first_segment + j] = SegmentWork{hub.row, (int64_t)j};
    }
}

// Insert into a bucket image in shared memory; probing wraps inside [0, n_buckets).  A per-bucket
// arrival counter hands out the slot, so an insert costs one shared-memory atomicAdd and one plain
// store (a CAS walk over the slots costs one atomic per occupied slot, and shared-memory atomics
// are what bounds these kernels).  Counters may run past 8; slots 0..7 are the bucket.
__device__ __forceinline__ bool smem_insert(uint32_t* __restrict__ image, uint32_t* __restrict__ count, int64_t n_buckets,
                                            int64_t start, uint32_t x) {
    int64_t bkt = start;
    for (int64_t probes = 0; probes < n_buckets; ++probes) {
        const uint32_t pos = atomicAdd(count + bkt, 1u);
        if (pos < 8u) {
            image[bkt * 8 + pos] = x;
            return true;
        }
        if (++bkt == n_buckets) bkt = 0;
    }
    return false;
}

// Short rows.  A CTA owns the rows that start inside its tile and are shorter than kHubDeg; such
// a row ends less than kHubDeg entries past the tile, so the CTA works on a window of kBuildRange
// entries.  The row of every entry is recovered with a max-scan over "a row starts here" codes.
template <bool WIDE>
__global__ void __launch_bounds__(kBuildThreads, 4) build_tiled_kernel(const BuildArgs a) {
    const IdxPtrT<WIDE> col_idx(a.col_idx);
    extern __shared__ __align__(16) uint32_t tiled_smem[];
    uint32_t* head = tiled_smem;                // [kHeadWords] row code of each entry of the window (index through head_at)
    uint32_t* image = tiled_smem + kHeadWords;  // [kBuildRange / 4 buckets][8 slots]; kHeadWords is a multiple of 4
    uint32_t* count = image + 2 * kBuildRange;  // [kBuildRange / 4] arrivals per bucket
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    __shared__ unsigned long long s_lo, s_hi;   // entry span [s_lo, s_hi) of the rows built here

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t e0 = tile * kBuildTile;
    const int64_t e1 = min(e0 + (int64_t)kBuildTile, a.nnz);
    const int64_t r0 = a.tile_row0[tile];
    const int64_t r1 = min(a.tile_row0[tile + 1], a.n_nodes - 1);
    const int64_t bucket0 = e0 >> 2;  // first bucket of the window (tiles are multiples of four entries)
    for (int i = tid; i < kHeadWords / 4; i += kBuildThreads) reinterpret_cast<uint4*>(head)[i] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < kRangePerThread / 2; ++k)
        reinterpret_cast<uint4*>(image)[k * kBuildThreads + tid] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
    if (tid < kBuildRange / 16) reinterpret_cast<uint4*>(count)[tid] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { s_lo = ~0ull; s_hi = 0; }
    __syncthreads();
    // Rows that start in [e0, e1): short ones get the code r - r0 + 1; a hub gets kNotOurs (nothing
    // else can start after it inside the tile: it is at least as long as the tile).
    for (int64_t r = r0 + tid; r <= r1; r += kBuildThreads) {
        const int64_t b = ldg_idx(a.row_ptr + r), e = ldg_idx(a.row_ptr + r + 1);
        if (e > b && b >= e0 && b < e1) {
            if (e - b >= kHubDeg) {
                head[head_at((int)(b - e0))] = kNotOurs;
            } else {
                head[head_at((int)(b - e0))] = (uint32_t)(r - r0 + 1);
                atomicMin(&s_lo, (unsigned long long)b);
                atomicMax(&s_hi, (unsigned long long)e);
            }
        }
    }
    __syncthreads();
    const int64_t own_end = (int64_t)s_hi;  // 0 when no short row starts here
    if (own_end == 0) return;
    // Entries past the last owned row belong to rows of later tiles.
    if (tid == 0 && own_end - e0 < kBuildRange) head[head_at((int)(own_end - e0))] = kNotOurs;
    __syncthreads();
    {   // inclusive max-scan of the codes; each thread scans kRangePerThread consecutive entries
        uint32_t own[kRangePerThread];
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kRangePerThread; ++k) {
            run = max(run, head[head_at(tid * kRangePerThread + k)]);
            own[k] = run;
        }
        uint32_t incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((tid & 31) >= d) incl = max(incl, o);
        }
        if ((tid & 31) == 31) warp_max[tid >> 5] = incl;
        __syncthreads();
        uint32_t before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if ((tid & 31) == 0) before = 0;
        for (int w = 0; w < (tid >> 5); ++w) before = max(before, warp_max[w]);
#pragma unroll
        for (int k = 0; k < kRangePerThread; ++k) head[head_at(tid * kRangePerThread + k)] = max(before, own[k]);
    }
    __syncthreads();
    // Fetch all of this thread's entries before the first insert (strided: coalesced), so the
    // window costs one DRAM latency, not one per entry.
    uint32_t xs[kRangePerThread];
#pragma unroll
    for (int k = 0; k < kRangePerThread; ++k) {
        const int idx = k * kBuildThreads + tid;
        const uint32_t code = head[head_at(idx)];
        const bool ours = code != 0 && code != kNotOurs && e0 + idx < own_end;
        xs[k] = ours ? slot_id(ldg64_stream(col_idx + (e0 + idx))) : kEmpty;
    }
    int64_t cur_row = -1, first = 0, nb = 0;
    bool ok = true;
    const uint64_t pol_keep = make_policy_evict_last();
#pragma unroll
    for (int k = 0; k < kRangePerThread; ++k) {
        if (xs[k] == kEmpty) continue;  // not ours, or an id the uint32 slots cannot hold
        const int64_t r = r0 + head[head_at(k * kBuildThreads + tid)] - 1;
        if (r != cur_row) {
            cur_row = r;
            const int64_t b = ldg_idx(a.row_ptr + r), en = ldg_idx(a.row_ptr + r + 1);
            if (en - b >= kMinTableDeg) table_span(b, en, first, nb); else nb = 0;
        }
        // a short row is a single segment: probing wraps over the whole row, capacity is guaranteed
        if (nb > 0) ok &= smem_insert(image + (first - bucket0) * 8, count + (first - bucket0), nb, home_bucket(xs[k], nb), xs[k]);
    }
    if (!ok) *a.failed = 1;
    __syncthreads();
    // Write the finished buckets of the owned rows: [ceil(s_lo/4), floor(s_hi/4)).
    const int64_t w_lo = ((int64_t)s_lo + 3) >> 2, w_hi = own_end >> 2;
    const int64_t n16 = (w_hi - w_lo) * 2;  // 16-byte pieces
    const uint4* src = reinterpret_cast<const uint4*>(image + (w_lo - bucket0) * 8);
    uint4* dst = reinterpret_cast<uint4*>(a.table + w_lo * 8);
    for (int64_t i = tid; i < n16; i += kBuildThreads) dst[i] = src[i];
}

// Hub rows: persistent CTAs (one per SM: the segment image takes 144 KB of shared memory), which
// pull segments from a shared counter so that the few very long rows do not unbalance the grid.
template <bool WIDE>
__global__ void __launch_bounds__(kHubThreads, 1) build_hub_kernel(const BuildArgs a) {
    const IdxPtrT<WIDE> col_idx(a.col_idx);
    extern __shared__ __align__(16) uint32_t hub_image[];  // up to 2*kSegBuckets-1 buckets
    uint32_t* hub_count = hub_image + 2 * kSegBuckets * 8;
    __shared__ unsigned long long s_next;
    const int tid = threadIdx.x;
    const int64_t n_segments = (int64_t)(*a.hub_counter & 0xFFFFFFFFull);
    if (n_segments > a.max_segs) return;  // flagged by hub_segments_kernel
    for (;;) {
        if (tid == 0) s_next = atomicAdd(a.next_segment, 1ull);
        __syncthreads();
        const int64_t c = (int64_t)s_next;
        if (c >= n_segments) break;
        const SegmentWork work = a.segments[c];
        const int64_t b = ldg_idx(a.row_ptr + work.row), e = ldg_idx(a.row_ptr + work.row + 1);
        int64_t first, nb;
        table_span(b, e, first, nb);
        const int64_t nseg = segment_count(nb);
        const int64_t lo = work.segment << kSegShift;
        const int64_t hi = (work.segment == nseg - 1) ? nb : lo + kSegBuckets;
        const int64_t size = hi - lo;
        for (int64_t i = tid; i < size * 2; i += kHubThreads)
            reinterpret_cast<uint4*>(hub_image)[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
        for (int64_t i = tid; i < size; i += kHubThreads) hub_count[i] = 0;
        __syncthreads();
        bool ok = true;
        const uint64_t pol_keep = make_policy_evict_last();
        // eight independent loads in flight per thread; the row comes from L2 after its first reader
        for (int64_t i = b + tid; i < e; i += 8 * kHubThreads) {
            uint32_t x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t idx = i + (int64_t)u * kHubThreads;
                x[u] = idx < e ? slot_id(ldg_idx(col_idx + idx)) : kEmpty;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (x[u] == kEmpty) continue;
                const int64_t home = home_bucket(x[u], nb);
                if (home >= lo && home < hi) {  // every entry of the row has its home in exactly one segment
                    ok &= smem_insert(hub_image, hub_count, size, home - lo, x[u]);
                }
            }
        }
        if (!ok) *a.failed = 1;
        __syncthreads();
        const uint4* src = reinterpret_cast<const uint4*>(hub_image);
        uint4* dst = reinterpret_cast<uint4*>(a.table + (first + lo) * 8);
        for (int64_t i = tid; i < size * 2; i += kHubThreads) dst[i] = src[i];
        __syncthreads();  // the image (and s_next) are reused by the next segment
    }
}

// Reference build (option build_mode = 0): every entry is inserted with a global CAS into a table
// cleared by cudaMemsetAsync.  Same placement rule, kept as the A/B baseline of the tiled build.
__device__ __forceinline__ bool global_insert(uint32_t* __restrict__ table, int64_t first, int64_t nb, uint32_t x) {
    int64_t bkt = home_bucket(x, nb), lo, hi;
    probe_segment(nb, bkt, lo, hi);
    for (int64_t probes = lo; probes < hi; ++probes) {
        uint32_t* slots = table + (first + bkt) * 8;
        // Snapshot the bucket, then claim the first EMPTY slot seen; a lost race just moves on.
        const uint4 lo4 = ld_relaxed_u32x4(slots), hi4 = ld_relaxed_u32x4(slots + 4);
        const uint32_t snap[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (snap[j] == x) return true;
            if (snap[j] == kEmpty) {
                const uint32_t old = atomicCAS(slots + j, kEmpty, x);
                if (old == kEmpty || old == x) return true;
            }
        }
        if (++bkt == hi) bkt = lo;
    }
    return false;
}

template <bool WIDE>
__global__ void __launch_bounds__(kBuildThreads, 4) build_flat_kernel(const BuildArgs a) {
    const IdxPtrT<WIDE> col_idx(a.col_idx);
    __shared__ uint32_t head[kBuildTile];
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t e0 = tile * kBuildTile;
    const int64_t e1 = min(e0 + (int64_t)kBuildTile, a.nnz);
    const int64_t r0 = a.tile_row0[tile];
    const int64_t r1 = min(a.tile_row0[tile + 1], a.n_nodes - 1);
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) head[k * kBuildThreads + tid] = 0;
    __syncthreads();
    for (int64_t r = r0 + 1 + tid; r <= r1; r += kBuildThreads) {
        const int64_t b = ldg_idx(a.row_ptr + r), e = ldg_idx(a.row_ptr + r + 1);
        if (e > b && b < e1) head[b - e0] = (uint32_t)(r - r0);
    }
    __syncthreads();
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
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((tid & 31) >= d) incl = max(incl, o);
    }
    if ((tid & 31) == 31) warp_max[tid >> 5] = incl;
    __syncthreads();
    uint32_t before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if ((tid & 31) == 0) before = 0;
    for (int w = 0; w < (tid >> 5); ++w) before = max(before, warp_max[w]);

    const int64_t mine = e0 + (int64_t)tid * kBuildPerThread;
    int64_t cur_row = -1, first = 0, nb = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        if (mine + k >= e1) break;
        const int64_t r = r0 + max(before, own[k]);
        if (r != cur_row) {
            cur_row = r;
            const int64_t b = ldg_idx(a.row_ptr + r), en = ldg_idx(a.row_ptr + r + 1);
            if (en - b >= kMinTableDeg) table_span(b, en, first, nb); else nb = 0;
        }
        const uint32_t x = slot_id(ldg64_stream(col_idx + (mine + k)));
        if (nb > 0 && x != kEmpty) ok &= global_insert(a.table, first, nb, x);
    }
    if (!ok) *a.failed = 1;
}

// Are all rows strictly increasing (sorted, no edge stored twice)?  Two counts decide it without
// knowing which row an entry belongs to: descents anywhere in col_idx (col[i] >= col[i+1]) can only
// sit on row boundaries of such a graph, so the rows are strict iff the number of descents equals
// the number of descents found AT row boundaries.
template <bool WIDE>
__global__ void __launch_bounds__(256) strict_descents_kernel(IdxPtr col_any, int64_t nnz,
                                                              unsigned long long* counts) {
    const IdxPtrT<WIDE> col_idx(col_any);
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    unsigned long long n = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < nnz; i += gsz)
        n += ldg64_stream(col_idx + i) >= ldg_idx(col_idx + i + 1) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, d);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counts, n);
}
__global__ void __launch_bounds__(256) strict_boundaries_kernel(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes,
                                                                int64_t nnz, unsigned long long* counts) {
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    unsigned long long n = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_nodes; r += gsz) {
        const int64_t b = ldg_idx(row_ptr + r), e = ldg_idx(row_ptr + r + 1);
        if (e > b && e < nnz) n += ldg_idx(col_idx + e - 1) >= ldg_idx(col_idx + e) ? 1 : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, d);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counts + 1, n);
}

// Packed uint32 copies of the rows too short for a hashed table (member_table.cuh).  Runs after the
// table build, whose whole-bucket writes leave EMPTY over these rows' bytes; a short row's bytes are
// its own, so nothing of a neighbouring row's table is touched.
template <bool WIDE>
__global__ void __launch_bounds__(256) short_rows_kernel(IdxPtr row_ptr, IdxPtr col_any, int64_t n_nodes,
                                                         uint32_t* __restrict__ table) {
    const IdxPtrT<WIDE> col_idx(col_any);
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_nodes; r += gsz) {
        const int64_t b = ldg_idx(row_ptr + r), e = ldg_idx(row_ptr + r + 1);
        const int64_t d = e - b;
        if (d <= 0 || d >= kMinTableDeg) continue;
        uint32_t ids[kMinTableDeg];
#pragma unroll
        for (int k = 0; k < kMinTableDeg - 1; ++k) ids[k] = k < d ? slot_id(ldg64_stream(col_idx + b + k)) : kEmpty;
        uint32_t* words = table + 2 * b;
#pragma unroll
        for (int k = 0; k < 2 * (kMinTableDeg - 1); ++k)
            if (k < 2 * d) words[k] = k < kMinTableDeg - 1 ? ids[k] : kEmpty;
    }
}

// Edge records (member_table.cuh): one streaming pass over col_idx; the two row-index reads per
// entry hit the L2-resident uint32 index (evict_last), the 16-byte records leave coalesced.
template <bool WIDE>
__global__ void __launch_bounds__(256) edge_records_kernel(IdxPtr col_any, int64_t nnz,
                                                           const uint32_t* __restrict__ row32, int64_t n_nodes,
                                                           uint4* __restrict__ records) {
    const IdxPtrT<WIDE> col_idx(col_any);
    constexpr int kBatch = 4;  // independent entries per thread and round
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += kBatch * gsz) {
        int64_t x[kBatch];
        uint32_t b[kBatch], e[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) x[u] = i + u * gsz < nnz ? ldg64_stream(col_idx + i + u * gsz) : -1;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const bool inside = (uint64_t)x[u] < (uint64_t)n_nodes;
            b[u] = inside ? (uint32_t)ldg32_keep(row32 + x[u], pol_keep) : 0u;
            e[u] = inside ? (uint32_t)ldg32_keep(row32 + x[u] + 1, pol_keep) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (i + u * gsz >= nnz) continue;
            const uint64_t id = (uint64_t)x[u];
            const uint32_t deg = e[u] - b[u];
            stg_u32x4_hint(records + i + u * gsz, (uint32_t)id, deg, b[u], record_last_word(id, deg, kBloomAll), pol_stream);
        }
    }
}

// Triangle Blooms of a kept graph (member_table.cuh): word .w of record (t -> v) becomes the 32-bit Bloom
// of adj(t) & adj(v).  The set is the same from either end, so each unordered pair is worked out once, by
// the entry that sits in the LONGER row (ties: the larger id): it walks the shorter row adj(v), asks row
// t's table about every w and, on the way, meets t itself -- at the position of the
// mirror entry (v -> t), which receives the same word.  A warp owns 32 consecutive CSR entries; the
// (entry, neighbour) pairs of the tile are flattened over the lanes, two per lane and round, so short
// rows do not idle the warp and two gathers are in flight per lane.  Pairs whose shorter row exceeds
// `cap` keep the saturated word: their Bloom would be full anyway and the work is quadratic in hub size.
// The same pass proves or refutes that the graph is symmetric (every (t -> v) has its (v -> t)), which
// the walk needs before it may look at a triangle from its far side.
//
// Which row an entry belongs to is a search of the row index (24 dependent L2 reads on c3).  A warp therefore takes runs
// of kBloomRun consecutive tiles: the first tile of a run bisects, the others gallop forward from the row where the
// previous tile ended (a read or two).
constexpr int kBloomWarps = 8;
constexpr int kBloomUnroll = 2;
constexpr int kBloomRun = 8;
template <bool WIDE>
__global__ void __launch_bounds__(kBloomWarps * 32) edge_bloom_kernel(IdxPtr col_any, int64_t nnz,
                                                                      const uint32_t* __restrict__ row32, int64_t n_nodes,
                                                                      uint4* __restrict__ records,
                                                                      const uint32_t* __restrict__ table,
                                                                      uint32_t* __restrict__ hub_bits, uint32_t hub_n_bits,
                                                                      uint32_t cap, const int* __restrict__ table_failed,
                                                                      int* __restrict__ asymmetric) {
    const IdxPtrT<WIDE> col_idx(col_any);
    __shared__ uint32_t s_bloom[kBloomWarps][32];
    __shared__ uint32_t s_mirror[kBloomWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const int64_t n_tiles = (nnz + 31) >> 5;
    constexpr uint32_t kNoMirror = 0xFFFFFFFFu;
    if (*table_failed != 0) {  // a table that overflowed cannot be asked: the words stay saturated, symmetry unproven
        if (threadIdx.x == 0) *asymmetric = 1;
        return;
    }
    const int64_t n_runs = (n_tiles + kBloomRun - 1) / kBloomRun;
    for (int64_t run = (int64_t)blockIdx.x * kBloomWarps + warp; run < n_runs; run += (int64_t)gridDim.x * kBloomWarps) {
      int64_t row_from = -1;  // a row at or before the rows of the next tile (-1: not known yet)
      for (int64_t tile = run * kBloomRun; tile < min((run + 1) * kBloomRun, n_tiles); ++tile) {
        const int64_t k = tile * 32 + lane;
        const bool valid = k < nnz;
        // this lane's entry: row t (a search of the L2-resident row index), neighbour v with its span from the record
        uint32_t t = 0, tb = 0, dt = 0, v = 0, vb = 0, dv = 0;
        if (valid) {
            int64_t lo = 0, hi = n_nodes;  // row32[lo] <= k < row32[hi]
            if (row_from >= 0) {  // gallop forward from where the previous tile ended
                lo = row_from;
                int64_t step = 1;
                while (lo + step < n_nodes && (int64_t)ldg32_keep(row32 + lo + step, pol_keep) <= k) { lo += step; step <<= 1; }
                hi = min(lo + step, n_nodes);
            }
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int64_t)ldg32_keep(row32 + mid, pol_keep) <= k) lo = mid; else hi = mid;
            }
            t = (uint32_t)lo;
            tb = (uint32_t)ldg32_keep(row32 + lo, pol_keep);
            dt = (uint32_t)ldg32_keep(row32 + lo + 1, pol_keep) - tb;
            const uint4 rec = records[k];
            v = rec.x; dv = rec.y; vb = rec.z;
        }
        const bool live = valid && dv > 0;  // a neighbour without out-edges (or outside the graph) keeps its word
        // an edge between two rows longer than the cap keeps its saturated word: it goes into the hub-pair filter instead
        if (hub_bits != nullptr && live && dt > cap && dv > cap) filter_insert(hub_bits, hub_n_bits, t, v, pol_keep);
        const bool owner = live && (dt > dv || (dt == dv && t >= v));
        const bool exact = owner && dv <= cap;
        // every other live entry only has to know that its mirror exists
        if (live && !exact && !is_member<true>((int64_t)t, (int64_t)vb, (int64_t)vb + dv, col_idx, table, pol_stream)) *asymmetric = 1;
        s_bloom[warp][lane] = 0;
        s_mirror[warp][lane] = kNoMirror;
        const uint32_t items = exact ? dv : 0u;
        uint32_t incl = items;  // inclusive prefix of the work items over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        __syncwarp();
        for (uint32_t i0 = 0; i0 < total; i0 += 32 * kBloomUnroll) {
            int j[kBloomUnroll];
            uint32_t pos[kBloomUnroll], row[kBloomUnroll], bb[kBloomUnroll], bd[kBloomUnroll];
            int64_t w[kBloomUnroll];
            bool act[kBloomUnroll];
#pragma unroll
            for (int u = 0; u < kBloomUnroll; ++u) {
                const uint32_t i = i0 + u * 32 + lane;
                // entry j that owns item i: the first lane whose inclusive prefix exceeds i
                int jj = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xFFFFFFFFu, incl, jj + step - 1);
                    if (probe <= i) jj += step;
                }
                j[u] = jj;
                const uint32_t j_incl = __shfl_sync(0xFFFFFFFFu, incl, jj);
                const uint32_t j_items = __shfl_sync(0xFFFFFFFFu, items, jj);
                const uint32_t j_vb = __shfl_sync(0xFFFFFFFFu, vb, jj);
                row[u] = __shfl_sync(0xFFFFFFFFu, t, jj);
                bb[u] = __shfl_sync(0xFFFFFFFFu, tb, jj);
                bd[u] = __shfl_sync(0xFFFFFFFFu, dt, jj);
                act[u] = i < total;
                pos[u] = j_vb + (i - (j_incl - j_items));
                w[u] = act[u] ? ldg64_hint(col_idx + pos[u], pol_stream) : -1;
            }
            bool maybe[kBloomUnroll];
#pragma unroll
            for (int u = 0; u < kBloomUnroll; ++u) {
                if (act[u] && w[u] == (int64_t)row[u]) s_mirror[warp][j[u]] = pos[u];
                maybe[u] = act[u];
            }
#pragma unroll
            for (int u = 0; u < kBloomUnroll; ++u)
                // the table of the longer row is asked once per neighbour of every shorter row it owns, by the warps of all SMs at
                // about the same time: worth keeping in L2 (evict_last), unlike the one-off sectors of a walk
                if (maybe[u] && is_member<true>(w[u], (int64_t)bb[u], (int64_t)bb[u] + bd[u], col_idx, table, pol_keep))
                    atomicOr(&s_bloom[warp][j[u]], bloom_bit(w[u]));
        }
        __syncwarp();
        if (exact) {
            const uint32_t word = s_bloom[warp][lane], mirror = s_mirror[warp][lane];
            reinterpret_cast<uint32_t*>(records + k)[3] = word;
            if (mirror != kNoMirror) reinterpret_cast<uint32_t*>(records + mirror)[3] = word;
            else *asymmetric = 1;
        }
        row_from = (int64_t)__shfl_sync(0xFFFFFFFFu, t, 31);  // (the last tile's lanes past nnz hold 0, and no tile follows it)
        __syncwarp();
      }
    }
}

// ------------------------------------------------------------------------------------------ host
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

CsrWorkspace csr_workspace_layout(int64_t n_nodes, int64_t nnz, bool uniform, bool records, bool filter) {
    CsrWorkspace w{};
    const bool ids_fit = (uint64_t)n_nodes < 0xFFFFFFFFull;   // neighbour ids must fit the uint32 table slots
    const bool offsets_fit = (uint64_t)nnz <= 0xFFFFFFFFull;   // row offsets must fit the uint32 row index
    w.has_table = !uniform && ids_fit && nnz > 0;
    w.has_row32 = offsets_fit && n_nodes > 0;
    w.n_tiles = (nnz + kBuildTile - 1) / kBuildTile;
    w.n_buckets = (nnz + 3) / 4;
    w.max_hubs = nnz / kHubDeg + 1;
    w.max_segs = w.n_buckets / kSegBuckets + w.max_hubs + 1;  // a hub has one segment, or segments of >= kSegBuckets buckets
    size_t off = 0;
    w.table = off;
    if (w.has_table) off += align256((size_t)w.n_buckets * 32);
    w.tile_row0 = off;
    if (w.has_table) off += align256((size_t)(w.n_tiles + 1) * 8);
    w.hub_list = off;
    if (w.has_table) off += align256((size_t)w.max_hubs * sizeof(HubEntry));
    w.seg_work = off;
    if (w.has_table) off += align256((size_t)w.max_segs * sizeof(SegmentWork));
    w.cells = off;
    if (nnz > 0) off += 256;
    w.row32 = off;
    if (w.has_row32) off += align256((size_t)(n_nodes + 1) * 4);
    w.has_records = records && w.has_row32 && nnz > 0;  // spans are stored as uint32 (start, degree)
    w.records = off;
    if (w.has_records) off += align256((size_t)nnz * 16);
    // hub-pair filter: 32 bits per CSR entry for small graphs, capped at edge_filter_mb (it has to stay L2-resident)
    w.filter = off;
    w.filter_bits = 0;
    const int64_t filter_mb = options().edge_filter_mb > 511 ? 511 : options().edge_filter_mb;
    if (w.has_table && filter && filter_mb > 0) {
        const size_t bytes = std::min((size_t)filter_mb << 20, align256((size_t)nnz * 4));
        w.filter_bits = (uint32_t)(bytes * 8);
        off += bytes;
    }
    w.total = off;
    return w;
}

int csr_prepare_device(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, void* workspace,
                       const CsrWorkspace& w, bool want_table, bool want_row32, bool want_strict, bool want_records,
                       int build_mode, int device, cudaStream_t st, CsrPrepared* out, int64_t bloom_cap) {
    want_table = want_table && w.has_table;
    want_row32 = want_row32 && w.has_row32;
    want_records = want_records && w.has_records && want_row32;
    want_strict = want_strict && nnz > 0;
    *out = CsrPrepared{};
    if (!want_table && !want_row32 && !want_strict) return TRW_OK;
    char* ws = (char*)workspace;
    BuildArgs b{};
    b.row_ptr = row_ptr; b.col_idx = col_idx; b.n_nodes = n_nodes; b.nnz = nnz;
    b.n_tiles = w.n_tiles; b.n_buckets = w.n_buckets; b.max_hubs = w.max_hubs; b.max_segs = w.max_segs;
    b.row32 = want_row32 ? (uint32_t*)(ws + w.row32) : nullptr;
    const bool want_filter = want_table && build_mode != 0 && w.filter_bits != 0;  // (filled by csr_add_blooms)
    int rc;
    if (want_table) {
        b.table = (uint32_t*)(ws + w.table);
        b.tile_row0 = (int64_t*)(ws + w.tile_row0);
        b.hub_counter = (unsigned long long*)(ws + w.cells);
        b.failed = (int*)(ws + w.cells + 64);
        b.next_segment = (unsigned long long*)(ws + w.cells + 32);
        if (build_mode != 0) {
            b.hubs = (HubEntry*)(ws + w.hub_list);
            b.segments = (SegmentWork*)(ws + w.seg_work);
        }
    }
    if (want_table || want_strict) {
        rc = check_cuda(cudaMemsetAsync(ws + w.cells, 0, 256, st), "cells memset");
        if (rc) return rc;
    }
    const int sms = sm_count(device);
    if (want_strict) {
        unsigned long long* counts = (unsigned long long*)(ws + w.cells + 128);
        if (col_idx.wide()) strict_descents_kernel<true><<<sms * 8, 256, 0, st>>>(col_idx, nnz, counts);
        else strict_descents_kernel<false><<<sms * 8, 256, 0, st>>>(col_idx, nnz, counts);
        strict_boundaries_kernel<<<sms * 8, 256, 0, st>>>(row_ptr, col_idx, n_nodes, nnz, counts);
        count_launch(2);
        rc = check_cuda(cudaGetLastError(), "strict-rows check launch");
        if (rc) return rc;
        out->strict_counts = counts;
    }
    if (!want_table && !want_row32) return TRW_OK;
    csr_prepass_kernel<<<sms * 8, 256, 0, st>>>(b);
    count_launch(1);
    rc = check_cuda(cudaGetLastError(), "csr_prepass launch");
    if (rc) return rc;
    if (want_table) {
        if (build_mode == 0) {
            rc = check_cuda(cudaMemsetAsync(b.table, 0xFF, (size_t)w.n_buckets * 32, st), "table memset");
            if (rc) return rc;
            if (col_idx.wide()) build_flat_kernel<true><<<(unsigned)w.n_tiles, kBuildThreads, 0, st>>>(b);
            else build_flat_kernel<false><<<(unsigned)w.n_tiles, kBuildThreads, 0, st>>>(b);
            count_launch(1);
        } else {
            static bool attr_set[64];
            if (device >= 64 || !attr_set[device]) {
                rc = check_cuda(cudaFuncSetAttribute(build_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)kTiledSmemBytes), "tiled build smem attribute");
                if (!rc) rc = check_cuda(cudaFuncSetAttribute(build_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                              (int)kTiledSmemBytes), "tiled build smem attribute");
                if (!rc) rc = check_cuda(cudaFuncSetAttribute(build_hub_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                              (int)kHubSmemBytes), "hub build smem attribute");
                if (!rc) rc = check_cuda(cudaFuncSetAttribute(build_hub_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                              (int)kHubSmemBytes), "hub build smem attribute");
                if (rc) return rc;
                if (device < 64) attr_set[device] = true;
            }
            hub_segments_kernel<<<sms, 256, 0, st>>>(b);
            if (col_idx.wide()) {
                build_tiled_kernel<true><<<(unsigned)w.n_tiles, kBuildThreads, kTiledSmemBytes, st>>>(b);
                build_hub_kernel<true><<<sms, kHubThreads, kHubSmemBytes, st>>>(b);
            } else {
                build_tiled_kernel<false><<<(unsigned)w.n_tiles, kBuildThreads, kTiledSmemBytes, st>>>(b);
                build_hub_kernel<false><<<sms, kHubThreads, kHubSmemBytes, st>>>(b);
            }
            count_launch(3);
        }
        if (col_idx.wide()) short_rows_kernel<true><<<sms * 8, 256, 0, st>>>(row_ptr, col_idx, n_nodes, b.table);
        else short_rows_kernel<false><<<sms * 8, 256, 0, st>>>(row_ptr, col_idx, n_nodes, b.table);
        count_launch(1);
        rc = check_cuda(cudaGetLastError(), "membership table build launch");
        if (rc) return rc;
        out->table = b.table;
        out->table_failed = b.failed;
        if (want_filter) { out->filter_space = (uint32_t*)(ws + w.filter); out->filter_space_bits = w.filter_bits; }
    }
    out->row32 = b.row32;
    if (want_records) {
        uint4* records = (uint4*)(ws + w.records);
        if (col_idx.wide()) edge_records_kernel<true><<<sms * 16, 256, 0, st>>>(col_idx, nnz, b.row32, n_nodes, records);
        else edge_records_kernel<false><<<sms * 16, 256, 0, st>>>(col_idx, nnz, b.row32, n_nodes, records);
        count_launch(1);
        rc = check_cuda(cudaGetLastError(), "edge records launch");
        if (rc) return rc;
        out->records = records;
        out->bloom_flag = (int*)(ws + w.cells + 192);
        if (bloom_cap > 0 && build_mode != 0) return csr_add_blooms(out, col_idx, n_nodes, nnz, bloom_cap, device, st);
    }
    return TRW_OK;
}

int csr_add_blooms(CsrPrepared* pr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, int64_t cap, int device, cudaStream_t st) {
    if (pr->asymmetric != nullptr || cap <= 0) return TRW_OK;  // already there
    if (!pr->table || !pr->records || !pr->row32 || !pr->bloom_flag || nnz <= 0) return TRW_OK;  // nothing to hang them on
    int rc = check_cuda(cudaMemsetAsync(pr->bloom_flag, 0, sizeof(int), st), "bloom flag memset");
    if (rc) return rc;
    const unsigned grid = (unsigned)sm_count(device) * 8;
    uint4* records = const_cast<uint4*>(pr->records);
    const uint32_t cap32 = (uint32_t)std::min<int64_t>(cap, 1 << 20);
    if (pr->filter_space) {
        rc = check_cuda(cudaMemsetAsync(pr->filter_space, 0, (size_t)pr->filter_space_bits / 8, st), "hub-pair filter memset");
        if (rc) return rc;
    }
    if (col_idx.wide())
        edge_bloom_kernel<true><<<grid, kBloomWarps * 32, 0, st>>>(col_idx, nnz, pr->row32, n_nodes, records, pr->table, pr->filter_space,
                                                                  pr->filter_space_bits, cap32, pr->table_failed, pr->bloom_flag);
    else
        edge_bloom_kernel<false><<<grid, kBloomWarps * 32, 0, st>>>(col_idx, nnz, pr->row32, n_nodes, records, pr->table, pr->filter_space,
                                                                   pr->filter_space_bits, cap32, pr->table_failed, pr->bloom_flag);
    count_launch(1);
    rc = check_cuda(cudaGetLastError(), "edge bloom launch");
    if (rc) return rc;
    pr->asymmetric = pr->bloom_flag;
    // (a pass that found the table overflowed returns at once and leaves the filter empty: the walk kernels do not consult
    // it then, they scan -- walk_csr.cu)
    if (pr->filter_space) pr->filter = EdgeFilter{pr->filter_space, pr->filter_space_bits, cap32};
    return TRW_OK;
}

}  // namespace trw
