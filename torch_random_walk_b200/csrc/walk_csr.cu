// CSR random walks for sm_100a: first-order (uniform) and second-order (node2vec) kernels.
//
// Replaces the reference's walk_gpu / uniform_walk_gpu / biased_walk_gpu
// (csrc/cuda/rw_cuda.cu:186-248 / 59-98 / 100-184).  Semantics kept:
//   * out[i,0] = targets[i]; every later entry is drawn from adj(previous entry);
//   * a node without out-edges keeps the walk where it is (rw_cuda.cu:25-30);
//   * node2vec: step 1 uniform, then propose x ~ U(adj(v)), accept with probability
//     prob_0 (x == t), prob_1 (x in adj(t)), prob_2 (otherwise), prob_k = {1/p,1,1/q}/max
//     (rw_cuda.cu:119-123, 146-179).
//
// Design.  Both kernels are dependent random gathers whose cost is the
// number of random DRAM fetches per step, so
//   * one thread owns one walk and the SMs are kept full of them;
//   * the node2vec loop is flattened to one rejection *trial* per iteration, so lanes whose
//     proposal was accepted move on instead of idling until the slowest lane is accepted;
//   * "x in adj(t)" is one 32-byte sector of a hashed copy of the adjacency that is built per
//     call into caller-provided workspace, not a scan or search of adj(t);
//   * row_ptr is re-encoded per call as uint32 offsets (half the bytes, fits the 126 MB L2 for
//     16 M nodes) and read with an L2 evict_last policy, while the never-reused gathers
//     (proposals, table buckets) and the output stores carry evict_first, so the stream of
//     random sectors does not wash the row index out of L2.
#include <cooperative_groups.h>

#include "trw_common.cuh"
#include "trw_options.h"
#include "walk_csr.h"

namespace cg = cooperative_groups;

namespace trw {

// ------------------------------------------------------------------------------------------
// Membership table.  Row t of the CSR owns the bytes [8*row_ptr[t], 8*row_ptr[t+1]) of a
// table as large as col_idx; the 32-byte buckets wholly inside that span hold the row's
// neighbour ids as uint32 (8 slots per bucket, EMPTY = 0xFFFFFFFF), open addressing over
// buckets.  A row of degree d >= kMinTableDeg owns >= (d-6)/4 buckets = 2d-12 >= d slots, so
// inserts always find room; shorter rows are scanned directly (<= 11 ids).  No per-row
// pointer is needed: the bucket range follows from the row_ptr pair the walk already holds.
// Lookup: hash -> bucket -> one sector; a hit, or any EMPTY slot (the bucket never
// overflowed), ends the probe.
// ------------------------------------------------------------------------------------------
constexpr int64_t kMinTableDeg = 12;
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

struct BuildArgs {
    const int64_t* row_ptr;
    const int64_t* col_idx;
    int64_t n_nodes, nnz;
    uint32_t* table;
    int64_t* tile_row0;          // [n_tiles + 1]: row holding the first entry of each tile
    unsigned long long* maxdeg;  // largest row length (decides how far ahead buckets are cleared)
    uint32_t* row32;             // optional compact copy of row_ptr
    int64_t n_tiles, n_buckets;
    int chunk_tiles;             // tiles per L2-resident chunk
};

// Pre-pass (ordinary launch): first row of every tile, the maximum degree, and the uint32 copy
// of row_ptr.  All three are embarrassingly parallel.
__global__ void __launch_bounds__(256) csr_prepass_kernel(const BuildArgs a) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    if (a.tile_row0) {
        for (int64_t j = gtid; j <= a.n_tiles; j += gsz) {
            const int64_t e = j * kBuildTile;
            a.tile_row0[j] = e < a.nnz ? row_of_entry(a.row_ptr, a.n_nodes, e) : a.n_nodes;
        }
    }
    unsigned long long best = 0;
    for (int64_t r = gtid; r <= a.n_nodes; r += gsz) {
        const int64_t b = __ldg(a.row_ptr + r);
        if (a.row32) a.row32[r] = (uint32_t)b;
        if (a.maxdeg && r < a.n_nodes) best = max(best, (unsigned long long)(__ldg(a.row_ptr + r + 1) - b));
    }
    if (a.maxdeg) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
        if ((threadIdx.x & 31) == 0 && best) atomicMax(a.maxdeg, best);
    }
}

__device__ __forceinline__ void clear_buckets(const BuildArgs& a, int64_t lo, int64_t hi, int64_t gtid, int64_t gsz) {
    for (int64_t b = lo + gtid; b < hi; b += gsz) stg_sector(a.table + b * 8, ~0ull, ~0ull, ~0ull, ~0ull);
}

__device__ __forceinline__ void table_insert(uint32_t* __restrict__ table, int64_t first, int64_t nb, uint32_t x) {
    int64_t bkt = (int64_t)__umul64hi((uint64_t)mix32(x) << 32, (uint64_t)nb);
    for (;;) {
        uint32_t* slots = table + (first + bkt) * 8;
        // Snapshot the bucket, then claim the first EMPTY slot seen; a lost race just moves on.
        const uint4 lo4 = ld_relaxed_u32x4(slots), hi4 = ld_relaxed_u32x4(slots + 4);
        const uint32_t snap[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (snap[j] == x) return;
            if (snap[j] == kEmpty) {
                const uint32_t old = atomicCAS(slots + j, kEmpty, x);
                if (old == kEmpty || old == x) return;
            }
        }
        if (++bkt == nb) bkt = 0;
    }
}

// One tile = kBuildTile consecutive CSR entries.  The rows they belong to are recovered with a
// shared-memory max-scan over "a row starts here" marks (edge-balanced: hubs and short rows cost
// the same per entry), then every entry is inserted into its row's buckets.
__device__ __forceinline__ void build_tile(const BuildArgs& a, int64_t tile, uint32_t* head, uint32_t* warp_max) {
    const int tid = threadIdx.x;
    const int64_t e0 = tile * kBuildTile;
    const int64_t e1 = min(e0 + (int64_t)kBuildTile, a.nnz);
    const int64_t r0 = a.tile_row0[tile];
    const int64_t r1 = min(a.tile_row0[tile + 1], a.n_nodes - 1);
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) head[k * kBuildThreads + tid] = 0;
    __syncthreads();
    for (int64_t r = r0 + 1 + tid; r <= r1; r += kBuildThreads) {
        const int64_t b = __ldg(a.row_ptr + r), e = __ldg(a.row_ptr + r + 1);
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

    // This thread's eight consecutive entries: fetch them first (two 256-bit loads when the
    // span is whole and aligned) so that the inserts below do not wait on them one by one.
    const int64_t mine = e0 + (int64_t)tid * kBuildPerThread;
    uint64_t x[kBuildPerThread];
    if (mine + kBuildPerThread <= e1 && (((uintptr_t)(a.col_idx + mine)) & 31) == 0) {
        const Sector64 s0 = ldg_sector(a.col_idx + mine), s1 = ldg_sector(a.col_idx + mine + 4);
        x[0] = s0.a; x[1] = s0.b; x[2] = s0.c; x[3] = s0.d;
        x[4] = s1.a; x[5] = s1.b; x[6] = s1.c; x[7] = s1.d;
    } else {
#pragma unroll
        for (int k = 0; k < kBuildPerThread; ++k) x[k] = mine + k < e1 ? (uint64_t)ldg64_stream(a.col_idx + mine + k) : 0;
    }
    int64_t cur_row = -1, first = 0, nb = 0;
#pragma unroll
    for (int k = 0; k < kBuildPerThread; ++k) {
        if (mine + k >= e1) break;
        const int64_t r = r0 + max(before, own[k]);
        if (r != cur_row) {
            cur_row = r;
            const int64_t b = __ldg(a.row_ptr + r), en = __ldg(a.row_ptr + r + 1);
            if (en - b >= kMinTableDeg) table_span(b, en, first, nb); else nb = 0;
        }
        if (nb > 0) table_insert(a.table, first, nb, (uint32_t)x[k]);
    }
    __syncthreads();  // head/warp_max are reused by the next tile
}

// Table build, cooperative and persistent.  The table is filled in chunks small enough to stay
// in L2: the CTAs clear the buckets of the chunk ahead with full-sector stores, synchronise the
// grid, then insert the current chunk's entries with L2-resident atomics, so each table byte
// goes to DRAM once (write-back) instead of once for a memset and again for every random CAS.
// Inserts of chunk k can reach at most max-degree entries past the chunk (the row of its last
// entry), which is how far ahead buckets are cleared before the chunk starts.
__global__ void __launch_bounds__(kBuildThreads, 4) build_member_table_coop_kernel(const BuildArgs a) {
    __shared__ uint32_t head[kBuildTile];
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    cg::grid_group grid = cg::this_grid();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    const int64_t chunk_entries = (int64_t)a.chunk_tiles * kBuildTile;
    const int64_t chunk_buckets = chunk_entries / 4;
    const int64_t n_chunks = (a.n_tiles + a.chunk_tiles - 1) / a.chunk_tiles;
    const int64_t ahead = (int64_t)(*a.maxdeg / (unsigned long long)chunk_entries) + 1;

    clear_buckets(a, 0, min((ahead + 1) * chunk_buckets, a.n_buckets), gtid, gsz);
    grid.sync();
    for (int64_t k = 0; k < n_chunks; ++k) {
        const int64_t t_end = min((k + 1) * a.chunk_tiles, a.n_tiles);
        for (int64_t tile = k * a.chunk_tiles + blockIdx.x; tile < t_end; tile += gridDim.x) build_tile(a, tile, head, warp_max);
        const int64_t c = k + 1 + ahead;
        if (c * chunk_buckets < a.n_buckets) clear_buckets(a, c * chunk_buckets, min((c + 1) * chunk_buckets, a.n_buckets), gtid, gsz);
        grid.sync();
    }
}

// Same build as one ordinary launch over a table cleared by cudaMemsetAsync (used when a
// cooperative launch is not possible, and as the A/B baseline: option build_mode = 0).
__global__ void __launch_bounds__(kBuildThreads, 4) build_member_table_flat_kernel(const BuildArgs a) {
    __shared__ uint32_t head[kBuildTile];
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    build_tile(a, blockIdx.x, head, warp_max);
}

// ------------------------------------------------------------------------------------------
// Table build, tiled through shared memory (the default).  Global atomics turned out to be the
// bound of the builds above (about 37 G CAS/s on B200, whether or not the lines are L2-resident),
// so they are kept for hub rows only.  A CTA owns the rows that START inside its tile of
// kBuildTile CSR entries and are shorter than kHubDeg: it assembles their buckets in shared
// memory with shared-memory CAS and writes them out with coalesced 16-byte stores -- no global
// atomic, no read-modify-write of table lines.  Such a row ends less than kHubDeg entries past
// the tile, so the CTA works on a window of kBuildRange entries.  Entries of hub rows
// (>= kHubDeg neighbours) inside the tile go through the global insert; their buckets were
// cleared by the memset that precedes the kernel.
// ------------------------------------------------------------------------------------------
constexpr int kHubDeg = kBuildTile;
constexpr int kBuildRange = kBuildTile + kHubDeg;
constexpr int kRangePerThread = kBuildRange / kBuildThreads;
constexpr uint32_t kNotOurs = 0xFFFFFFFFu;
constexpr size_t kTiledSmemBytes = (size_t)kBuildRange * 4 + (size_t)kBuildRange * 2 * 4;  // head + bucket image

__device__ __forceinline__ void smem_table_insert(uint32_t* __restrict__ stab, int64_t local_first, int64_t nb, uint32_t x) {
    const uint32_t h = mix32(x);
    int64_t bkt = (int64_t)__umul64hi((uint64_t)h << 32, (uint64_t)nb);
    for (int64_t probes = 0; probes < nb; ++probes) {
        uint32_t* slots = stab + (local_first + bkt) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t old = atomicCAS(slots + j, kEmpty, x);
            if (old == kEmpty || old == x) return;
        }
        if (++bkt == nb) bkt = 0;
    }
}

__global__ void __launch_bounds__(kBuildThreads, 4) build_member_table_tiled_kernel(const BuildArgs a) {
    extern __shared__ __align__(16) uint32_t tiled_smem[];
    uint32_t* head = tiled_smem;               // [kBuildRange] row code of each entry of the window
    uint32_t* stab = tiled_smem + kBuildRange; // [kBuildRange / 4 buckets][8 slots]
    __shared__ uint32_t warp_max[kBuildThreads / 32];
    __shared__ unsigned long long s_lo, s_hi;  // entry span [s_lo, s_hi) of the rows built in shared memory
    __shared__ long long s_hub_row;            // the hub row that starts in this tile, if any

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t e0 = tile * kBuildTile;
    const int64_t e1 = min(e0 + (int64_t)kBuildTile, a.nnz);
    const int64_t r0 = a.tile_row0[tile];
    const int64_t r1 = min(a.tile_row0[tile + 1], a.n_nodes - 1);
    const int64_t bucket0 = e0 >> 2;  // first bucket of the window (tiles are multiples of four entries)
#pragma unroll
    for (int k = 0; k < kRangePerThread; ++k) head[k * kBuildThreads + tid] = 0;
#pragma unroll
    for (int k = 0; k < 2 * kRangePerThread; ++k) stab[k * kBuildThreads + tid] = kEmpty;
    if (tid == 0) { s_lo = ~0ull; s_hi = 0; s_hub_row = -1; }
    __syncthreads();
    // Rows that start in [e0, e1): short ones get the code r - r0 + 1, a hub gets kNotOurs (nothing
    // else can start after it inside the tile: it is at least as long as the tile).
    for (int64_t r = r0 + tid; r <= r1; r += kBuildThreads) {
        const int64_t b = __ldg(a.row_ptr + r), e = __ldg(a.row_ptr + r + 1);
        if (e > b && b >= e0 && b < e1) {
            if (e - b >= kHubDeg) {
                head[b - e0] = kNotOurs;
                s_hub_row = r;
            } else {
                head[b - e0] = (uint32_t)(r - r0 + 1);
                atomicMin(&s_lo, (unsigned long long)b);
                atomicMax(&s_hi, (unsigned long long)e);
            }
        }
    }
    __syncthreads();
    // Entries past the last owned row belong to rows of later tiles.
    if (tid == 0 && s_hi > 0 && (int64_t)s_hi - e0 < kBuildRange) head[(int64_t)s_hi - e0] = kNotOurs;
    __syncthreads();
    // Inclusive max-scan of the codes; each thread scans kRangePerThread consecutive entries.
    {
        uint32_t own[kRangePerThread];
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kRangePerThread; ++k) {
            run = max(run, head[tid * kRangePerThread + k]);
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
        for (int k = 0; k < kRangePerThread; ++k) head[tid * kRangePerThread + k] = max(before, own[k]);
    }
    __syncthreads();
    // Row r0 reaches into the tile from the left when it did not start here; it is a hub's job
    // only if it is a hub (otherwise the tile where it starts builds it).
    const int64_t r0_b = __ldg(a.row_ptr + r0), r0_e = __ldg(a.row_ptr + r0 + 1);
    const bool r0_hub = r0_b < e0 && (r0_e - r0_b) >= kHubDeg;
    const int64_t hub_row = s_hub_row;
    int64_t hub_first = 0, hub_nb = 0, r0_first = 0, r0_nb = 0;
    if (hub_row >= 0) table_span(__ldg(a.row_ptr + hub_row), __ldg(a.row_ptr + hub_row + 1), hub_first, hub_nb);
    if (r0_hub) table_span(r0_b, r0_e, r0_first, r0_nb);
    const int64_t own_end = s_hi;  // 0 when no short row starts here
    int64_t cur_row = -1, first = 0, nb = 0;
    for (int k = 0; k < kRangePerThread; ++k) {
        const int idx = k * kBuildThreads + tid;  // strided: coalesced loads, hub work spread over all threads
        const int64_t e = e0 + idx;
        if (e >= a.nnz) break;
        const uint32_t code = head[idx];
        if (code == 0) {
            if (e < e1 && r0_hub) table_insert(a.table, r0_first, r0_nb, (uint32_t)ldg64_stream(a.col_idx + e));
        } else if (code == kNotOurs) {
            if (e < e1 && hub_row >= 0) table_insert(a.table, hub_first, hub_nb, (uint32_t)ldg64_stream(a.col_idx + e));
        } else if (e < own_end) {
            const int64_t r = r0 + code - 1;
            if (r != cur_row) {
                cur_row = r;
                const int64_t b = __ldg(a.row_ptr + r), en = __ldg(a.row_ptr + r + 1);
                if (en - b >= kMinTableDeg) table_span(b, en, first, nb); else nb = 0;
            }
            if (nb > 0) smem_table_insert(stab, first - bucket0, nb, (uint32_t)ldg64_stream(a.col_idx + e));
        }
    }
    __syncthreads();
    // Write the finished buckets of the owned rows: [ceil(s_lo/4), floor(s_hi/4)).
    if (own_end > 0) {
        const int64_t w_lo = ((int64_t)s_lo + 3) >> 2, w_hi = own_end >> 2;
        const int64_t n16 = (w_hi - w_lo) * 2;  // 16-byte pieces
        const uint4* src = reinterpret_cast<const uint4*>(stab + (w_lo - bucket0) * 8);
        uint4* dst = reinterpret_cast<uint4*>(a.table + w_lo * 8);
        for (int64_t i = tid; i < n16; i += kBuildThreads) dst[i] = src[i];
    }
}

// x in adj(t)?  (b,e) = row span of t.
template <bool TABLE>
__device__ __forceinline__ bool is_member(int64_t x, int64_t b, int64_t e, const int64_t* __restrict__ col_idx,
                                          const uint32_t* __restrict__ table, uint64_t pol_stream) {
    if (TABLE && e - b >= kMinTableDeg) {
        int64_t first, nb;
        table_span(b, e, first, nb);
        const uint32_t x32 = (uint32_t)x;
        int64_t bkt = (int64_t)__umul64hi((uint64_t)mix32(x32) << 32, (uint64_t)nb);
        for (int64_t probes = 0; probes < nb; ++probes) {
            const Sector64 s = ldg_sector_hint(table + (first + bkt) * 8, pol_stream);
            const uint32_t w[8] = {(uint32_t)s.a, (uint32_t)(s.a >> 32), (uint32_t)s.b, (uint32_t)(s.b >> 32),
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

// ------------------------------------------------------------------------------------------
template <bool ROW32>
__device__ __forceinline__ void load_row(const WalkArgs& a, int64_t v, int64_t& b, int64_t& e, uint64_t pol_keep) {
    if ((uint64_t)v < (uint64_t)a.n_nodes) {
        if (ROW32) {
            b = ldg32_keep(a.row32 + v, pol_keep);
            e = ldg32_keep(a.row32 + v + 1, pol_keep);
        } else {
            b = ldg64_keep(a.row_ptr + v, pol_keep);
            e = ldg64_keep(a.row_ptr + v + 1, pol_keep);
        }
    } else {
        b = e = 0;  // id outside the graph: treated as a node without out-edges
    }
}

// Neighbour of v at a uniformly random position, or v itself when it has none (rw_cuda.cu:8-31).
__device__ __forceinline__ int64_t pick_neighbor(const WalkArgs& a, int64_t v, int64_t b, int64_t e, uint32_t r0,
                                                 uint32_t r1, uint64_t pol_stream) {
    const int64_t deg = e - b;
    if (deg <= 0) return v;
    const int64_t idx = b + bounded(r0, r1, deg);
    if ((uint64_t)idx >= (uint64_t)a.nnz) return v;
    return ldg64_hint(a.col_idx + idx, pol_stream);
}

template <int BLOCK, bool STAGE>
struct RowOut {
    RowStager<BLOCK> st;
    int64_t* row;
    __device__ __forceinline__ void init(int64_t (*ring)[BLOCK], int64_t* r, int tid, uint64_t pol) {
        row = r;
        if (STAGE) st.init(ring, r, tid, pol);
    }
    __device__ __forceinline__ void put(int s, int64_t v, bool last) {
        if (STAGE) st.put(s, v, last); else row[s] = v;
    }
};

// First-order walk: one thread per walk, two dependent gathers per step (row span, then the
// chosen col_idx entry), one Philox block per four steps.
template <int BLOCK, bool STAGE, bool ROW32>
__global__ void __launch_bounds__(BLOCK, 8) uniform_walk_kernel(const WalkArgs a) {
    __shared__ int64_t ring[STAGE ? 4 : 1][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowOut<BLOCK, STAGE> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x, pol_stream);

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
        load_row<ROW32>(a, v, b, e, pol_keep);
        v = pick_neighbor(a, v, b, e, r, r * 0x9E3779B1u + (uint32_t)s, pol_stream);
        o.put(s, v, s == L);
    }
}

// Second-order walk.  One iteration of the loop = one rejection trial of this thread's walk.
template <int BLOCK, int MIN_CTAS, bool STAGE, bool TABLE, bool SPECULATE, bool ROW32>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) node2vec_walk_kernel(const WalkArgs a) {
    __shared__ int64_t ring[STAGE ? 4 : 1][BLOCK];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n_walks) return;
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const uint64_t wid = (uint64_t)(a.walk_id_offset + i);
    RowOut<BLOCK, STAGE> o;
    o.init(ring, a.out + i * a.out_row_stride, threadIdx.x, pol_stream);
    const int L = a.walk_length;
    const uint32_t wlo = (uint32_t)wid, whi = (uint32_t)(wid >> 32);

    int64_t t = __ldg(a.targets + i);
    o.put(0, t, L == 0);
    if (L == 0) return;
    int64_t tb, te;
    load_row<ROW32>(a, t, tb, te, pol_keep);
    uint4 rnd = philox4x32_10(make_uint4(wlo, whi, 1u, 0u), a.key);
    int64_t v = pick_neighbor(a, t, tb, te, rnd.x, rnd.z, pol_stream);  // first step is uniform (rw_cuda.cu:138)
    o.put(1, v, L == 1);
    if (L == 1) return;
    int64_t vb, ve;
    load_row<ROW32>(a, v, vb, ve, pol_keep);

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
        const int64_t x = pick_neighbor(a, v, vb, ve, r, r_hi, pol_stream);
        const bool back = (x == t);
        const bool possible = back ? (u < a.thr0) : (u < thr_far);
        int64_t xb = 0, xe = 0;
        if (SPECULATE && possible && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
        bool accept;
        if (u < thr_any) accept = true;
        else if (back) accept = u < a.thr0;
        else if (!possible) accept = false;
        else if (a.thr1 == a.thr2) accept = true;  // q == 1: membership cannot change the answer
        else accept = u < (is_member<TABLE>(x, tb, te, a.col_idx, a.table, pol_stream) ? a.thr1 : a.thr2);
        if (accept) {
            o.put(s, x, s == L);
            if (!SPECULATE && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
            t = v; tb = vb; te = ve;
            v = x; vb = xb; ve = xe;
            ++s;
            trial = 0;
        } else {
            ++trial;
        }
    }
}

// ------------------------------------------------------------------------------------------ launchers
template <bool STAGE, bool ROW32>
static void launch_uniform(const WalkArgs& a, cudaStream_t st) {
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    uniform_walk_kernel<BLOCK, STAGE, ROW32><<<grid, BLOCK, 0, st>>>(a);
}

template <int MIN_CTAS, bool STAGE, bool TABLE, bool ROW32>
static void launch_n2v3(const WalkArgs& a, bool speculate, cudaStream_t st) {
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    if (speculate) node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, true, ROW32><<<grid, BLOCK, 0, st>>>(a);
    else node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, false, ROW32><<<grid, BLOCK, 0, st>>>(a);
}

template <bool STAGE, bool TABLE, bool ROW32>
static void launch_n2v2(const WalkArgs& a, bool speculate, int min_ctas, cudaStream_t st) {
    if (min_ctas >= 6) launch_n2v3<6, STAGE, TABLE, ROW32>(a, speculate, st);
    else if (min_ctas == 5) launch_n2v3<5, STAGE, TABLE, ROW32>(a, speculate, st);
    else launch_n2v3<4, STAGE, TABLE, ROW32>(a, speculate, st);
}

static void launch_n2v(const WalkArgs& a, bool stage, bool table, bool row32, bool speculate, int min_ctas, cudaStream_t st) {
    if (!stage) {  // plain 8-byte stores: A/B path only, kept to one variant per table mode
        if (table) launch_n2v3<4, false, true, false>(a, speculate, st);
        else launch_n2v3<4, false, false, false>(a, speculate, st);
        return;
    }
    if (table) {
        if (row32) launch_n2v2<true, true, true>(a, speculate, min_ctas, st);
        else launch_n2v2<true, true, false>(a, speculate, min_ctas, st);
    } else {
        if (row32) launch_n2v2<true, false, true>(a, speculate, min_ctas, st);
        else launch_n2v2<true, false, false>(a, speculate, min_ctas, st);
    }
}

static uint64_t threshold(double prob) {
    double t = prob * 4294967296.0;
    if (!(t > 0.0)) return 0;
    if (t >= 4294967296.0) return 4294967296ull;
    return (uint64_t)t;
}

// Workspace layout (all offsets 256-byte aligned).
struct WsLayout {
    size_t table, tile_row0, maxdeg, row32, total;
    int64_t n_tiles, n_buckets;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static WsLayout ws_layout(int64_t n_nodes, int64_t nnz, bool uniform) {
    WsLayout w{};
    const bool ids_fit = (uint64_t)n_nodes < 0xFFFFFFFFull;    // neighbour ids must fit the uint32 table slots
    const bool offsets_fit = (uint64_t)nnz <= 0xFFFFFFFFull;    // row offsets must fit the uint32 row index
    const bool want_table = !uniform && ids_fit && nnz > 0;
    w.n_tiles = (nnz + kBuildTile - 1) / kBuildTile;
    w.n_buckets = (nnz + 3) / 4;
    size_t off = 0;
    w.table = off;
    if (want_table) off += align256((size_t)w.n_buckets * 32);
    w.tile_row0 = off;
    if (want_table) off += align256((size_t)(w.n_tiles + 1) * 8);
    w.maxdeg = off;
    if (want_table) off += 256;
    w.row32 = off;
    if (offsets_fit && n_nodes > 0) off += align256((size_t)(n_nodes + 1) * 4);
    w.total = off;
    return w;
}

// Optional persisting-L2 window over the row index (experiment: option persist_row_ptr).
static void set_row_window(cudaStream_t st, const void* base, size_t bytes, int device, bool on) {
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
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    } else {
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);  // give the set-aside lines back
    }
    cudaGetLastError();
}

// Validates the graph-side arguments, derives the acceptance thresholds, re-encodes row_ptr and
// (node2vec only) builds the membership table into `workspace`.  After this the plan can launch
// any number of shards of start nodes (trw_walk_csr launches one; trw_walk_csr_host one per chunk).
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
    a.out = nullptr; a.out_row_stride = 0; a.table = nullptr; a.row32 = nullptr;
    a.thr0 = a.thr1 = a.thr2 = 0;
    plan->device = device;
    plan->uniform = (p == 1.0 && q == 1.0);  // rw_cuda.cu:226
    plan->table = false;
    plan->stage = opt.stage_output != 0;
    plan->persist = opt.persist_row_ptr != 0;
    plan->min_ctas = (int)opt.n2v_min_ctas;
    plan->speculate = false;
    if (!plan->uniform) {
        const double mx = fmax(fmax(1.0 / p, 1.0), 1.0 / q);  // rw_cuda.cu:119-123
        const double p0 = 1.0 / p / mx, p1 = 1.0 / mx, p2 = 1.0 / q / mx;
        a.thr0 = threshold(p0); a.thr1 = threshold(p1); a.thr2 = threshold(p2);
        // Fetch row_ptr[x] before the verdict only when most proposals are accepted anyway.
        plan->speculate = opt.n2v_speculate < 0 ? (fmin(p1, p2) >= 0.5) : (opt.n2v_speculate != 0);
    }
    if (workspace == nullptr) return TRW_OK;  // reference-style path: int64 row_ptr, linear-scan membership
    const WsLayout w = ws_layout(n_nodes, nnz, plan->uniform);
    if (w.total == 0) return TRW_OK;
    if (workspace_bytes < w.total || ((uintptr_t)workspace & 255)) {
        set_error("trw_walk_csr: workspace needs %zu bytes at 256-byte alignment (got %zu)", w.total, workspace_bytes);
        return TRW_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    const bool want_table = !plan->uniform && opt.n2v_table != 0 && w.tile_row0 > w.table && a.thr1 != a.thr2;
    const bool want_row32 = opt.row32 != 0 && w.total > w.row32;
    if (!want_table && !want_row32) return TRW_OK;

    BuildArgs b{};
    b.row_ptr = row_ptr; b.col_idx = col_idx; b.n_nodes = n_nodes; b.nnz = nnz;
    b.n_tiles = w.n_tiles; b.n_buckets = w.n_buckets;
    b.row32 = want_row32 ? (uint32_t*)(ws + w.row32) : nullptr;
    if (want_table) {
        b.table = (uint32_t*)(ws + w.table);
        b.tile_row0 = (int64_t*)(ws + w.tile_row0);
        b.maxdeg = (unsigned long long*)(ws + w.maxdeg);
    }
    timing_begin(0, st);
    int rc;
    if (want_table) {
        rc = check_cuda(cudaMemsetAsync(b.maxdeg, 0, 8, st), "maxdeg memset");
        if (rc) return rc;
    }
    const int sms = sm_count(device);
    csr_prepass_kernel<<<sms * 8, 256, 0, st>>>(b);
    count_launch(1);
    rc = check_cuda(cudaGetLastError(), "csr_prepass launch");
    if (rc) return rc;
    if (want_table) {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (opt.build_mode == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, build_member_table_coop_kernel, kBuildThreads, 0);
        if (opt.build_mode == 1 && coop && per_sm > 0) {
            const int grid = sms * per_sm;
            b.chunk_tiles = grid * (int)(opt.build_tiles_per_cta > 0 ? opt.build_tiles_per_cta : 2);
            void* args[] = {&b};
            rc = check_cuda(cudaLaunchCooperativeKernel((void*)build_member_table_coop_kernel, dim3(grid), dim3(kBuildThreads),
                                                        args, 0, st), "build_member_table (cooperative) launch");
        } else {
            rc = check_cuda(cudaMemsetAsync(b.table, 0xFF, (size_t)w.n_buckets * 32, st), "table memset");
            if (rc) return rc;
            if (opt.build_mode == 0) {
                build_member_table_flat_kernel<<<(unsigned)w.n_tiles, kBuildThreads, 0, st>>>(b);
            } else {
                static bool attr_set[64];
                if (device < 64 && !attr_set[device]) {
                    rc = check_cuda(cudaFuncSetAttribute(build_member_table_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)kTiledSmemBytes), "tiled build smem attribute");
                    if (rc) return rc;
                    attr_set[device] = true;
                }
                build_member_table_tiled_kernel<<<(unsigned)w.n_tiles, kBuildThreads, kTiledSmemBytes, st>>>(b);
            }
            rc = check_cuda(cudaGetLastError(), "build_member_table launch");
        }
        count_launch(1);
        if (rc) return rc;
        a.table = b.table;
        plan->table = true;
    }
    timing_end(0, st);
    a.row32 = b.row32;
    return TRW_OK;
}

int csr_walk_launch(const CsrWalkPlan& plan, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                    int64_t* out, int64_t out_row_stride, cudaStream_t st) {
    if (n_walks <= 0) return TRW_OK;
    WalkArgs a = plan.a;
    a.targets = targets; a.n_walks = n_walks; a.walk_id_offset = walk_id_offset;
    a.out = out; a.out_row_stride = out_row_stride;
    const bool stage = plan.stage && (((uintptr_t)out & 7) == 0);
    const bool row32 = a.row32 != nullptr;
    if (plan.persist) {
        if (row32) set_row_window(st, a.row32, (size_t)(a.n_nodes + 1) * 4, plan.device, true);
        else set_row_window(st, a.row_ptr, (size_t)(a.n_nodes + 1) * 8, plan.device, true);
    }
    timing_begin(1, st);
    if (plan.uniform) {
        if (!stage) launch_uniform<false, false>(a, st);
        else if (row32) launch_uniform<true, true>(a, st);
        else launch_uniform<true, false>(a, st);
    } else {
        launch_n2v(a, stage, plan.table, row32 && stage, plan.speculate, plan.min_ctas, st);
    }
    timing_end(1, st);
    count_launch(1);
    const int rc = check_cuda(cudaGetLastError(), "walk kernel launch");
    if (plan.persist) set_row_window(st, nullptr, 0, plan.device, false);
    return rc;
}

}  // namespace trw

using namespace trw;

extern "C" size_t trw_walk_csr_workspace_bytes(int64_t n_nodes, int64_t nnz, double p, double q) {
    if (n_nodes < 0 || nnz < 0) return 0;
    return ws_layout(n_nodes, nnz, p == 1.0 && q == 1.0).total;
}

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
