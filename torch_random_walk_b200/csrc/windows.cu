// Window generation for sm_100a: one coalesced gather kernel per call.
//
// Replaces the four launchers/kernels of csrc/cuda/windows_cuda.cu (to_windows_gpu :67-119 /
// create_windows :7-65; to_windows_cbow_gpu :187-239 / :122-185; to_windows_triples_gpu
// :375-435 / :241-373; to_windows_triples_cbow_gpu :584-644 / :438-582).
//
// The reference gives each walk to one thread, which then writes its windows with 8-byte
// stores scattered over three tensors.  Here the problem is turned around: every output tensor
// is a dense array whose element e is a pure function of e (which walk, which window, which
// slot), so a CTA stages a tile of walk rows in shared memory once and then streams the three
// contiguous output segments that belong to that tile, coalesced: node windows as whole 32-byte
// rows per thread where the shape allows (else 16-byte pairs through window_element()), triple
// rows produced whole into per-warp stages (triple_tile()).  Positives and targets are bit-exact
// with the reference; negatives are Philox draws indexed by the output element (one block per
// aligned quad of draws), so they do not depend on the launch shape or on which path wrote them.
#include <algorithm>
#include <cmath>
#include <vector>

#include "trw_common.cuh"
#include "trw_options.h"

namespace trw {

enum WinMode { kSkipGram = 0, kCbow = 1, kTriples = 2, kTriplesCbow = 3 };

// Exact division of a 32-bit numerator by a runtime constant (magic = floor(2^64/d)+1).
struct FastDiv {
    uint32_t d;
    uint64_t magic;
    __host__ void set(uint32_t div) { d = div; magic = div > 1 ? (~0ull / div) + 1 : 0; }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d > 1 ? (uint32_t)__umul64hi(n, magic) : n; }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

struct WinArgs {
    const int64_t* walks;
    int64_t n_walks;
    int wl;        // columns of `walks`
    int W;         // window_size
    int mid;       // W / 2
    int per_walk;  // windows (skip-gram/CBOW) or target triples per walk
    int tile_walks;
    int use_smem;
    int bulk;          // triple modes: the warps' stages leave as bulk stores (cp.async.bulk) instead of 16-byte stores
    int direct_pos;    // triple modes: positive windows go from the tile to global memory in 16-byte pieces, no stage (window_size <= 10)
    int64_t num_nodes, pad;
    const int64_t* triples;
    int64_t n_triples;
    const uint4* triples16;   // optional 16-byte copy of `triples` (x, y, z = head, relation, tail as uint32), see compact_triples_kernel
    const int* triples16_bad; // device flag: some value did not fit, the copy must not be used
    const uint2* alias;       // optional alias table over [0, num_nodes) for the node windows' negatives ({threshold, alias} per node; see alias_draw)
    uint2 key;
    int64_t* out[3];       // outputs in the API's order
    uint32_t epw[3];       // elements of out[k] per walk
    FastDiv by_per_walk;   // / per_walk
    FastDiv by_row;        // / (W-1)   (node windows)  or  / (2W) (triple windows)
};

// Draw #draw of stream `stream` for item g: 64 random bits as two words.
__device__ __forceinline__ uint2 draw64(const uint2 key, uint64_t g, uint32_t stream, uint32_t draw) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), stream, draw >> 1), key);
    return (draw & 1u) ? make_uint2(r.z, r.w) : make_uint2(r.x, r.y);
}

// One draw from an alias table (Walker/Vose): a uniform cell i, then i itself with probability threshold_i / 2^32, else the
// cell's alias.  Two random words per draw, one 8-byte gather.  (negatives from a non-uniform
// distribution such as word2vec's degree^0.75; the reference draws uniformly, windows_cuda.cu:57-62.)
__device__ __forceinline__ int64_t alias_draw(const uint2* __restrict__ alias, uint32_t n, uint32_t r_cell, uint32_t r_keep) {
    const uint32_t i = __umulhi(r_cell, n);
    const uint2 c = __ldg(alias + i);
    return (int64_t)(r_keep < c.x ? i : c.y);
}

template <int MODE>
__device__ __forceinline__ int64_t window_element(const WinArgs& a, const int64_t* tile, int which, uint32_t e, uint64_t g) {
    // e: element index inside this tile's segment of out[which]; g: the same index in the whole tensor.
    if (MODE == kSkipGram || MODE == kCbow) {  // element path of the node windows (fast row paths: windows_kernel)
        const int pos_out = (MODE == kSkipGram) ? 1 : 2, neg_out = (MODE == kSkipGram) ? 2 : 1;
        if (which == 0) {  // target_nodes / pos_nodes
            uint32_t i, s;
            a.by_per_walk.divmod(e, i, s);
            return tile[(size_t)i * a.wl + s + a.mid];
        }
        if (which == pos_out) {  // pos_windows / windows: the window without its middle element
            uint32_t k, j, i, s;
            a.by_row.divmod(e, k, j);
            a.by_per_walk.divmod(k, i, s);
            return tile[(size_t)i * a.wl + s + j + (j >= (uint32_t)a.mid ? 1u : 0u)];
        }
        (void)neg_out;
        if (MODE == kSkipGram) {  // neg_windows: uniform node id (windows_cuda.cu:57-62)
            if (a.alias != nullptr) {  // ... or a draw from the caller's distribution: two draws per Philox block, shared by the aligned pair
                const uint4 r = philox4x32_10(make_uint4((uint32_t)(g >> 1), (uint32_t)(g >> 33), 5u, 0x414C4941u), a.key);
                return (g & 1) ? alias_draw(a.alias, (uint32_t)a.num_nodes, r.z, r.w) : alias_draw(a.alias, (uint32_t)a.num_nodes, r.x, r.y);
            }
            if ((uint64_t)a.num_nodes <= 0xFFFFFFFFull) {  // four draws per Philox block, shared by the aligned quad
                const uint4 r = philox4x32_10(make_uint4((uint32_t)(g >> 2), (uint32_t)(g >> 34), 1u, 0x51554144u), a.key);
                const uint32_t w = (g & 2) ? ((g & 1) ? r.w : r.z) : ((g & 1) ? r.y : r.x);
                return (int64_t)__umulhi(w, (uint32_t)a.num_nodes);
            }
            uint2 r = draw64(a.key, g, 1u, 0u);
            return bounded(r.x, r.y, a.num_nodes);
        }
        // CBOW neg_nodes: redraw while equal to the positive node, at most 101 times (:160-167)
        uint32_t i, s;
        a.by_per_walk.divmod(e, i, s);
        const int64_t pos = tile[(size_t)i * a.wl + s + a.mid];
        int64_t neg = 0;
        for (uint32_t attempt = 0; attempt <= 101u; ++attempt) {
            uint2 r = draw64(a.key, g, a.alias != nullptr ? 6u : 2u, attempt);
            neg = a.alias != nullptr ? alias_draw(a.alias, (uint32_t)a.num_nodes, r.x, r.y) : bounded(r.x, r.y, a.num_nodes);
            if (neg != pos) break;
        }
        return neg;
    }
    return 0;  // triple modes never come here: triple_tile() produces their rows whole
}

// ------------------------------------------------------------------------------------------
// Triple modes.  Every output row is a triple (24 bytes); producing the outputs element by element
// pays the index arithmetic and, for the negatives, the random draw three times per row, and kept
// the kernel issue-bound at 3-4 TB/s (59 % of the issue slots, ncu).  So rows are produced whole --
// one thread decodes (walk, target, window row) once and computes the three values -- into a
// per-warp shared-memory stage, and the stage leaves with coalesced 16-byte stores: per-lane 24-byte pieces
// written directly would touch every 128-byte line of the warp's span three times, and requests,
// not bytes, are what the memory system charges for.
// ------------------------------------------------------------------------------------------
constexpr size_t kMaxTileSmem = 176 * 1024;  // dynamic shared memory of a tile: 227 KiB per CTA less the 40 KiB of static stages, with room to spare
constexpr int kNegChunkRows = 2048;  // negative rows drawn per round (8 KiB of row indices)
constexpr int kWarpRows = 64;  // rows a warp stages per round: 1536 bytes, three 16-byte pieces per lane

// Every warp takes blocks of kWarpRows consecutive rows of the segment: row_fn(row, stage_slot) writes one row's three
// values, two rows per lane; then the warp streams its 1.5 KB to dst.  Only warp-level synchronisation.
//
// BULK: the warp's 1.5 KB leave as ONE bulk store of the copy engine (cp.async.bulk.global.shared::cta, TMA's linear
// form, SASS UBLKCP) issued by lane 0 from one of two stages per warp, instead of 16-byte stores from every lane.
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}

struct NoPre {
    __device__ __forceinline__ void operator()(uint32_t, uint32_t) const {}
};
// pre(first, cnt) is called by every lane before the rows of a round are produced (and followed by a __syncwarp):
// per-round work that the rows share, such as drawing the round's random row indices.
template <int BLOCK, bool BULK, class RowFn, class Pre = NoPre>
__device__ __forceinline__ void emit_rows(int64_t* __restrict__ dst, uint32_t n_rows, int64_t* __restrict__ stage, uint32_t& round,
                                          RowFn row_fn, Pre pre = Pre()) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = (((uintptr_t)dst) & 15) == 0;
    for (uint32_t first = warp * kWarpRows; first < n_rows; first += (BLOCK / 32) * kWarpRows) {
        const uint32_t cnt = min((uint32_t)kWarpRows, n_rows - first);
        pre(first, cnt);
        __syncwarp();
        int64_t* wstage = stage + (warp * (BULK ? 2 : 1) + (BULK ? (round & 1u) : 0u)) * (kWarpRows * 3);
        if (BULK) {  // this stage was handed to the copy engine two rounds ago
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        for (uint32_t g = lane; g < cnt; g += 32) row_fn(first + g, wstage + 3u * g);
        if (BULK && vec) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' writes -> visible to the copy engine
            __syncwarp();
            int64_t* out = dst + (uint64_t)first * 3u;
            const uint32_t even = cnt & ~1u;  // the engine moves multiples of 16 bytes: an odd last row goes by hand
            if (lane == 0) {
                if (even) bulk_store(out, wstage, even * 24u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (cnt != even && lane < 3) out[(uint64_t)even * 3u + lane] = wstage[(size_t)even * 3u + lane];
            ++round;
            continue;
        }
        __syncwarp();
        const uint32_t n_el = cnt * 3u;
        int64_t* out = dst + (uint64_t)first * 3u;  // first is even: 16-byte alignment carries over from dst
        if (vec) {
            for (uint32_t k = lane; k < (n_el >> 1); k += 32)
                reinterpret_cast<longlong2*>(out)[k] = reinterpret_cast<const longlong2*>(wstage)[k];
            if ((n_el & 1u) && lane == 0) out[n_el - 1] = wstage[n_el - 1];
        } else {
            for (uint32_t k = lane; k < n_el; k += 32) out[k] = wstage[k];
        }
        __syncwarp();
    }
}

template <int MODE, int BLOCK, bool BULK>
__device__ __forceinline__ void triple_tile(const WinArgs& a, const int64_t* tile, int64_t i0, int tw, int64_t* stage) {
    uint32_t round = 0;
    const uint32_t K = (uint32_t)a.per_walk, R = (uint32_t)(2 * a.W);
    const int pos_out = (MODE == kTriples) ? 1 : 2, neg_out = (MODE == kTriples) ? 2 : 1;
    // targets / pos_triples: row (i, ti) = walk[i][2*ti .. 2*ti+2]
    emit_rows<BLOCK, BULK>(a.out[0] + (uint64_t)i0 * K * 3u, (uint32_t)tw * K, stage, round, [&](uint32_t row, int64_t* o) {
        uint32_t i, ti;
        a.by_per_walk.divmod(row, i, ti);
        const int64_t* w = tile + (size_t)i * a.wl + 2u * ti;
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    });
    // Positive windows.  The 2W rows of one (walk, target) are 6W consecutive elements = 3W 16-byte pieces, and which walk
    // position (relative to the target) a piece's two elements come from depends on the piece alone.  So for 3W <= 32 a
    // lane keeps ONE piece slot for the whole tile -- its two source offsets are worked out once -- and a warp round
    // writes floor(32 / 3W) whole blocks: per piece two shared-memory loads and one 16-byte store, against the stage's
    // three loads, three 8-byte stage stores, one 16-byte stage load and the store (the kernel is bound by L1/shared-memory
    // wavefronts, ncu).  An earlier element-per-lane form (8-byte stores, offsets recomputed per element) lost to the stage.
    int64_t* const pos_dst = a.out[pos_out] + (uint64_t)i0 * K * R * 3u;
    const uint32_t P = 3u * (uint32_t)a.W;  // pieces per block
    const bool direct = R > 0 && a.direct_pos != 0 && P <= 32u && (((uintptr_t)pos_dst) & 15) == 0;
    if (direct) {
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t per_round = 32u / P, blk = lane / P, jp = lane - blk * P;
        // slot j of a block = row h = j / 3, column c = j % 3.  Left rows (h < W): (walk[ri], walk[ri], walk[ri+1]) with
        // ri = r - 2(h+1), the first two only where ri >= 1, the third where ri >= -1 (windows_cuda.cu:284-313); right rows:
        // walk[idx + c] with idx = r + 2(h-W+1) - 1 where it exists (:315-345).  As (offset from r, lowest valid position):
        int off[2], lo[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int j = 2 * (int)jp + k, h = j / 3, c = j - 3 * h;
            if (h < a.W) { off[k] = -2 * (h + 1) + (c == 2 ? 1 : 0); lo[k] = c == 2 ? 0 : 1; }
            else { off[k] = 2 * (h - a.W + 1) - 1 + c; lo[k] = 0; }
        }
        const uint32_t n_blocks = (uint32_t)tw * K;
        if (blk < per_round) {
            for (uint32_t b = warp * per_round + blk; b < n_blocks; b += (BLOCK / 32) * per_round) {
                uint32_t i, ti;
                a.by_per_walk.divmod(b, i, ti);
                const int64_t* w = tile + (size_t)i * a.wl;
                const int s0 = 2 * (int)ti + 1 + off[0], s1 = 2 * (int)ti + 1 + off[1];
                longlong2 v;
                v.x = (s0 >= lo[0] && s0 < a.wl) ? w[s0] : a.pad;
                v.y = (s1 >= lo[1] && s1 < a.wl) ? w[s1] : a.pad;
                reinterpret_cast<longlong2*>(pos_dst)[(uint64_t)b * P + jp] = v;
            }
        }
    }
    // (3W > 32, or an odd destination:) row (i, ti, h), the three slots at once, through the per-warp stage
    if (R > 0 && !direct) {
        emit_rows<BLOCK, BULK>(a.out[pos_out] + (uint64_t)i0 * K * R * 3u, (uint32_t)tw * K * R, stage, round, [&](uint32_t row, int64_t* o) {
            uint32_t k, h, i, ti;
            a.by_row.divmod(row, k, h);
            a.by_per_walk.divmod(k, i, ti);
            const int64_t* w = tile + (size_t)i * a.wl;
            const int r = 2 * (int)ti + 1;
            if ((int)h < a.W) {  // left rows: (walk[ri], walk[ri], walk[ri+1]) -- windows_cuda.cu:284-313, head slot as the reference has it
                const int ri = r - 2 * ((int)h + 1);
                const int64_t rel = ri >= 1 ? w[ri] : a.pad;
                o[0] = rel; o[1] = rel;
                o[2] = ri >= -1 ? w[ri + 1] : a.pad;
            } else {             // right rows: walk[idx .. idx+2] where it exists (:315-345)
                const int idx = r + 2 * ((int)h - a.W + 1) - 1;
                o[0] = idx < a.wl ? w[idx] : a.pad;
                o[1] = idx + 1 < a.wl ? w[idx + 1] : a.pad;
                o[2] = idx + 2 < a.wl ? w[idx + 2] : a.pad;
            }
        });
    }
    // The negatives gather rows of `triples`.  As int64 a row is three 8-byte loads that straddle sectors; the caller may
    // pass a 16-byte uint32 copy (one LDG.128 per row, a third of the instructions), valid when every id fits 32 bits.
    const bool small_table = (uint64_t)a.n_triples <= 0xFFFFFFFFull;
    const uint4* t16 = (a.triples16 != nullptr && small_table && *a.triples16_bad == 0) ? a.triples16 : nullptr;
    if (MODE == kTriples && t16 != nullptr) {
        // neg_windows from the 16-byte copy.  ncu: the kernel is bound by L1/shared-memory wavefronts (92 %), not by HBM or
        // issue slots, so rows do not go through a stage here: a warp draws the row indices of 64 rows (four per Philox
        // block; same blocks and word order as the int64 path below) and then stores element by element, coalesced --
        // the three lanes of a row read one sector of the table between them.
        __shared__ uint32_t drawn[BLOCK / 32][kWarpRows];
        const uint64_t row0 = (uint64_t)i0 * K * R;
        const uint32_t n_rows = (uint32_t)tw * K * R;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t* mine = drawn[warp];
        int64_t* dst = a.out[neg_out] + row0 * 3u;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(t16);
        const uint32_t nt = (uint32_t)a.n_triples;
        for (uint32_t first = warp * kWarpRows; first < n_rows; first += (BLOCK / 32) * kWarpRows) {
            const uint32_t cnt = min((uint32_t)kWarpRows, n_rows - first);
            if (lane * 4u < cnt) {
                const uint64_t gq = (row0 + first) / 4u + lane;
                const uint4 r = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), 3u, 0x51554144u), a.key);
                mine[lane * 4 + 0] = __umulhi(r.x, nt); mine[lane * 4 + 1] = __umulhi(r.y, nt);
                mine[lane * 4 + 2] = __umulhi(r.z, nt); mine[lane * 4 + 3] = __umulhi(r.w, nt);
            }
            __syncwarp();
            int64_t* out = dst + (uint64_t)first * 3u;
            const uint32_t n = cnt * 3u;
#pragma unroll 2
            for (uint32_t e = lane; e < n; e += 32) {
                const uint32_t row = (e * 0xAAABu) >> 17;  // e / 3 for e < 2^15
                const uint32_t c = e - row * 3u;
                out[e] = (int64_t)__ldg(words + 4u * mine[row] + c);
            }
            __syncwarp();
        }
    } else if (MODE == kTriples) {
        // neg_windows: every row a uniformly drawn row of `triples` (windows_cuda.cu:353-365).  The row
        // indices of a chunk are drawn first, four per Philox block (the tile's first row is a multiple
        // of four), then the rows are gathered element by element straight into the coalesced stores
        // (measured: gathering whole rows per lane through the stage is slower, 3.3 vs 4.3 TB/s -- the
        // gathers hit the L2-resident table and want one sector request per ~3 lanes, not three per lane).
        __shared__ uint32_t neg_rows[kNegChunkRows];
        const int tid = threadIdx.x;
        const uint64_t row0 = (uint64_t)i0 * K * R;
        const uint32_t n_rows = (uint32_t)tw * K * R;
        int64_t* dst = a.out[neg_out] + row0 * 3u;
        const bool small = small_table;
        for (uint32_t base = 0; base < n_rows; base += kNegChunkRows) {
            const uint32_t rows = min((uint32_t)kNegChunkRows, n_rows - base);
            for (uint32_t q = tid; q * 4u < rows; q += BLOCK) {
                const uint64_t gq = (row0 + base) / 4u + q;
                const uint4 r = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), 3u, 0x51554144u), a.key);
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q * 4u + j < rows) neg_rows[q * 4u + j] = small ? __umulhi(w[j], (uint32_t)a.n_triples) : w[j];
            }
            __syncthreads();
            const uint32_t n = rows * 3u;
            for (uint32_t e = tid; e < n; e += BLOCK) {
                const uint32_t row = e / 3u, c = e - row * 3u;
                int64_t idx = neg_rows[row];
                if (!small) {  // more than 2^32 triples: widen the draw with a second block (never in practice)
                    const uint2 r = draw64(a.key, row0 + base + row, 3u, 1u);
                    idx = bounded((uint32_t)idx, r.x, a.n_triples);
                }
                dst[(uint64_t)base * 3u + e] = __ldg(a.triples + idx * 3 + c);
            }
            __syncthreads();
        }
    } else {
        // neg_triples: one row per target, redrawn while identical to the positive triple (windows_cuda.cu:485-505)
        emit_rows<BLOCK, BULK>(a.out[neg_out] + (uint64_t)i0 * K * 3u, (uint32_t)tw * K, stage, round, [&](uint32_t row, int64_t* o) {
            uint32_t i, ti;
            a.by_per_walk.divmod(row, i, ti);
            const uint64_t grow = ((uint64_t)i0 + i) * K + ti;
            const int64_t* w = tile + (size_t)i * a.wl + 2u * ti;
            const int64_t ph = w[0], pr = w[1], pt = w[2];
            int64_t nh = 0, nr = 0, nt = 0;
            for (uint32_t attempt = 0; attempt <= 101u; ++attempt) {
                const uint2 r = draw64(a.key, grow, 4u, attempt);
                const int64_t pick = bounded(r.x, r.y, a.n_triples);
                if (t16 != nullptr) {
                    const uint4 t = __ldg(t16 + pick);
                    nh = (int64_t)t.x; nr = (int64_t)t.y; nt = (int64_t)t.z;
                } else {
                    const int64_t* t = a.triples + pick * 3;
                    nh = __ldg(t); nr = __ldg(t + 1); nt = __ldg(t + 2);
                }
                if (nh != ph || nr != pr || nt != pt) break;
            }
            o[0] = nh; o[1] = nr; o[2] = nt;
        });
    }
    if (BULK && (threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stages must outlive their copies
}

template <int MODE, int BLOCK, bool BULK = false>
__global__ void __launch_bounds__(BLOCK) windows_kernel(const WinArgs a) {
    extern __shared__ __align__(16) int64_t smem_tile[];
    const int64_t i0 = (int64_t)blockIdx.x * a.tile_walks;
    const int tw = (int)min((int64_t)a.tile_walks, a.n_walks - i0);
    const int64_t* gsrc = a.walks + i0 * a.wl;
    const int64_t* tile = gsrc;
    if (a.use_smem) {
        const int n = tw * a.wl;
        if ((((uintptr_t)gsrc) & 15) == 0) {
            const int n2 = n >> 1;
            const longlong2* src2 = reinterpret_cast<const longlong2*>(gsrc);
            longlong2* dst2 = reinterpret_cast<longlong2*>(smem_tile);
            for (int k = threadIdx.x; k < n2; k += BLOCK) dst2[k] = __ldg(src2 + k);
            if ((n & 1) && threadIdx.x == 0) smem_tile[n - 1] = __ldg(gsrc + n - 1);
        } else {
            for (int k = threadIdx.x; k < n; k += BLOCK) smem_tile[k] = __ldg(gsrc + k);
        }
        __syncthreads();
        tile = smem_tile;
    }
    if constexpr (MODE == kTriples || MODE == kTriplesCbow) {
        __shared__ __align__(128) int64_t stage[(BLOCK / 32) * kWarpRows * 3 * (BULK ? 2 : 1)];  // per warp: one stage, two for bulk stores
        triple_tile<MODE, BLOCK, BULK>(a, tile, i0, tw, stage);
    } else {
#pragma unroll
    for (int which = 0; which < 3; ++which) {
        const uint64_t gbase = (uint64_t)i0 * a.epw[which];
        int64_t* dst = a.out[which] + gbase;
        const uint32_t n = (uint32_t)tw * a.epw[which];
        // Fast paths of the node windows: whole 32-byte rows per thread.  The generic loop below pays two
        // divisions per 8-byte element and, for the negatives, one Philox block per element although a
        // block yields the four draws of an aligned quad -- that redundancy alone kept the kernel
        // issue-bound at 4.0 TB/s.
        if (MODE == kSkipGram && which == 2 && a.alias != nullptr && (gbase & 1u) == 0 && (((uintptr_t)dst) & 15) == 0) {
            const uint32_t np = n >> 1, nn = (uint32_t)a.num_nodes;  // a pair of negatives per thread: one Philox block, two gathers, one 16-byte store
            for (uint32_t q = threadIdx.x; q < np; q += BLOCK) {
                const uint64_t gp = (gbase >> 1) + q;  // same block and word order as window_element()
                const uint4 r = philox4x32_10(make_uint4((uint32_t)gp, (uint32_t)(gp >> 32), 5u, 0x414C4941u), a.key);
                longlong2 v;
                v.x = alias_draw(a.alias, nn, r.x, r.y);
                v.y = alias_draw(a.alias, nn, r.z, r.w);
                reinterpret_cast<longlong2*>(dst)[q] = v;
            }
            if ((n & 1u) && threadIdx.x == 0) dst[n - 1] = window_element<MODE>(a, tile, which, n - 1, gbase + n - 1);
            continue;
        }
        if (MODE == kSkipGram && which == 2 && a.alias == nullptr && (uint64_t)a.num_nodes <= 0xFFFFFFFFull && (gbase & 3u) == 0 &&
            (((uintptr_t)dst) & 31) == 0) {
            const uint32_t nq = n >> 2, nn = (uint32_t)a.num_nodes;
            for (uint32_t q = threadIdx.x; q < nq; q += BLOCK) {
                const uint64_t gq = (gbase >> 2) + q;  // same block and word order as window_element()
                const uint4 r = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), 1u, 0x51554144u), a.key);
                stg_sector(dst + 4u * q, __umulhi(r.x, nn), __umulhi(r.y, nn), __umulhi(r.z, nn), __umulhi(r.w, nn));
            }
            for (uint32_t k = (nq << 2) + threadIdx.x; k < n; k += BLOCK) dst[k] = window_element<MODE>(a, tile, which, k, gbase + k);
            continue;
        }
        if ((MODE == kSkipGram ? which == 1 : (MODE == kCbow && which == 2)) && a.W == 5 && (((uintptr_t)dst) & 31) == 0) {
            const uint32_t nw = (uint32_t)tw * (uint32_t)a.per_walk;  // one window = one 32-byte row (w0, w1, w3, w4)
            for (uint32_t k = threadIdx.x; k < nw; k += BLOCK) {
                uint32_t i, sidx;
                a.by_per_walk.divmod(k, i, sidx);
                const int64_t* w = tile + (size_t)i * a.wl + sidx;
                stg_sector(dst + 4u * k, (uint64_t)w[0], (uint64_t)w[1], (uint64_t)w[3], (uint64_t)w[4]);
            }
            continue;
        }
        if ((((uintptr_t)dst) & 15) == 0) {
            const uint32_t n2 = n >> 1;
            for (uint32_t k = threadIdx.x; k < n2; k += BLOCK) {
                longlong2 v;
                v.x = window_element<MODE>(a, tile, which, 2 * k, gbase + 2 * k);
                v.y = window_element<MODE>(a, tile, which, 2 * k + 1, gbase + 2 * k + 1);
                reinterpret_cast<longlong2*>(dst)[k] = v;
            }
            if ((n & 1u) && threadIdx.x == 0) dst[n - 1] = window_element<MODE>(a, tile, which, n - 1, gbase + n - 1);
        } else {
            for (uint32_t k = threadIdx.x; k < n; k += BLOCK) dst[k] = window_element<MODE>(a, tile, which, k, gbase + k);
        }
    }
    }
}

// 16-byte copy of the triple table for the negative gathers: (head, relation, tail, 0) as uint32.  *bad is raised when a
// value does not fit (negative or >= 2^32); the window kernel then reads the int64 table as before.
__global__ void __launch_bounds__(256) compact_triples_kernel(const int64_t* __restrict__ triples, int64_t n, uint4* __restrict__ out,
                                                              int* __restrict__ bad) {
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    bool wide = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gsz) {
        const int64_t h = __ldg(triples + 3 * i), r = __ldg(triples + 3 * i + 1), t = __ldg(triples + 3 * i + 2);
        wide |= ((uint64_t)h | (uint64_t)r | (uint64_t)t) > 0xFFFFFFFFull;
        out[i] = make_uint4((uint32_t)h, (uint32_t)r, (uint32_t)t, 0u);
    }
    if (wide) *bad = 1;
}

static size_t triples_workspace_bytes(int64_t n_triples) { return n_triples > 0 ? (((size_t)n_triples * 16 + 255) & ~(size_t)255) + 256 : 0; }

template <int MODE>
static int launch_windows(const char* name, const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                          int64_t num_nodes, int64_t pad, const int64_t* triples, int64_t n_triples, int64_t seed,
                          int64_t* o0, int64_t* o1, int64_t* o2, int device, void* stream, void* workspace = nullptr,
                          size_t workspace_bytes = 0, const uint64_t* alias_table = nullptr) {
    constexpr bool kTripleMode = (MODE == kTriples || MODE == kTriplesCbow);
    if (n_walks < 0 || walk_cols < 0 || window_size < 0) { set_error("%s: negative size", name); return TRW_ERR_ARG; }
    if (walk_cols > (1 << 24) || window_size > (1 << 15)) {
        set_error("%s: walk rows longer than 2^24 or windows wider than 2^15 are not supported", name);
        return TRW_ERR_ARG;
    }
    int64_t per_walk;
    uint32_t epw[3];
    const int W = window_size;
    if (!kTripleMode) {
        if (W < 1) { set_error("%s: window_size must be >= 1", name); return TRW_ERR_ARG; }
        per_walk = walk_cols - W + 1;
        if (per_walk < 0) per_walk = 0;  // the reference would ask torch for a negative size here
        epw[0] = (uint32_t)per_walk;
        epw[1] = epw[2] = (uint32_t)(per_walk * (W - 1));
        if (MODE == kCbow) epw[1] = (uint32_t)per_walk;
    } else {
        per_walk = walk_cols >= 1 ? (walk_cols - 1) / 2 : 0;
        epw[0] = (uint32_t)(per_walk * 3);
        epw[1] = epw[2] = (uint32_t)(per_walk * 2 * W * 3);
        if (MODE == kTriplesCbow) epw[1] = (uint32_t)(per_walk * 3);
    }
    const int64_t max_epw = (int64_t)per_walk * (kTripleMode ? (int64_t)6 * (W > 0 ? W : 1) : (W > 1 ? W - 1 : 1));
    if (max_epw >= (1ll << 29)) { set_error("%s: more than 2^29 output elements per walk", name); return TRW_ERR_ARG; }
    if (n_walks == 0 || per_walk == 0) return resolve_device(device) < 0 ? TRW_ERR_DEVICE : TRW_OK;
    if (!walks || !o0 || (epw[1] && !o1) || (epw[2] && !o2)) { set_error("%s: null pointer", name); return TRW_ERR_ARG; }
    if (!kTripleMode && num_nodes <= 0) { set_error("%s: num_nodes must be positive", name); return TRW_ERR_ARG; }
    if (kTripleMode && (n_triples <= 0 || !triples)) { set_error("%s: triples must not be empty", name); return TRW_ERR_ARG; }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("%s: cudaSetDevice(%d) failed", name, d); return TRW_ERR_DEVICE; }

    WinArgs a;
    a.walks = walks; a.n_walks = n_walks; a.wl = (int)walk_cols; a.W = W; a.mid = W / 2; a.per_walk = (int)per_walk;
    a.num_nodes = num_nodes; a.pad = pad; a.triples = triples; a.n_triples = n_triples;
    a.triples16 = nullptr; a.triples16_bad = nullptr;
    a.alias = reinterpret_cast<const uint2*>(alias_table);
    if (alias_table != nullptr && (kTripleMode || (uint64_t)num_nodes > 0xFFFFFFFFull || ((uintptr_t)alias_table & 7))) {
        set_error("%s: an alias table needs node windows, num_nodes < 2^32 and 8-byte alignment", name);
        return TRW_ERR_ARG;
    }
    if (kTripleMode && workspace != nullptr && options().win_table16 != 0) {
        if (workspace_bytes < triples_workspace_bytes(n_triples) || ((uintptr_t)workspace & 255)) {
            set_error("%s: workspace needs %zu bytes at 256-byte alignment", name, triples_workspace_bytes(n_triples));
            return TRW_ERR_WORKSPACE;
        }
        uint4* t16 = reinterpret_cast<uint4*>(workspace);
        int* bad = reinterpret_cast<int*>((char*)workspace + triples_workspace_bytes(n_triples) - 256);
        int rc = check_cuda(cudaMemsetAsync(bad, 0, sizeof(int), (cudaStream_t)stream), "triples flag memset");
        if (rc) return rc;
        const int64_t blocks = std::min<int64_t>((n_triples + 255) / 256, (int64_t)sm_count(d) * 8);
        compact_triples_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(triples, n_triples, t16, bad);
        count_launch(1);
        a.triples16 = t16; a.triples16_bad = bad;
    }
    a.key = philox_key(seed, kTagWindows);
    a.out[0] = o0; a.out[1] = o1; a.out[2] = o2;
    for (int k = 0; k < 3; ++k) a.epw[k] = epw[k];
    a.by_per_walk.set((uint32_t)per_walk);
    a.by_row.set(kTripleMode ? (uint32_t)(2 * W) : (uint32_t)(W - 1));
    // Tile: a multiple of four walk rows, about 32 KiB of them, and < 2^31 elements per output segment.
    constexpr int BLOCK = 256;
    // (triple modes add a 12 KiB output stage and 8 KiB of drawn row indices: a 16 KiB tile keeps six CTAs per SM)
    int64_t tw = ((kTripleMode ? 16 : 32) * 1024) / (walk_cols * 8);
    const int64_t cap = (1ll << 30) / (max_epw > 0 ? max_epw : 1);
    if (tw > cap) tw = cap;
    tw &= ~3ll;  // a multiple of four: every tile's output segments start on a 32-byte boundary of their tensor
    if (tw < 4) tw = 4;
    size_t smem = (size_t)tw * walk_cols * 8;
    a.use_smem = 1;
    if (smem > kMaxTileSmem) { a.use_smem = 0; smem = 0; }
    a.tile_walks = (int)tw;
    // bulk stores need 16-byte aligned segments: true for whole tensors from any allocator worth the name and for
    // tiles that start on a multiple of four walks
    a.direct_pos = options().win_direct_pos != 0 ? 1 : 0;
    a.bulk = 0;
    if (kTripleMode && options().win_bulk != 0 && (((uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)o2) & 15) == 0) a.bulk = 1;
    const bool bulk = kTripleMode && a.bulk != 0;
    if (kTripleMode || smem > 24 * 1024) {  // static shared memory (stages, drawn row indices) counts against the 48 KiB default too
        // raise the opt-in limit once per device and kernel; the call is far too slow to repeat per launch
        static size_t opted_in[64][2];
        if (d >= 64 || opted_in[d][bulk] < smem) {
            int rc = bulk ? check_cuda(cudaFuncSetAttribute(windows_kernel<MODE, BLOCK, kTripleMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                            (int)kMaxTileSmem), name)
                          : check_cuda(cudaFuncSetAttribute(windows_kernel<MODE, BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                            (int)kMaxTileSmem), name);
            if (rc) return rc;
            if (d < 64) opted_in[d][bulk] = kMaxTileSmem;
        }
    }
    const int64_t tiles = (n_walks + tw - 1) / tw;
    if (tiles > 0x7FFFFFFFll) { set_error("%s: too many tiles", name); return TRW_ERR_ARG; }
    if (bulk) windows_kernel<MODE, BLOCK, kTripleMode><<<(unsigned)tiles, BLOCK, smem, (cudaStream_t)stream>>>(a);
    else windows_kernel<MODE, BLOCK, false><<<(unsigned)tiles, BLOCK, smem, (cudaStream_t)stream>>>(a);
    count_launch(1);
    return check_cuda(cudaGetLastError(), name);
}

}  // namespace trw

using namespace trw;

extern "C" int trw_windows(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size, int64_t num_nodes,
                           int64_t seed, int64_t* target, int64_t* pos, int64_t* neg, int device, void* stream) {
    return launch_windows<kSkipGram>("trw_windows", walks, n_walks, walk_cols, window_size, num_nodes, 0, nullptr, 0, seed,
                                     target, pos, neg, device, stream);
}

extern "C" int trw_windows_cbow(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                                int64_t num_nodes, int64_t seed, int64_t* pos_nodes, int64_t* neg_nodes, int64_t* windows,
                                int device, void* stream) {
    return launch_windows<kCbow>("trw_windows_cbow", walks, n_walks, walk_cols, window_size, num_nodes, 0, nullptr, 0, seed,
                                 pos_nodes, neg_nodes, windows, device, stream);
}

extern "C" int trw_windows_triples(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                                   int64_t num_nodes, int64_t padding_idx, const int64_t* triples, int64_t n_triples,
                                   int64_t seed, int64_t* target, int64_t* pos, int64_t* neg, int device, void* stream) {
    return launch_windows<kTriples>("trw_windows_triples", walks, n_walks, walk_cols, window_size, num_nodes, padding_idx,
                                    triples, n_triples, seed, target, pos, neg, device, stream);
}

extern "C" size_t trw_windows_triples_workspace_bytes(int64_t n_triples) { return triples_workspace_bytes(n_triples); }

extern "C" int trw_windows_triples_ws(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size, int64_t num_nodes,
                                      int64_t padding_idx, const int64_t* triples, int64_t n_triples, int64_t seed, int64_t* target,
                                      int64_t* pos, int64_t* neg, void* workspace, size_t workspace_bytes, int device, void* stream) {
    return launch_windows<kTriples>("trw_windows_triples", walks, n_walks, walk_cols, window_size, num_nodes, padding_idx, triples,
                                    n_triples, seed, target, pos, neg, device, stream, workspace, workspace_bytes);
}

extern "C" int trw_windows_triples_cbow_ws(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                                           int64_t num_nodes, int64_t padding_idx, const int64_t* triples, int64_t n_triples,
                                           int64_t seed, int64_t* pos_triples, int64_t* neg_triples, int64_t* pos_windows,
                                           void* workspace, size_t workspace_bytes, int device, void* stream) {
    return launch_windows<kTriplesCbow>("trw_windows_triples_cbow", walks, n_walks, walk_cols, window_size, num_nodes, padding_idx,
                                        triples, n_triples, seed, pos_triples, neg_triples, pos_windows, device, stream, workspace,
                                        workspace_bytes);
}

extern "C" int trw_windows_triples_cbow(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                                        int64_t num_nodes, int64_t padding_idx, const int64_t* triples, int64_t n_triples,
                                        int64_t seed, int64_t* pos_triples, int64_t* neg_triples, int64_t* pos_windows,
                                        int device, void* stream) {
    return launch_windows<kTriplesCbow>("trw_windows_triples_cbow", walks, n_walks, walk_cols, window_size, num_nodes,
                                        padding_idx, triples, n_triples, seed, pos_triples, neg_triples, pos_windows, device,
                                        stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Negatives from a caller-supplied distribution (optional): an alias table over the node ids.
// trw_alias_table_build runs on the HOST (Vose's O(n) construction): cell i = {threshold_i, alias_i} packed as
// threshold | alias << 32, P(v) proportional to weights[v]^power (weights >= 0, not all zero).
extern "C" int trw_alias_table_build(const double* weights, int64_t n, double power, uint64_t* table_out) {
    if (!weights || !table_out || n <= 0 || (uint64_t)n > 0xFFFFFFFFull) { set_error("trw_alias_table_build: bad argument"); return TRW_ERR_ARG; }
    std::vector<double> p((size_t)n);
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double w = weights[i];
        if (!(w >= 0.0)) { set_error("trw_alias_table_build: weights must be >= 0"); return TRW_ERR_ARG; }
        p[(size_t)i] = w > 0.0 ? pow(w, power) : 0.0;
        total += p[(size_t)i];
    }
    if (!(total > 0.0)) { set_error("trw_alias_table_build: all weights are zero"); return TRW_ERR_ARG; }
    const double scale = (double)n / total;  // cell mass 1 = the average
    std::vector<uint32_t> small, large;
    small.reserve((size_t)n);
    large.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        p[(size_t)i] *= scale;
        (p[(size_t)i] < 1.0 ? small : large).push_back((uint32_t)i);
    }
    auto cell = [](double keep, uint32_t alias) {
        const double t = keep * 4294967296.0;
        const uint64_t thr = t >= 4294967295.0 ? 0xFFFFFFFFull : (t <= 0.0 ? 0ull : (uint64_t)t);
        return thr | ((uint64_t)alias << 32);
    };
    while (!small.empty() && !large.empty()) {
        const uint32_t s_ = small.back(), l_ = large.back();
        small.pop_back();
        table_out[s_] = cell(p[s_], l_);
        p[l_] -= 1.0 - p[s_];
        if (p[l_] < 1.0) { large.pop_back(); small.push_back(l_); }
    }
    for (uint32_t i : large) table_out[i] = cell(1.0, i);
    for (uint32_t i : small) table_out[i] = cell(1.0, i);  // rounding leftovers: mass 1 up to floating-point error
    return TRW_OK;
}

extern "C" int trw_windows_alias(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size, int64_t num_nodes,
                                 int64_t seed, const uint64_t* alias_table, int64_t* target, int64_t* pos, int64_t* neg, int device,
                                 void* stream) {
    return launch_windows<kSkipGram>("trw_windows_alias", walks, n_walks, walk_cols, window_size, num_nodes, 0, nullptr, 0, seed,
                                     target, pos, neg, device, stream, nullptr, 0, alias_table);
}

extern "C" int trw_windows_cbow_alias(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size, int64_t num_nodes,
                                      int64_t seed, const uint64_t* alias_table, int64_t* pos_nodes, int64_t* neg_nodes, int64_t* windows,
                                      int device, void* stream) {
    return launch_windows<kCbow>("trw_windows_cbow_alias", walks, n_walks, walk_cols, window_size, num_nodes, 0, nullptr, 0, seed,
                                 pos_nodes, neg_nodes, windows, device, stream, nullptr, 0, alias_table);
}
