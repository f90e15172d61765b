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
// Design (DESIGN.md sections 3-4).  Both kernels are dependent random gathers.  On B200 a gather that
// misses L2 moves a full 128-byte line of HBM, and what saturates first is the rate of requests
// leaving L2 for memory (~48 G/s, loads and stores alike); L2 hits and shared-memory footprint
// (which shrinks the L1 the in-flight gathers land in) cost too.  Hence:
//   * one thread owns one walk; about 1024 resident threads per SM saturate the request rate;
//   * the proposal gather reads a 16-byte edge record that carries the next node's row span, so a
//     step never touches the row index (member_table.cuh; without records the uint32 row index is
//     read with an L2 evict_last policy while everything streamed carries evict_first);
//   * the node2vec loop is flattened to one *trial* per iteration and picks, per (p, q), the
//     tightest exact sampling scheme: two-sided mixture, return-edge folding or plain rejection;
//   * "x in adj(t)" is one 32-byte sector of a hashed copy of the adjacency (member_table.cuh),
//     not a scan or search of adj(t), and is skipped whenever the uniform draw alone decides;
//   * output leaves in whole 128-byte lines written cooperatively by the warp (LineStager), one
//     request per line; the loops are therefore warp-converged, one trial per lane per iteration.
// Everything derived from the graph alone (csr_graph_prepare) is separate from the per-call plan
// (csr_walk_plan), so it is built per call by trw_walk_csr and once by trw_csr_graph_prepare.
#include <mutex>
#include <new>
#include <unordered_map>

#include "member_table.cuh"
#include "trw_common.cuh"
#include "trw_options.h"
#include "walk_csr.h"

namespace trw {

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

// Proposal: the neighbour of v at a uniformly random position, or v itself when it has none
// (rw_cuda.cu:8-31).  REC: the gather is an edge record and also returns the neighbour's own row
// span in (xb, xe); otherwise (xb, xe) are left for the caller to load.
// `bloom` receives the triangle Bloom of the edge (v -> result) when the record carries one, else kBloomAll.
template <bool REC>
__device__ __forceinline__ int64_t propose(const WalkArgs& a, int64_t v, int64_t b, int64_t e, uint32_t r0, uint32_t r1,
                                           uint64_t pol_stream, int64_t& xb, int64_t& xe, uint32_t& bloom) {
    const int64_t deg = e - b;
    const int64_t idx = deg > 0 ? b + bounded(r0, r1, deg) : -1;
    bloom = kBloomAll;
    if ((uint64_t)idx >= (uint64_t)a.nnz) {  // no out-edge (or a span outside col_idx): the walk stays on v
        xb = b; xe = e;
        return v;
    }
    if (REC) {
        const uint4 rec = ldg_u32x4_hint(a.records + idx, pol_stream);
        xb = (int64_t)rec.z;
        xe = xb + (int64_t)rec.y;
        if (rec.y != 0) { bloom = rec.w; return (int64_t)rec.x; }  // member_table.cuh: the last word is a Bloom ...
        return (int64_t)(((uint64_t)rec.w << 32) | rec.x);         // ... or the high word of an id without out-edges
    }
    return ldg64_hint(a.col_idx + idx, pol_stream);
}
template <bool REC>
__device__ __forceinline__ int64_t propose(const WalkArgs& a, int64_t v, int64_t b, int64_t e, uint32_t r0, uint32_t r1,
                                           uint64_t pol_stream, int64_t& xb, int64_t& xe) {
    uint32_t unused;
    return propose<REC>(a, v, b, e, r0, r1, pol_stream, xb, xe, unused);
}

// The two free looks at the triangle (t, v, x) before any memory is touched (member_table.cuh): the Bloom of
// (t -> v) carried along since v was proposed, and -- on a symmetric graph -- the Bloom that came with the
// proposal of x.  false: x is certainly not in adj(t).
__device__ __forceinline__ bool triangle_maybe(uint32_t carried, int64_t x, bool symmetric, uint32_t fresh, uint32_t other_bit) {
    return (carried & bloom_bit(x)) != 0 && (!symmetric || (fresh & other_bit) != 0);
}

// "x in adj(row)?" once the triangle Blooms have said "maybe": for a pair of hubs the L2-resident hub-pair filter first (a
// clear bit is a proof), the table sector only for another "maybe".  Identical answers with and without the filter.
// d_x: out-degree of x (0 when not known -- only an edge record brings it along: the filter is then not asked).  A table that failed to build leaves the filter
// unfilled, so without a table there is no filter either.
template <bool TABLE>
__device__ __forceinline__ bool is_member_filtered(const WalkArgs& a, int64_t x, int64_t d_x, int64_t row, int64_t b, int64_t e,
                                                   const uint32_t* table, uint64_t pol_stream, uint64_t pol_keep) {
    if (TABLE && table != nullptr && e > b && !filter_maybe(a.filter, row, x, e - b, d_x, pol_keep)) return false;
    return is_member<TABLE>(x, b, e, a.col_idx, table, pol_stream);
}

// Output of one walk row: line-staged cooperative stores (STAGE) or plain 8-byte stores.  put() is a
// warp collective (see LineStager): every lane of the warp calls it, `have` marks the lanes that append.
template <int BLOCK, bool STAGE, int SLOTS = 16>
struct RowOut {
    LineStager<BLOCK, SLOTS> st;
    int64_t* row;
    __device__ __forceinline__ void init(int64_t* ring, int64_t* r, int tid, uint64_t pol) {
        row = r;
        if (STAGE) st.init(ring, r, tid, pol);
    }
    __device__ __forceinline__ void put(bool have, int s, int64_t v, bool last) {
        if (STAGE) st.put(have, s, v, last);
        else if (have) row[s] = v;
    }
};
template <int BLOCK, bool STAGE, int SLOTS = 16>
struct RowOutSmem {
    static constexpr int kWords = STAGE ? BLOCK * LineStager<BLOCK, SLOTS>::kPitch : 1;
};

// A/B for SURVEY section 8 (f1), the fused walk -> window pipeline: the node2vec kernel emits the skip-gram windows of
// width 5 (targets and positive windows, the two outputs of to_windows that depend on the walk) instead of the walk
// rows, so the walks never make their round trip through HBM.  Element s of a walk is the target of window s-2 and
// completes window s-4 = (w[s-4], w[s-3], w[s-1], w[s]); both streams are contiguous per walk and leave through line
// stagers like the walk row does.  Measured, not shipped: profiles/r02_summary.md.
template <int BLOCK>
struct WindowOut {
    static constexpr int kTargetSlots = 8, kPosSlots = 16;
    static constexpr int kWords = BLOCK * (LineStager<BLOCK, kTargetSlots>::kPitch + LineStager<BLOCK, kPosSlots>::kPitch);
    LineStager<BLOCK, kTargetSlots> tgt;
    LineStager<BLOCK, kPosSlots> pos;
    int64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;  // the last four elements of the walk
    int L = 0;
    __device__ __forceinline__ void init(int64_t* ring, const WalkArgs& a, int64_t i, int tid, uint64_t pol) {
        L = a.walk_length;
        const int64_t per_walk = (int64_t)L - 3;  // windows of width 5 in a row of L + 1 elements
        tgt.init(ring, a.win_target + i * per_walk, tid, pol);
        pos.init(ring + BLOCK * LineStager<BLOCK, kTargetSlots>::kPitch, a.win_pos + i * per_walk * 4, tid, pol);
    }
    __device__ __forceinline__ void put(bool have, int s, int64_t v, bool) {
        tgt.put(have && s >= 2 && s <= L - 2, s - 2, v, s == L - 2);
        pos.put4(have && s >= 4, 4 * (s - 4), h0, h1, h3, v, s == L);
        if (have) { h0 = h1; h1 = h2; h2 = h3; h3 = v; }
    }
};
template <int BLOCK, bool STAGE, int SLOTS>
__device__ __forceinline__ void init_out(RowOut<BLOCK, STAGE, SLOTS>& o, int64_t* ring, const WalkArgs& a, int64_t i, int tid, uint64_t pol) {
    o.init(ring, a.out + i * a.out_row_stride, tid, pol);
}
template <int BLOCK>
__device__ __forceinline__ void init_out(WindowOut<BLOCK>& o, int64_t* ring, const WalkArgs& a, int64_t i, int tid, uint64_t pol) {
    o.init(ring, a, i, tid, pol);
}
template <int BLOCK, bool STAGE, int SLOTS, bool WIN>
struct OutSel {
    using type = RowOut<BLOCK, STAGE, SLOTS>;
    static constexpr int kWords = RowOutSmem<BLOCK, STAGE, SLOTS>::kWords;
};
template <int BLOCK, bool STAGE, int SLOTS>
struct OutSel<BLOCK, STAGE, SLOTS, true> {
    using type = WindowOut<BLOCK>;
    static constexpr int kWords = WindowOut<BLOCK>::kWords;
};

// First-order walk: one thread per walk, two dependent gathers per step (row span from L2, then
// the chosen col_idx entry from HBM), one Philox block per four steps.
template <int BLOCK, bool STAGE, bool ROW32, bool REC>
__global__ void __launch_bounds__(BLOCK, 4) uniform_walk_kernel(const WalkArgs a) {  // four resident CTAs: what its carve-out targets (launchers below)
    __shared__ int64_t ring[RowOutSmem<BLOCK, STAGE>::kWords];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool live = i < a.n_walks;  // lanes past the end stay for the warp-collective stores
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const uint64_t wid = global_walk_id(a, i);
    RowOut<BLOCK, STAGE> o;
    o.init(ring, a.out + (live ? i : 0) * a.out_row_stride, threadIdx.x, output_policy(a.store_mode));

    int64_t v = live ? __ldg(a.targets + i) : 0;
    const int L = a.walk_length;
    o.put(live, 0, v, L == 0);
    int64_t b = 0, e = 0;
    if (REC && L > 0 && live) load_row<ROW32>(a, v, b, e, pol_keep);  // only the start node is looked up in the row index
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int s = 1; s <= L; ++s) {
        if (live) {
            const int k = (s - 1) & 3;
            if (k == 0)
                rnd = philox4x32_10(make_uint4((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)((s - 1) >> 2), 0x80000000u), a.key);
            const uint32_t r = rnd.x;
            rnd.x = rnd.y; rnd.y = rnd.z; rnd.z = rnd.w;
            if (!REC) load_row<ROW32>(a, v, b, e, pol_keep);
            int64_t xb, xe;
            v = propose<REC>(a, v, b, e, r, r * 0x9E3779B1u + (uint32_t)s, pol_stream, xb, xe);
            if (REC) { b = xb; e = xe; }
        }
        o.put(live, s, v, s == L);
    }
}

// Elements per cooperative output store piece of the node2vec kernel.  Measured in one run on the benchmark
// graph (boxes differ by +-2 %): the plain-rejection variant is 1 % faster with 64-byte pieces and the L1 they
// leave free (26.14 vs 26.43 ms on C3; option n2v_slots for the A/B), the mixture/folding variants with whole
// lines (40.0 vs 40.8 ms at p=0.5 q=2).
template <bool FOLD>
constexpr int kN2vSlots = FOLD ? 16 : 8;

// Second-order walk.  One iteration of the loop = one rejection trial of this thread's walk.
//
// FOLD (used when the return edge carries the largest weight, 1/p > max(1, 1/q)): plain rejection
// needs an envelope of 1/p over every neighbour although only the return edge reaches it, which
// costs max(1,1/q)*p times more trials than necessary (4x at p=0.25).  Following KnightKing
// (Yang et al., SOSP'19) the excess of the return edge, e = 1/p - M' with M' = max(1, 1/q), is
// folded out of the envelope: a trial draws a point uniformly from deg(v) bars of height M' plus
// one extra bar of height e that belongs to t.  The extra bar is accepted iff t is a neighbour of
// v (one membership probe, needed only when the point lands there: probability e/(deg*M'+e));
// a point in bar x at height h is accepted iff h < min(w(x), M').  Every neighbour is therefore
// still drawn with probability proportional to its node2vec weight (t: M' + e = 1/p).  This is
// only exact when no edge is stored twice, which the prepare step verifies (strict_counts).
template <int BLOCK, int MIN_CTAS, bool STAGE, bool TABLE, bool SPECULATE, bool ROW32, bool FOLD, bool REC, int SLOTS = kN2vSlots<FOLD>,
          bool WIN = false>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) node2vec_walk_kernel(const WalkArgs a) {
    __shared__ int64_t ring[OutSel<BLOCK, STAGE, SLOTS, WIN>::kWords];
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool live = i < a.n_walks;  // lanes past the end stay for the warp-collective stores
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const uint64_t wid = global_walk_id(a, i);
    typename OutSel<BLOCK, STAGE, SLOTS, WIN>::type o;
    init_out(o, ring, a, live ? i : 0, threadIdx.x, output_policy(a.store_mode));
    const int L = a.walk_length;
    const uint32_t wlo = (uint32_t)wid, whi = (uint32_t)(wid >> 32);
    // a table whose build reported an overflowing segment is not trusted: scan instead
    const uint32_t* table = (TABLE && a.table != nullptr && *a.table_failed == 0) ? a.table : nullptr;
    // `symmetric`: every stored (t -> v) has its (v -> t), proven by the pass that built the triangle Blooms
    // (which leaves them saturated and this flag down when the table build failed); then t is a neighbour of
    // v at every step and a triangle may be looked at from its far side.
    const bool symmetric = REC && TABLE && a.asymmetric != nullptr && *a.asymmetric == 0;

    int64_t t = live ? __ldg(a.targets + i) : 0;
    o.put(live, 0, t, L == 0);
    if (L == 0) return;
    int64_t tb = 0, te = 0, vb = 0, ve = 0, v = 0;
    uint32_t ctx = kBloomAll;  // Bloom of adj(t) & adj(v) for the current (t, v), or all ones
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (live) {
        load_row<ROW32>(a, t, tb, te, pol_keep);
        rnd = philox4x32_10(make_uint4(wlo, whi, 1u, 0u), a.key);
        v = propose<REC>(a, t, tb, te, rnd.x, rnd.z, pol_stream, vb, ve, ctx);  // first step is uniform (rw_cuda.cu:138)
    }
    o.put(live, 1, v, L == 1);
    if (L == 1) return;
    if (!REC && live) load_row<ROW32>(a, v, vb, ve, pol_keep);

    // One loop iteration = one rejection trial of every lane that still walks, then the collective
    // store of whatever the lanes accepted; the warp leaves the loop together.
    int s = live ? 2 : L + 1;
    uint32_t trial = 0;
    if (FOLD && a.mix && a.strict_counts[0] == a.strict_counts[1]) {
        // Two-sided mixture (q > 1, p <= q; rows without duplicate edges).  With weights scaled so that
        // a common neighbour weighs 1, every neighbour of v weighs c = 1/q, a common neighbour 1 - c
        // more and the return edge 1/p - c more:  w = c*[x in adj(v)] + (1-c)*[x in adj(v) & adj(t)] +
        // (1/p-c)*[x == t].  A trial picks one of the three terms in proportion to an upper bound of
        // its mass and samples it directly:
        //   A  c*deg(v):                 a uniform neighbour of v, always accepted (one gather);
        //   B  (1-c)*min(deg v, deg t):  a uniform entry of the SHORTER of the two rows, accepted iff it
        //                                is in the other one (gather + probe) -- the common neighbours are
        //                                proposed from the side where they are dense;
        //   C  1/p - c:                  t itself, accepted iff t is a neighbour of v (one probe).
        // Plain rejection spends q trials per step on the far neighbours' low weight; here the envelope
        // exceeds the true mass only by the misses of B, so a step costs about one gather when deg(t)
        // is much smaller than deg(v) and never more than plain rejection does.
        //
        // The three terms share ONE gather site and ONE membership site: which row is drawn from and which
        // question is asked are selected by predicates, not by branches, so the lanes of a warp issue their
        // gathers (and then their probes) together whatever term each of them picked.
        const double c = a.fold_env, back = a.fold_excess;
        uint32_t thr_a = 0, thr_ab = 0;
        bool have_thr = false, t_side = false;
        while (__any_sync(0xFFFFFFFFu, s <= L)) {
            bool accept = false;
            int64_t x = 0;
            const int s_now = s;
            if (s <= L) {
                const int64_t dv = ve - vb, dt = te - tb;
                if (!have_thr) {
                    const double mass_a = c * (double)dv;
                    const double mass_b = (1.0 - c) * (double)max((int64_t)0, min(dv, dt));
                    const double total = mass_a + mass_b + back;
                    thr_a = (uint32_t)fmin(mass_a / total * 4294967296.0, 4294967295.0);
                    thr_ab = (uint32_t)fmin((mass_a + mass_b) / total * 4294967296.0, 4294967295.0);
                    t_side = dt < dv;
                    have_thr = true;
                }
                rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, trial), a.key);
                int64_t xb = vb, xe = ve;
                uint32_t xw = kBloomAll;
                const bool stay = dv <= 0;  // no out-edge: the walk stays on v (rw_cuda.cu:25-30)
                const bool term_a = !stay && rnd.z < thr_a;
                const bool term_b = !stay && !term_a && rnd.z < thr_ab;
                const bool term_c = !stay && !term_a && !term_b;
                const bool from_t = term_b && t_side;  // B draws from the shorter row
                x = v;
                if (term_a || term_b) x = propose<REC>(a, from_t ? t : v, from_t ? tb : vb, from_t ? te : ve, rnd.x, rnd.w, pol_stream, xb, xe, xw);
                if (term_c) { x = t; xb = tb; xe = te; }
                // B asks whether x is also in the row it was NOT drawn from (the triangle Blooms look first);
                // C asks whether t is a neighbour of v at all, which a symmetric graph has already answered
                const bool ask = term_b ? (x != t && triangle_maybe(ctx, x, symmetric, xw, bloom_bit(from_t ? v : t))) : (term_c && !symmetric);
                const bool in_v = from_t || term_c;
                bool member = false;
                if (ask) member = is_member_filtered<TABLE>(a, term_c ? t : x, REC ? xe - xb : 0, in_v ? v : t, in_v ? vb : tb, in_v ? ve : te, table, pol_stream, pol_keep);
                accept = stay || term_a || member || (term_c && symmetric);
                if (accept) {
                    if (!REC && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
                    // the Bloom of the new (t, v): the word that came with x when x was drawn from adj(v); the old one
                    // after a return (adj(v) & adj(t) is the same set from either end); unknown otherwise
                    ctx = (term_a || (term_b && !from_t)) ? xw : (term_c ? ctx : kBloomAll);
                    t = v; tb = vb; te = ve;
                    v = x; vb = xb; ve = xe;
                    ++s;
                    trial = 0;
                    have_thr = false;
                } else {
                    ++trial;
                }
            }
            o.put(accept, s_now, x, s_now == L);
        }
        return;
    }
    if (FOLD && !a.mix && a.strict_counts[0] == a.strict_counts[1]) {
        // rows are strictly increasing (no duplicate edges): the folded envelope is exact
        const uint64_t fthr_any = min(a.fthr1, a.fthr2), fthr_top = max(a.fthr1, a.fthr2);
        uint32_t thr_extra = 0;  // P(point lands in the extra bar) for the current v, scaled by 2^32
        bool have_extra = false;
        while (__any_sync(0xFFFFFFFFu, s <= L)) {
            bool accept = false;
            int64_t x = 0;
            const int s_now = s;
            if (s <= L) {
                if (!have_extra) {
                    const double area = (double)(ve - vb) * a.fold_env + a.fold_excess;
                    thr_extra = (uint32_t)fmin(a.fold_excess / area * 4294967296.0, 4294967295.0);
                    have_extra = true;
                }
                rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, trial), a.key);
                int64_t xb = vb, xe = ve;
                uint32_t xw = kBloomAll;
                const bool stay = ve <= vb;  // no out-edge: the walk stays on v (rw_cuda.cu:25-30), nothing to sample
                const bool extra = !stay && rnd.z < thr_extra;  // the point fell into the extra bar of t
                x = v;
                if (!stay && !extra) x = propose<REC>(a, v, vb, ve, rnd.x, rnd.w, pol_stream, xb, xe, xw);
                if (extra) { x = t; xb = tb; xe = te; xw = ctx; }  // adj(v) & adj(t) is the same set from either end
                const uint32_t u = rnd.y;
                // one membership site: "t in adj(v)?" for the extra bar (answered by symmetry when the graph has it),
                // "x in adj(t)?" when the draw alone does not decide a regular bar
                const bool regular = !stay && !extra && x != t && u >= fthr_any && u < fthr_top;
                const bool ask = extra ? !symmetric : (regular && triangle_maybe(ctx, x, symmetric, xw, bloom_bit(t)));
                bool member = false;
                if (ask) member = is_member_filtered<TABLE>(a, extra ? t : x, REC ? xe - xb : 0, extra ? v : t, extra ? vb : tb, extra ? ve : te, table, pol_stream, pol_keep);
                if (stay) accept = true;
                else if (extra) accept = symmetric || member;
                else if (x == t || u < fthr_any) accept = true;
                else if (u >= fthr_top) accept = false;
                else accept = u < (member ? a.fthr1 : a.fthr2);
                if (accept) {
                    if (!REC && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
                    t = v; tb = vb; te = ve;
                    v = x; vb = xb; ve = xe;
                    ctx = xw;
                    ++s;
                    trial = 0;
                    have_extra = false;
                } else {
                    ++trial;
                }
            }
            o.put(accept, s_now, x, s_now == L);
        }
        return;
    }
    const uint64_t thr_any = min(a.thr0, min(a.thr1, a.thr2));  // below this every class accepts
    const uint64_t thr_far = max(a.thr1, a.thr2);               // at or above this only x == t can accept
    while (__any_sync(0xFFFFFFFFu, s <= L)) {
        bool accept = false;
        int64_t x = 0;
        const int s_now = s;
        if (s <= L) {
            uint32_t r, u, r_hi;
            if ((trial & 1u) == 0u) {
                rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, trial >> 1), a.key);
                r = rnd.x; u = rnd.y; r_hi = rnd.z;
            } else {
                r = rnd.z; u = rnd.w; r_hi = rnd.x;
            }
            int64_t xb = 0, xe = 0;
            uint32_t xw;
            x = propose<REC>(a, v, vb, ve, r, r_hi, pol_stream, xb, xe, xw);
            const bool back = (x == t);
            const bool possible = back ? (u < a.thr0) : (u < thr_far);
            if (!REC && SPECULATE && possible && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
            if (u < thr_any) accept = true;
            else if (back) accept = u < a.thr0;
            else if (!possible) accept = false;
            else if (a.thr1 == a.thr2) accept = true;  // q == 1: membership cannot change the answer
            else {
                const bool member = triangle_maybe(ctx, x, symmetric, xw, bloom_bit(t)) &&
                                    is_member_filtered<TABLE>(a, x, REC ? xe - xb : 0, t, tb, te, table, pol_stream, pol_keep);
                accept = u < (member ? a.thr1 : a.thr2);
            }
            if (accept) {
                if (!REC && !SPECULATE && s < L) load_row<ROW32>(a, x, xb, xe, pol_keep);
                t = v; tb = vb; te = ve;
                v = x; vb = xb; ve = xe;
                ctx = xw;
                ++s;
                trial = 0;
            } else {
                ++trial;
            }
        }
        o.put(accept, s_now, x, s_now == L);
    }
}

// ------------------------------------------------------------------------------------------ A/B: one warp per walk
// The design BASELINE.json's north star sketches: a warp owns a walk; at a node of up to kCdfMax neighbours it reads the
// whole neighbourhood cooperatively (one coalesced record load per lane), weighs every neighbour (1/p, 1, 1/q -- membership
// through the same Blooms and table as the shipped kernel), and samples from the exact CDF by a shuffle prefix sum; above
// that it falls back to rejection, which all lanes run in lock step (same addresses: one request per load).  Exact (same
// law, other draws: statistics tested).  Kept as option n2v_warp for the A/B in profiles/: with 32 threads per walk an SM
// holds 32 times fewer walks in flight, and on every graph measured a step costs more lines than a rejection trial.
constexpr int kCdfMax = 64;

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) node2vec_warp_walk_kernel(const WalkArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (i >= a.n_walks) return;  // the whole warp leaves together
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    const uint64_t wid = global_walk_id(a, i);
    const uint32_t wlo = (uint32_t)wid, whi = (uint32_t)(wid >> 32);
    const uint32_t* table = (a.table != nullptr && *a.table_failed == 0) ? a.table : nullptr;
    const bool symmetric = a.asymmetric != nullptr && *a.asymmetric == 0;
    int64_t* row = a.out + i * a.out_row_stride;
    const int L = a.walk_length;
    int64_t t = __ldg(a.targets + i);
    if (lane == 0) stg64_hint(row, t, pol_stream);
    if (L == 0) return;
    int64_t tb = 0, te = 0, vb = 0, ve = 0;
    uint32_t ctx = kBloomAll;
    load_row<true>(a, t, tb, te, pol_keep);
    uint4 rnd = philox4x32_10(make_uint4(wlo, whi, 1u, 0u), a.key);
    int64_t v = propose<true>(a, t, tb, te, rnd.x, rnd.z, pol_stream, vb, ve, ctx);
    if (lane == 0) stg64_hint(row + 1, v, pol_stream);
    const uint64_t thr_any = min(a.thr0, min(a.thr1, a.thr2)), thr_far = max(a.thr1, a.thr2);
    for (int s = 2; s <= L; ++s) {
        const int64_t dv = ve - vb;
        int64_t x = v, xb = vb, xe = ve;
        uint32_t xw = kBloomAll;
        if (dv > 0 && dv <= kCdfMax) {
            // exact CDF over the neighbourhood: two rounds of 32 lanes
            rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, 0u), a.key);
            const double u = (double)((((uint64_t)rnd.x << 32) | rnd.y) >> 11) * (1.0 / 9007199254740992.0);
            double prefix[2] = {0.0, 0.0}, carry = 0.0;
            uint4 rec[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int64_t j = (int64_t)r * 32 + lane;
                const bool valid = j < dv;
                rec[r] = valid ? ldg_u32x4_hint(a.records + vb + j, pol_stream) : make_uint4(0, 0, 0, 0);
                double w = 0.0;
                if (valid) {
                    const int64_t xj = rec[r].y != 0 ? (int64_t)rec[r].x : (int64_t)(((uint64_t)rec[r].w << 32) | rec[r].x);
                    const uint32_t wj = rec[r].y != 0 ? rec[r].w : kBloomAll;
                    if (xj == t) w = a.w_back;
                    else if (a.w_common == a.w_far) w = a.w_far;
                    else w = (triangle_maybe(ctx, xj, symmetric, wj, bloom_bit(t)) &&
                              is_member_filtered<true>(a, xj, 0, t, tb, te, table, pol_stream, pol_keep)) ? a.w_common : a.w_far;
                }
                double incl = w;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += o;
                }
                prefix[r] = carry + incl;
                carry = __shfl_sync(0xFFFFFFFFu, prefix[r], 31);
            }
            const double target = u * carry;
            // the first neighbour whose inclusive prefix exceeds the target (the last valid one if rounding leaves none)
            const unsigned hit0 = __ballot_sync(0xFFFFFFFFu, lane < dv && prefix[0] > target);
            const unsigned hit1 = __ballot_sync(0xFFFFFFFFu, (int64_t)32 + lane < dv && prefix[1] > target);
            int round = 0, src = 0;
            if (hit0) src = __ffs(hit0) - 1;
            else if (hit1) { round = 1; src = __ffs(hit1) - 1; }
            else { round = dv > 32 ? 1 : 0; src = (int)((dv - 1) & 31); }
            const uint4 pick = make_uint4(__shfl_sync(0xFFFFFFFFu, round ? rec[1].x : rec[0].x, src),
                                          __shfl_sync(0xFFFFFFFFu, round ? rec[1].y : rec[0].y, src),
                                          __shfl_sync(0xFFFFFFFFu, round ? rec[1].z : rec[0].z, src),
                                          __shfl_sync(0xFFFFFFFFu, round ? rec[1].w : rec[0].w, src));
            xb = (int64_t)pick.z;
            xe = xb + (int64_t)pick.y;
            if (pick.y != 0) { x = (int64_t)pick.x; xw = pick.w; }
            else x = (int64_t)(((uint64_t)pick.w << 32) | pick.x);
        } else if (dv > 0) {
            // rejection, every lane in lock step (rw_cuda.cu:146-179)
            for (uint32_t trial = 0;; ++trial) {
                rnd = philox4x32_10(make_uint4(wlo, whi, (uint32_t)s, trial), a.key);
                x = propose<true>(a, v, vb, ve, rnd.x, rnd.z, pol_stream, xb, xe, xw);
                const uint32_t u = rnd.y;
                bool accept;
                if (u < thr_any) accept = true;
                else if (x == t) accept = u < a.thr0;
                else if (u >= thr_far) accept = false;
                else if (a.thr1 == a.thr2) accept = true;
                else accept = u < ((triangle_maybe(ctx, x, symmetric, xw, bloom_bit(t)) &&
                                    is_member_filtered<true>(a, x, 0, t, tb, te, table, pol_stream, pol_keep)) ? a.thr1 : a.thr2);
                if (accept) break;
            }
        }
        t = v; tb = vb; te = ve;
        v = x; vb = xb; ve = xe;
        ctx = xw;
        if (lane == 0) stg64_hint(row + s, v, pol_stream);
    }
}

// ------------------------------------------------------------------------------------------ launchers
// Shared memory and L1 share one 256 KB array per SM, and the L1 side is where the in-flight
// gathers land.  Measured on the benchmark graph (profiles/r01_summary.md): the first-order kernel
// is fastest with four resident CTAs (1024 threads, 164 KB configuration) and loses 40 % when its
// staging rings are allowed to take 209 KB (six CTAs, 28 KB of L1 left), so its carve-out is pinned.
// The node2vec kernel is flat between 768 and 1280 threads and keeps the driver's choice.
constexpr int kUniformCarveoutKb = 164;

template <typename K>
static void set_carveout(K kernel, int64_t kb) {
    int pct = kb <= 0 ? (int)cudaSharedmemCarveoutDefault : (int)(kb * 100 / 228);  // rounded down: the driver rounds up to a configuration
    if (pct > 100) pct = 100;
    // one attribute call per kernel, device and value, not one per launch (K is the same pointer type for every
    // instantiation, so the cache is keyed by the function's address)
    static std::mutex mu;
    static std::unordered_map<uint64_t, int> last;
    int d = 0;
    cudaGetDevice(&d);
    const uint64_t key = (uint64_t)(uintptr_t)reinterpret_cast<const void*>(kernel) ^ ((uint64_t)(uint32_t)d << 56);
    std::lock_guard<std::mutex> lock(mu);
    auto it = last.find(key);
    if (it != last.end() && it->second == pct) return;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) == cudaSuccess) last[key] = pct;
}
#define TRW_LAUNCH(kernel, carveout_kb, grid, block, st, args)                                              \
    do {                                                                                                    \
        set_carveout(kernel, options().smem_carveout_kb > 0 ? options().smem_carveout_kb : (carveout_kb));  \
        kernel<<<grid, block, 0, st>>>(args);                                                               \
    } while (0)

template <bool STAGE, bool ROW32, bool REC>
static void launch_uniform(const WalkArgs& a, cudaStream_t st) {
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    TRW_LAUNCH((uniform_walk_kernel<BLOCK, STAGE, ROW32, REC>), STAGE ? kUniformCarveoutKb : 0, grid, BLOCK, st, a);
}

template <int MIN_CTAS, bool STAGE, bool TABLE, bool ROW32, bool REC>
static void launch_n2v3(const WalkArgs& a, bool speculate, bool fold, cudaStream_t st) {
    constexpr int BLOCK = 256;
    const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
    if (fold) TRW_LAUNCH((node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, false, ROW32, true, REC>), 0, grid, BLOCK, st, a);
    else if (speculate && !REC) TRW_LAUNCH((node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, true, ROW32, false, false>), 0, grid, BLOCK, st, a);
    else if (REC && STAGE && options().n2v_slots == 16) TRW_LAUNCH((node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, false, ROW32, false, REC, 16>), 0, grid, BLOCK, st, a);
    else TRW_LAUNCH((node2vec_walk_kernel<BLOCK, MIN_CTAS, STAGE, TABLE, false, ROW32, false, REC>), 0, grid, BLOCK, st, a);
}

template <bool STAGE, bool TABLE, bool ROW32, bool REC>
static void launch_n2v2(const WalkArgs& a, bool speculate, bool fold, int min_ctas, cudaStream_t st) {
    if (min_ctas >= 6) launch_n2v3<6, STAGE, TABLE, ROW32, REC>(a, speculate, fold, st);
    else if (min_ctas == 5) launch_n2v3<5, STAGE, TABLE, ROW32, REC>(a, speculate, fold, st);
    else launch_n2v3<4, STAGE, TABLE, ROW32, REC>(a, speculate, fold, st);
}

static void launch_n2v(const WalkArgs& a, bool stage, bool table, bool row32, bool rec, bool speculate, bool fold,
                       int min_ctas, cudaStream_t st) {
    if (!stage) {  // plain 8-byte stores: A/B path only, kept to one variant per table mode
        if (table) launch_n2v3<4, false, true, false, false>(a, speculate, fold, st);
        else launch_n2v3<4, false, false, false, false>(a, speculate, fold, st);
        return;
    }
    if (rec) {  // edge records imply the uint32 row index (used for the start nodes only)
        if (table) launch_n2v2<true, true, true, true>(a, speculate, fold, min_ctas, st);
        else launch_n2v2<true, false, true, true>(a, speculate, fold, min_ctas, st);
        return;
    }
    if (table) {
        if (row32) launch_n2v2<true, true, true, false>(a, speculate, fold, min_ctas, st);
        else launch_n2v2<true, true, false, false>(a, speculate, fold, min_ctas, st);
    } else {
        if (row32) launch_n2v2<true, false, true, false>(a, speculate, fold, min_ctas, st);
        else launch_n2v2<true, false, false, false>(a, speculate, fold, min_ctas, st);
    }
}

static uint64_t threshold(double prob) {
    double t = prob * 4294967296.0;
    if (!(t > 0.0)) return 0;
    if (t >= 4294967296.0) return 4294967296ull;
    return (uint64_t)t;
}

// Optional persisting-L2 window over the row index (experiment: option persist_row_ptr; measured
// slower than the per-load policies on B200, profiles/r01_probe2_options.json).
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

// ------------------------------------------------------------------------------------------ graph side
// Everything the walk derives from the graph alone: the uint32 row index, the membership table,
// the strict-rows counters and the edge records, built into `workspace` (csr_workspace_layout with
// the same `uniform` and `want_records`).  Independent of p, q, the seed and the start nodes, so one
// prepared graph serves any number of walk calls (trw_csr_graph_* keeps it across calls;
// trw_walk_csr builds it per call).
int csr_graph_prepare(CsrGraph* g, IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz,
                      bool uniform, bool want_table, bool want_strict, bool want_records, void* workspace,
                      size_t workspace_bytes, int device, cudaStream_t st, int64_t bloom_cap) {
    if (n_nodes < 0 || nnz < 0) { set_error("trw_walk_csr: negative size"); return TRW_ERR_ARG; }
    if (!row_ptr || (nnz > 0 && !col_idx)) { set_error("trw_walk_csr: null pointer"); return TRW_ERR_ARG; }
    const Options& opt = options();
    *g = CsrGraph{};
    g->row_ptr = row_ptr; g->col_idx = col_idx; g->n_nodes = n_nodes; g->nnz = nnz; g->device = device;
    if (workspace == nullptr) return TRW_OK;  // reference-style path: int64 row_ptr, linear-scan membership
    const CsrWorkspace w = csr_workspace_layout(n_nodes, nnz, uniform, want_records);
    if (w.total == 0) return TRW_OK;
    if (workspace_bytes < w.total || ((uintptr_t)workspace & 255)) {
        set_error("trw_walk_csr: workspace needs %zu bytes at 256-byte alignment (got %zu)", w.total, workspace_bytes);
        return TRW_ERR_WORKSPACE;
    }
    timing_begin(0, st);
    const int rc = csr_prepare_device(row_ptr, col_idx, n_nodes, nnz, workspace, w, want_table && opt.n2v_table != 0,
                                      opt.row32 != 0, want_strict, want_records && opt.stage_output != 0,
                                      (int)opt.build_mode, device, st, &g->prepared, bloom_cap);
    timing_end(0, st);
    return rc;
}

// Walk-side parameters on top of a prepared graph: the acceptance thresholds of (p, q), the
// Philox key, and which kernel variant what the graph holds allows.
int csr_walk_plan(CsrWalkPlan* plan, const CsrGraph& g, double p, double q, int walk_length, int64_t seed) {
    if (walk_length < 0) { set_error("trw_walk_csr: negative walk_length"); return TRW_ERR_ARG; }
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_csr: p and q must be positive"); return TRW_ERR_ARG; }
    const Options& opt = options();
    WalkArgs& a = plan->a;
    a = WalkArgs{};
    a.row_ptr = g.row_ptr; a.col_idx = g.col_idx; a.n_nodes = g.n_nodes; a.nnz = g.nnz;
    a.walk_length = walk_length; a.key = philox_key(seed, kTagWalkCsr);
    a.store_mode = (int)opt.store_mode;
    a.table = g.prepared.table;
    a.table_failed = g.prepared.table_failed;
    a.row32 = g.prepared.row32;
    a.records = g.prepared.records;
    a.filter = g.prepared.filter;
    a.asymmetric = g.prepared.asymmetric;
    a.strict_counts = g.prepared.strict_counts;
    plan->device = g.device;
    plan->uniform = (p == 1.0 && q == 1.0);  // rw_cuda.cu:226
    plan->stage = opt.stage_output != 0;
    plan->persist = opt.persist_row_ptr != 0;
    // measured on the benchmark graphs: 5 CTAs/SM with edge records, 4 without (the row lookups like a larger L1)
    plan->min_ctas = opt.n2v_min_ctas > 0 ? (int)opt.n2v_min_ctas : (a.records != nullptr ? 5 : 4);
    plan->speculate = false;
    plan->fold = false;
    plan->table = false;
    if (!plan->uniform) {
        const double mx = fmax(fmax(1.0 / p, 1.0), 1.0 / q);  // rw_cuda.cu:119-123
        const double p0 = 1.0 / p / mx, p1 = 1.0 / mx, p2 = 1.0 / q / mx;
        a.thr0 = threshold(p0); a.thr1 = threshold(p1); a.thr2 = threshold(p2);
        a.w_back = 1.0 / p; a.w_common = 1.0; a.w_far = 1.0 / q;
        // Fetch row_ptr[x] before the verdict only when most proposals are accepted anyway.
        plan->speculate = opt.n2v_speculate < 0 ? (fmin(p1, p2) >= 0.5) : (opt.n2v_speculate != 0);
        // Return-edge folding applies when 1/p is the strict maximum of the three weights and the
        // graph side has verified (strict_counts) that no edge is stored twice.
        const double env = fmax(1.0, 1.0 / q);
        if (opt.n2v_mix != 0 && q > 1.0 && p <= q && a.strict_counts != nullptr) {
            // two-sided mixture (see node2vec_walk_kernel); it rides on the strict-rows kernel variant
            plan->fold = true;
            a.mix = 1;
            a.fold_env = 1.0 / q;
            a.fold_excess = 1.0 / p - 1.0 / q;
        } else if (opt.n2v_fold != 0 && 1.0 / p > env && a.strict_counts != nullptr) {
            plan->fold = true;
            a.fold_env = env;
            a.fold_excess = 1.0 / p - env;
            a.fthr1 = threshold(1.0 / env);
            a.fthr2 = threshold(1.0 / q / env);
        }
        // membership matters when common and far neighbours weigh differently (q != 1), and for the extra bar of the
        // folded return edge ("is t a neighbour of v?"), which a scan of a hub row would make the whole cost of q == 1
        plan->table = a.table != nullptr && (a.thr1 != a.thr2 || plan->fold);
        if (!plan->table) a.table = nullptr;
        if (!plan->table || opt.edge_filter_mb <= 0) a.filter = EdgeFilter{};
        if (!plan->table) a.asymmetric = nullptr;
    }
    return TRW_OK;
}

// What a one-shot call builds for (p, q): the table only when membership can change a verdict,
// the strict-rows check only when folding applies, the records only when the walk is long enough
// to pay for them (one extra pass over col_idx per call; a kept graph always has them).
void csr_one_shot_needs(double p, double q, int64_t nnz, int64_t n_walks, int walk_length, bool* uniform,
                        bool* want_table, bool* want_strict, bool* want_records) {
    const Options& opt = options();
    *uniform = (p == 1.0 && q == 1.0);
    *want_strict = !*uniform && ((opt.n2v_fold != 0 && 1.0 / p > fmax(1.0, 1.0 / q)) || (opt.n2v_mix != 0 && q > 1.0 && p <= q));
    *want_table = !*uniform && (q != 1.0 || *want_strict);  // thr1 == thr2 iff q == 1; folding asks "t in adj(v)?"
    // Records save about a tenth of the walk's gathers and cost one pass over col_idx: worth building
    // per call when the walk fetches more than ~3 random lines per CSR entry (measured break-even on
    // the benchmark graphs; lines per step: 1 first-order, ~1.5 for q <= 1, ~0.85 q + 0.8 for q > 1).
    const double lines_per_step = *uniform ? 1.0 : (q <= 1.0 ? 1.5 : 0.85 * q + 0.8);
    const double lines = (double)n_walks * (double)walk_length * lines_per_step;
    *want_records = opt.records > 0 || (opt.records < 0 && lines >= 3.2 * (double)nnz);
}

int csr_walk_launch(const CsrWalkPlan& plan, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                    int64_t* out, int64_t out_row_stride, cudaStream_t st, int64_t id_block, int64_t id_stride) {
    if (n_walks <= 0) return TRW_OK;
    if (id_block < 0 || (id_block > 0 && id_stride < id_block)) { set_error("walk ids: stride must be at least one block"); return TRW_ERR_ARG; }
    WalkArgs a = plan.a;
    a.targets = targets; a.n_walks = n_walks; a.walk_id_offset = walk_id_offset;
    a.id_block = id_block; a.id_stride = id_stride;
    a.out = out; a.out_row_stride = out_row_stride;
    const bool stage = plan.stage && (((uintptr_t)out & 7) == 0);
    const bool row32 = a.row32 != nullptr;
    const bool rec = a.records != nullptr && row32 && stage;
    if (plan.persist) {
        if (row32) set_row_window(st, a.row32, (size_t)(a.n_nodes + 1) * 4, plan.device, true);
        else set_row_window(st, a.row_ptr.base, (size_t)(a.n_nodes + 1) << a.row_ptr.shift, plan.device, true);
    }
    timing_begin(1, st);
    if (a.win_target != nullptr) {  // A/B: fused walk -> skip-gram windows (csr_walk_windows5 has checked the preconditions)
        constexpr int BLOCK = 128;  // two stagers per thread: 26 KB per CTA, six CTAs (768 threads) per SM
        const unsigned grid = (unsigned)((a.n_walks + BLOCK - 1) / BLOCK);
        TRW_LAUNCH((node2vec_walk_kernel<BLOCK, 6, true, true, false, true, false, true, 16, true>), 0, grid, BLOCK, st, a);
    } else if (!plan.uniform && options().n2v_warp != 0 && rec && plan.table) {  // A/B: the warp-per-walk design (see node2vec_warp_walk_kernel)
        constexpr int BLOCK = 256;
        const int64_t blocks = (a.n_walks * 32 + BLOCK - 1) / BLOCK;
        node2vec_warp_walk_kernel<BLOCK><<<(unsigned)blocks, BLOCK, 0, st>>>(a);
    } else if (plan.uniform) {
        if (!stage) launch_uniform<false, false, false>(a, st);
        else if (rec) launch_uniform<true, true, true>(a, st);
        else if (row32) launch_uniform<true, true, false>(a, st);
        else launch_uniform<true, false, false>(a, st);
    } else {
        launch_n2v(a, stage, plan.table, row32 && stage, rec, plan.speculate, plan.fold, plan.min_ctas, st);
    }
    timing_end(1, st);
    count_launch(1);
    const int rc = check_cuda(cudaGetLastError(), "walk kernel launch");
    if (plan.persist) set_row_window(st, nullptr, 0, plan.device, false);
    return rc;
}

}  // namespace trw

using namespace trw;

// Opaque handle of the C ABI: a prepared graph plus the flags it was built with.
struct trw_csr_graph {
    CsrGraph g;
};

namespace trw {
int csr_graph_of_handle(const trw_csr_graph* h, CsrGraph* out) {
    if (!h || !out) { set_error("null prepared graph"); return TRW_ERR_ARG; }
    *out = h->g;
    return TRW_OK;
}
}  // namespace trw

extern "C" size_t trw_walk_csr_workspace_bytes(int64_t n_nodes, int64_t nnz, double p, double q) {
    if (n_nodes < 0 || nnz < 0) return 0;
    // sized for the larger of the two one-shot layouts (with records), so the auto rule never outgrows it
    return csr_workspace_layout(n_nodes, nnz, p == 1.0 && q == 1.0, options().records != 0).total;
}

extern "C" size_t trw_walk_csr_workspace_bytes_for(int64_t n_nodes, int64_t nnz, double p, double q, int64_t n_walks,
                                                   int walk_length) {
    if (n_nodes < 0 || nnz < 0 || n_walks < 0 || walk_length < 0) return 0;
    bool uniform, want_table, want_strict, want_records;
    csr_one_shot_needs(p, q, nnz, n_walks, walk_length, &uniform, &want_table, &want_strict, &want_records);
    return csr_workspace_layout(n_nodes, nnz, uniform, want_records).total;
}

static int walk_args_check(const char* fn, const int64_t* targets, int64_t n_walks, int walk_length, int64_t* out,
                           int64_t out_row_stride) {
    if (n_walks < 0 || walk_length < 0 || out_row_stride < (int64_t)walk_length + 1) {
        set_error("%s: negative size or out_row_stride < walk_length+1", fn);
        return TRW_ERR_ARG;
    }
    if (n_walks > 0 && (!targets || !out)) { set_error("%s: null pointer", fn); return TRW_ERR_ARG; }
    return TRW_OK;
}

static int elem_bytes_ok(const char* fn, int row_ptr_bytes, int col_idx_bytes) {
    if ((row_ptr_bytes != 4 && row_ptr_bytes != 8) || (col_idx_bytes != 4 && col_idx_bytes != 8)) {
        set_error("%s: CSR elements must be 4 or 8 bytes wide (got %d, %d)", fn, row_ptr_bytes, col_idx_bytes);
        return TRW_ERR_ARG;
    }
    return TRW_OK;
}

static int walk_csr_one_shot(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz,
                             const int64_t* targets, int64_t n_walks, int64_t walk_id_offset, double p, double q,
                             int walk_length, int64_t seed, int64_t* out, int64_t out_row_stride, void* workspace,
                             size_t workspace_bytes, int device, void* stream) {
    int rc = walk_args_check("trw_walk_csr", targets, n_walks, walk_length, out, out_row_stride);
    if (rc) return rc;
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_walk_csr: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (!(p > 0.0) || !(q > 0.0)) { set_error("trw_walk_csr: p and q must be positive"); return TRW_ERR_ARG; }
    bool uniform, want_table, want_strict, want_records;
    csr_one_shot_needs(p, q, nnz, n_walks, walk_length, &uniform, &want_table, &want_strict, &want_records);
    // the records are dropped rather than refused when the caller sized the workspace without them
    if (want_records && workspace && workspace_bytes < csr_workspace_layout(n_nodes, nnz, uniform, true).total)
        want_records = false;
    CsrGraph g;
    rc = csr_graph_prepare(&g, row_ptr, col_idx, n_nodes, nnz, uniform, want_table, want_strict, want_records, workspace,
                           workspace_bytes, d, st);
    if (rc) return rc;
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, g, p, q, walk_length, seed);
    if (rc) return rc;
    return csr_walk_launch(plan, targets, n_walks, walk_id_offset, out, out_row_stride, st);
}

extern "C" int trw_walk_csr(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                            const int64_t* targets, int64_t n_walks, int64_t walk_id_offset, double p, double q,
                            int walk_length, int64_t seed, int64_t* out, int64_t out_row_stride, void* workspace,
                            size_t workspace_bytes, int device, void* stream) {
    return walk_csr_one_shot(IdxPtr(row_ptr), IdxPtr(col_idx), n_nodes, nnz, targets, n_walks, walk_id_offset, p, q, walk_length, seed,
                             out, out_row_stride, workspace, workspace_bytes, device, stream);
}

extern "C" int trw_walk_csr_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                                  int64_t n_nodes, int64_t nnz, const int64_t* targets, int64_t n_walks,
                                  int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed, int64_t* out,
                                  int64_t out_row_stride, void* workspace, size_t workspace_bytes, int device, void* stream) {
    const int rc = elem_bytes_ok("trw_walk_csr_typed", row_ptr_bytes, col_idx_bytes);
    if (rc) return rc;
    return walk_csr_one_shot(IdxPtr(row_ptr, row_ptr_bytes), IdxPtr(col_idx, col_idx_bytes), n_nodes, nnz, targets, n_walks,
                             walk_id_offset, p, q, walk_length, seed, out, out_row_stride, workspace, workspace_bytes, device, stream);
}

extern "C" size_t trw_csr_graph_workspace_bytes(int64_t n_nodes, int64_t nnz) {
    if (n_nodes < 0 || nnz < 0) return 0;
    return csr_workspace_layout(n_nodes, nnz, false, options().records != 0).total;
}

extern "C" int trw_csr_graph_prepare(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                                     void* workspace, size_t workspace_bytes, int device, void* stream,
                                     trw_csr_graph** out_graph) {
    return trw_csr_graph_prepare_ex(row_ptr, col_idx, n_nodes, nnz, workspace, workspace_bytes, device, stream, -1, out_graph);
}

static int graph_prepare_impl(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, void* workspace, size_t workspace_bytes,
                              int device, void* stream, int64_t bloom_cap, trw_csr_graph** out_graph);

extern "C" int trw_csr_graph_prepare_ex(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                                        void* workspace, size_t workspace_bytes, int device, void* stream,
                                        int64_t bloom_cap, trw_csr_graph** out_graph) {
    return graph_prepare_impl(IdxPtr(row_ptr), IdxPtr(col_idx), n_nodes, nnz, workspace, workspace_bytes, device, stream, bloom_cap,
                              out_graph);
}

extern "C" int trw_csr_graph_prepare_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                                           int64_t n_nodes, int64_t nnz, void* workspace, size_t workspace_bytes, int device,
                                           void* stream, int64_t bloom_cap, trw_csr_graph** out_graph) {
    const int rc = elem_bytes_ok("trw_csr_graph_prepare_typed", row_ptr_bytes, col_idx_bytes);
    if (rc) return rc;
    return graph_prepare_impl(IdxPtr(row_ptr, row_ptr_bytes), IdxPtr(col_idx, col_idx_bytes), n_nodes, nnz, workspace, workspace_bytes,
                              device, stream, bloom_cap, out_graph);
}

static int graph_prepare_impl(IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz, void* workspace, size_t workspace_bytes,
                              int device, void* stream, int64_t bloom_cap, trw_csr_graph** out_graph) {
    if (!out_graph) { set_error("trw_csr_graph_prepare: null out_graph"); return TRW_ERR_ARG; }
    *out_graph = nullptr;
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_csr_graph_prepare: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    if (n_nodes > 0 && nnz > 0 && workspace == nullptr) {
        set_error("trw_csr_graph_prepare: a prepared graph needs its workspace (trw_csr_graph_workspace_bytes)");
        return TRW_ERR_WORKSPACE;
    }
    trw_csr_graph* h = new (std::nothrow) trw_csr_graph();
    if (!h) { set_error("trw_csr_graph_prepare: out of host memory"); return TRW_ERR_ARG; }
    const int rc = csr_graph_prepare(&h->g, row_ptr, col_idx, n_nodes, nnz, /*uniform=*/false, /*want_table=*/true,
                                     /*want_strict=*/true, /*want_records=*/options().records != 0, workspace,
                                     workspace_bytes, d, (cudaStream_t)stream, bloom_cap < 0 ? options().edge_bloom_cap : bloom_cap);
    if (rc) { delete h; return rc; }
    *out_graph = h;
    return TRW_OK;
}

extern "C" int trw_walk_csr_prepared(const trw_csr_graph* graph, const int64_t* targets, int64_t n_walks,
                                     int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                                     int64_t* out, int64_t out_row_stride, void* stream) {
    if (!graph) { set_error("trw_walk_csr_prepared: null graph"); return TRW_ERR_ARG; }
    int rc = walk_args_check("trw_walk_csr_prepared", targets, n_walks, walk_length, out, out_row_stride);
    if (rc) return rc;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(graph->g.device);
    if (!guard.ok) { set_error("trw_walk_csr_prepared: cudaSetDevice(%d) failed", graph->g.device); return TRW_ERR_DEVICE; }
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, graph->g, p, q, walk_length, seed);
    if (rc) return rc;
    return csr_walk_launch(plan, targets, n_walks, walk_id_offset, out, out_row_stride, (cudaStream_t)stream);
}

// The same walk over a prepared graph whose CSR arrays now live at (row_ptr, col_idx): for callers that have
// verified (trw_csr_checksum) that these arrays hold what was prepared.  Nothing of the handle is changed.
extern "C" int trw_walk_csr_prepared_at(const trw_csr_graph* graph, const void* row_ptr, const void* col_idx,
                                        const int64_t* targets, int64_t n_walks, int64_t walk_id_offset, int64_t walk_id_block,
                                        int64_t walk_id_stride, double p, double q, int walk_length, int64_t seed, int64_t* out,
                                        int64_t out_row_stride, void* stream) {
    if (!graph) { set_error("trw_walk_csr_prepared_at: null graph"); return TRW_ERR_ARG; }
    if (!row_ptr || (graph->g.nnz > 0 && !col_idx)) { set_error("trw_walk_csr_prepared_at: null pointer"); return TRW_ERR_ARG; }
    int rc = walk_args_check("trw_walk_csr_prepared_at", targets, n_walks, walk_length, out, out_row_stride);
    if (rc) return rc;
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(graph->g.device);
    if (!guard.ok) { set_error("trw_walk_csr_prepared_at: cudaSetDevice(%d) failed", graph->g.device); return TRW_ERR_DEVICE; }
    CsrGraph g = graph->g;  // the arrays move, their element widths are those the graph was prepared with
    g.row_ptr.base = reinterpret_cast<const char*>(row_ptr);
    g.col_idx.base = reinterpret_cast<const char*>(col_idx);
    CsrWalkPlan plan;
    rc = csr_walk_plan(&plan, g, p, q, walk_length, seed);
    if (rc) return rc;
    return csr_walk_launch(plan, targets, n_walks, walk_id_offset, out, out_row_stride, (cudaStream_t)stream, walk_id_block,
                           walk_id_stride);
}

// Adds the triangle Blooms (member_table.cuh) to a graph that was prepared without them: one pass over the
// edge records in the graph's workspace.  `cap` <= 0 selects option edge_bloom_cap.
extern "C" int trw_csr_graph_add_blooms(trw_csr_graph* graph, const void* row_ptr, const void* col_idx, int64_t cap,
                                        void* stream) {
    if (!graph) { set_error("trw_csr_graph_add_blooms: null graph"); return TRW_ERR_ARG; }
    DeviceGuard guard(graph->g.device);
    if (!guard.ok) { set_error("trw_csr_graph_add_blooms: cudaSetDevice(%d) failed", graph->g.device); return TRW_ERR_DEVICE; }
    if (cap <= 0) cap = options().edge_bloom_cap;
    if (cap <= 0) return TRW_OK;
    timing_begin(0, (cudaStream_t)stream);
    IdxPtr ci = graph->g.col_idx;
    if (col_idx) ci.base = reinterpret_cast<const char*>(col_idx);
    const int rc = csr_add_blooms(&graph->g.prepared, ci, graph->g.n_nodes, graph->g.nnz, cap,
                                  graph->g.device, (cudaStream_t)stream);
    timing_end(0, (cudaStream_t)stream);
    (void)row_ptr;
    return rc;
}

// A/B entry of the fused walk -> window pipeline (WindowOut): node2vec walks of a kept graph whose skip-gram windows of
// width 5 go straight to target[n_walks * (L-3)] and pos[n_walks * (L-3), 4]; the walks themselves are not written.
extern "C" int trw_walk_csr_prepared_windows5(const trw_csr_graph* graph, const int64_t* targets, int64_t n_walks,
                                              int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                                              int64_t* window_target, int64_t* window_pos, void* stream) {
    if (!graph || !targets || !window_target || !window_pos || n_walks < 0) { set_error("trw_walk_csr_prepared_windows5: bad argument"); return TRW_ERR_ARG; }
    if (walk_length < 4) { set_error("trw_walk_csr_prepared_windows5: walk_length must be at least 4"); return TRW_ERR_ARG; }
    if ((((uintptr_t)window_target) & 7) || (((uintptr_t)window_pos) & 31)) { set_error("trw_walk_csr_prepared_windows5: outputs must be 32-byte aligned"); return TRW_ERR_ARG; }
    if (n_walks == 0) return TRW_OK;
    DeviceGuard guard(graph->g.device);
    if (!guard.ok) { set_error("trw_walk_csr_prepared_windows5: cudaSetDevice(%d) failed", graph->g.device); return TRW_ERR_DEVICE; }
    CsrWalkPlan plan;
    int rc = csr_walk_plan(&plan, graph->g, p, q, walk_length, seed);
    if (rc) return rc;
    if (plan.uniform || plan.fold || !plan.table || plan.a.records == nullptr || plan.a.row32 == nullptr) {
        set_error("trw_walk_csr_prepared_windows5: needs a kept graph with records and a law that takes plain rejection (1/p <= max(1, 1/q), q <= 1 or p > q)");
        return TRW_ERR_ARG;
    }
    plan.a.win_target = window_target;
    plan.a.win_pos = window_pos;
    // `out` is unused by this variant; a non-null aligned dummy keeps the launcher on its staged path
    return csr_walk_launch(plan, targets, n_walks, walk_id_offset, window_target, (int64_t)walk_length + 1, (cudaStream_t)stream);
}

extern "C" void trw_csr_graph_destroy(trw_csr_graph* graph) { delete graph; }

extern "C" int trw_csr_graph_info(const trw_csr_graph* graph, void* stream, int64_t* out, int n_out) {
    if (!graph || !out || n_out < 6) { set_error("trw_csr_graph_info: null argument or room for fewer than 6 values"); return TRW_ERR_ARG; }
    DeviceGuard guard(graph->g.device);
    if (!guard.ok) { set_error("trw_csr_graph_info: cudaSetDevice(%d) failed", graph->g.device); return TRW_ERR_DEVICE; }
    const CsrPrepared& pr = graph->g.prepared;
    cudaStream_t st = (cudaStream_t)stream;
    int flags[2] = {0, 0};  // table_failed, asymmetric
    if (pr.table_failed) {
        int rc = check_cuda(cudaMemcpyAsync(&flags[0], pr.table_failed, sizeof(int), cudaMemcpyDeviceToHost, st), "read table flag");
        if (rc) return rc;
    }
    if (pr.asymmetric) {
        int rc = check_cuda(cudaMemcpyAsync(&flags[1], pr.asymmetric, sizeof(int), cudaMemcpyDeviceToHost, st), "read symmetry flag");
        if (rc) return rc;
    }
    int rc = check_cuda(cudaStreamSynchronize(st), "trw_csr_graph_info");
    if (rc) return rc;
    out[0] = pr.table != nullptr;
    out[1] = pr.records != nullptr;
    out[2] = pr.filter.bits ? (int64_t)pr.filter.n_bits : 0;
    out[3] = pr.asymmetric != nullptr;                       // triangle Blooms were computed
    out[4] = pr.asymmetric ? (flags[1] == 0 ? 1 : 0) : -1;   // symmetric: 1 yes, 0 no, -1 not checked
    out[5] = flags[0];
    return TRW_OK;
}
