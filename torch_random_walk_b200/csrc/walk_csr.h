// Internal interface of the CSR walk (shared by trw_walk_csr and trw_walk_csr_host).
#pragma once
#include "member_table.cuh"
#include "trw_common.cuh"

namespace trw {

struct WalkArgs {
    IdxPtr row_ptr;
    IdxPtr col_idx;
    int64_t n_nodes, nnz;
    const int64_t* targets;
    int64_t n_walks, walk_id_offset;
    int64_t id_block, id_stride;  // block-cyclic shards: local walk i has global id offset + (i / block) * stride + i % block (block 0: offset + i)
    int walk_length;
    int store_mode;  // experiment: L2 policy of the output stores (see output_policy)
    uint2 key;
    int64_t* out;
    int64_t out_row_stride;
    const uint32_t* table;
    const int* table_failed;  // device flag raised by the build when a hub segment overflowed
    const uint32_t* row32;  // uint32 copy of row_ptr (nullptr: read the int64 row_ptr)
    const uint4* records;   // edge records (member_table.cuh), or nullptr
    EdgeFilter filter;      // L2-resident edge filter in front of the table (bits == nullptr: none)
    const int* asymmetric;  // device flag of the triangle-Bloom pass (member_table.cuh); nullptr: symmetry unknown
    uint64_t thr0, thr1, thr2;  // acceptance thresholds on a 32-bit uniform, scaled by 2^32
    // return-edge folding (see node2vec_walk_kernel): envelope M', excess 1/p - M', thresholds 1/M', (1/q)/M'
    uint64_t fthr1, fthr2;
    double fold_env, fold_excess;
    int64_t* win_target = nullptr;  // A/B of the fused walk -> window pipeline (WindowOut): skip-gram targets ...
    int64_t* win_pos = nullptr;     // ... and positive windows of width 5 instead of the walk rows
    double w_back, w_common, w_far;  // the node2vec weights themselves, 1/p, 1, 1/q (exact-CDF A/B kernel)
    int mix;  // 1: two-sided mixture sampling (q > 1, p <= q); fold_env = 1/q, fold_excess = 1/p - 1/q
    const unsigned long long* strict_counts;  // [descents in col_idx, descents at row boundaries]; equal = rows strictly increasing
};

// A CSR graph as the walk kernels see it: the caller's arrays plus what csr_graph_prepare derived
// from them into the workspace (all optional: a null member selects the slower generic path).
struct CsrGraph {
    IdxPtr row_ptr;
    IdxPtr col_idx;
    int64_t n_nodes = 0, nnz = 0;
    int device = 0;
    CsrPrepared prepared;
};

struct CsrWalkPlan {
    WalkArgs a;  // graph side and (p, q, seed) filled by csr_walk_plan; shard side by csr_walk_launch
    int device;
    int min_ctas;
    bool uniform, table, speculate, stage, persist, fold;
};

int csr_graph_prepare(CsrGraph* g, IdxPtr row_ptr, IdxPtr col_idx, int64_t n_nodes, int64_t nnz,
                      bool uniform, bool want_table, bool want_strict, bool want_records, void* workspace,
                      size_t workspace_bytes, int device, cudaStream_t st, int64_t bloom_cap = 0);
int csr_graph_of_handle(const trw_csr_graph* h, CsrGraph* out);  // the graph behind a C-ABI handle (walk_csr.cu)
int csr_walk_plan(CsrWalkPlan* plan, const CsrGraph& g, double p, double q, int walk_length, int64_t seed);
void csr_one_shot_needs(double p, double q, int64_t nnz, int64_t n_walks, int walk_length, bool* uniform,
                        bool* want_table, bool* want_strict, bool* want_records);
int csr_walk_launch(const CsrWalkPlan& plan, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                    int64_t* out, int64_t out_row_stride, cudaStream_t st, int64_t id_block = 0, int64_t id_stride = 0);

// Global id of local walk i (the Philox counter): contiguous shards add an offset, block-cyclic ones
// (dist.py) own every stride-th block of `block` consecutive ids.
__device__ __forceinline__ uint64_t global_walk_id(const WalkArgs& a, int64_t i) {
    if (a.id_block <= 0) return (uint64_t)(a.walk_id_offset + i);
    const int64_t blk = i / a.id_block;
    return (uint64_t)(a.walk_id_offset + blk * a.id_stride + (i - blk * a.id_block));
}

}  // namespace trw
