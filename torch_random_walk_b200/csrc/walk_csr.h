// Internal interface of the CSR walk (shared by trw_walk_csr and trw_walk_csr_host).
#pragma once
#include "trw_common.cuh"

namespace trw {

struct WalkArgs {
    const int64_t* row_ptr;
    const int64_t* col_idx;
    int64_t n_nodes, nnz;
    const int64_t* targets;
    int64_t n_walks, walk_id_offset;
    int walk_length;
    int store_mode;  // experiment: L2 policy of the output stores (see output_policy)
    uint2 key;
    int64_t* out;
    int64_t out_row_stride;
    const uint32_t* table;
    const int* table_failed;  // device flag raised by the build when a hub segment overflowed
    const uint32_t* row32;  // uint32 copy of row_ptr (nullptr: read the int64 row_ptr)
    uint64_t thr0, thr1, thr2;  // acceptance thresholds on a 32-bit uniform, scaled by 2^32
    // return-edge folding (see node2vec_walk_kernel): envelope M', excess 1/p - M', thresholds 1/M', (1/q)/M'
    uint64_t fthr1, fthr2;
    double fold_env, fold_excess;
    const unsigned long long* strict_counts;  // [descents in col_idx, descents at row boundaries]; equal = rows strictly increasing
};

struct CsrWalkPlan {
    WalkArgs a;  // graph side filled by csr_walk_prepare; shard side by csr_walk_launch
    int device;
    int min_ctas;
    bool uniform, table, speculate, stage, persist, fold;
};

int csr_walk_prepare(CsrWalkPlan* plan, const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                     double p, double q, int walk_length, int64_t seed, void* workspace, size_t workspace_bytes,
                     int device, cudaStream_t st);
int csr_walk_launch(const CsrWalkPlan& plan, const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                    int64_t* out, int64_t out_row_stride, cudaStream_t st);

}  // namespace trw
