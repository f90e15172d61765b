/*
 * trw_b200.h -- C ABI of libtrw_b200.so: the B200 (sm_100a) random-walk sampler and window
 * generator behind torch_rw's `torch_rw_native` module.
 *
 * Drop-in boundary.  Each entry point below replaces one function of the reference's pybind11
 * module (/root/reference/csrc/rw_init.cpp:133-141); the reference interface it stands in for
 * is cited on each declaration.  Signatures carry plain pointers and sizes only: device
 * pointers to contiguous row-major int64 arrays, a CUDA device ordinal and a cudaStream_t
 * passed as void*.  Nothing here allocates device memory, synchronises the stream or touches
 * host copies of the data (the *_host entry point excepted, which says so): the binding
 * (torch_random_walk_b200/native.py, or the cgo/JNI/ctypes stub of INTEGRATION.md) owns
 * allocation, exactly as the reference's launchers call torch::empty before their kernels.
 *
 * Every function returns TRW_OK (0) or a negative status; trw_last_error() then holds a
 * message for the calling thread.  There is no CPU fallback: without a usable sm_100 device
 * every launch returns TRW_ERR_DEVICE.
 *
 * Randomness: counter-based Philox4x32-10 keyed by (seed, entry-point tag) and indexed by
 * (global walk id, step, trial) or (output element).  Output therefore depends only on the
 * arguments -- not on grid shape, stream, rerun, or on how walks are sharded over GPUs, as
 * long as each shard passes its global `walk_id_offset`.
 */
#ifndef TRW_B200_H
#define TRW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRW_ABI_VERSION 3

enum {
    TRW_OK = 0,
    TRW_ERR_ARG = -1,     /* bad argument (null pointer, negative size, ...)        */
    TRW_ERR_DEVICE = -2,  /* no CUDA device / not an sm_100 part / cudaSetDevice     */
    TRW_ERR_CUDA = -3,    /* a CUDA runtime call or kernel launch failed             */
    TRW_ERR_WORKSPACE = -4 /* workspace pointer missing or too small                 */
};

/* ABI version of the loaded library (== TRW_ABI_VERSION it was built with). */
int trw_abi_version(void);

/* Message describing the last non-zero status returned to this thread ("" if none). */
const char* trw_last_error(void);

/* Number of kernels this library has launched since load / since the last reset
 * (bench.py's `gpu_launches`). */
int64_t trw_launch_count(void);
void trw_reset_launch_count(void);

/* TRW_OK when `device` is a compute-capability-10.x GPU the kernels can run on. */
int trw_device_check(int device);

/* ---------------------------------------------------------------------------------------
 * CSR walks.   Replaces walk() -> walk_gpu(): csrc/rw_init.cpp:11-25,
 * csrc/cuda/rw_cuda.cu:186-248 (kernels :59-98 uniform, :100-184 node2vec).
 *
 *   row_ptr[n_nodes+1], col_idx[nnz], targets[n_walks]  -> out[n_walks, walk_length+1]
 *   out_row_stride: elements between consecutive rows of `out` (>= walk_length+1).
 *   p == 1.0 && q == 1.0 selects the first-order kernel, as rw_cuda.cu:226 does.
 *   walk_id_offset: global index of this call's first walk (0 for an unsharded call).
 *   A node without out-edges keeps the walk on that node (rw_cuda.cu:25-30).
 *
 * The node2vec path tests "x in adj(t)" against a per-call hashed copy of the adjacency that
 * it builds in `workspace` (trw_walk_csr_workspace_bytes bytes, 256-byte aligned device
 * memory, contents undefined before and after).  With workspace == NULL it falls back to the
 * reference's linear scan of adj(t).
 * ------------------------------------------------------------------------------------- */
size_t trw_walk_csr_workspace_bytes(int64_t n_nodes, int64_t nnz, double p, double q);
/* The exact size this call shape uses: shorter than the bound above when the walk is too short
 * to pay for the per-call edge records (about 3 gathered lines per CSR entry). */
size_t trw_walk_csr_workspace_bytes_for(int64_t n_nodes, int64_t nnz, double p, double q,
                                        int64_t n_walks, int walk_length);

int trw_walk_csr(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                 const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                 double p, double q, int walk_length, int64_t seed,
                 int64_t* out, int64_t out_row_stride,
                 void* workspace, size_t workspace_bytes,
                 int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Prepared graphs.  Everything trw_walk_csr derives from the graph alone (uint32 row index,
 * hashed membership table, duplicate-edge check, 16-byte edge records) does not depend on p, q,
 * the seed or the start nodes.  A caller that walks the same graph repeatedly (one rw.walk per
 * epoch) prepares it once and passes the handle; the walk call is then the walk kernel alone.
 * The reference has no counterpart (its walk_gpu is stateless, csrc/cuda/rw_cuda.cu:186-248);
 * results are bit-identical to trw_walk_csr on the same arguments.
 *
 *   workspace: trw_csr_graph_workspace_bytes(n_nodes, nnz) bytes of 256-byte aligned device
 *   memory.  It, row_ptr and col_idx belong to the caller and must stay alive and unchanged
 *   until trw_csr_graph_destroy.  Preparation is enqueued on `stream`; walks on another stream
 *   must be ordered after it by the caller.
 * ------------------------------------------------------------------------------------- */
typedef struct trw_csr_graph trw_csr_graph;
size_t trw_csr_graph_workspace_bytes(int64_t n_nodes, int64_t nnz);
int trw_csr_graph_prepare(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                          void* workspace, size_t workspace_bytes, int device, void* stream,
                          trw_csr_graph** out_graph);
/* As above with the triangle-Bloom pass chosen per call: bloom_cap < 0 the library default (option
 * edge_bloom_cap, 256), 0 none (trw_csr_graph_add_blooms can add them later), > 0 that cap. */
int trw_csr_graph_prepare_ex(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                             void* workspace, size_t workspace_bytes, int device, void* stream,
                             int64_t bloom_cap, trw_csr_graph** out_graph);
int trw_walk_csr_prepared(const trw_csr_graph* graph, const int64_t* targets, int64_t n_walks,
                          int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                          int64_t* out, int64_t out_row_stride, void* stream);
/* The same walk with the CSR arrays at another address: for callers that have verified, with
 * trw_csr_checksum, that (row_ptr, col_idx) hold exactly what the graph was prepared from.  This is how the
 * Python binding keeps a prepared graph across rw.walk calls without holding on to the caller's tensors.
 * Global walk ids (the Philox counters): local walk i is walk_id_offset + i, or, for a block-cyclic shard
 * (walk_id_block > 0), walk_id_offset + (i / walk_id_block) * walk_id_stride + i % walk_id_block -- rank r of
 * W ranks with blocks of B walks passes (r*B, B, W*B). */
int trw_walk_csr_prepared_at(const trw_csr_graph* graph, const void* row_ptr, const void* col_idx,
                             const int64_t* targets, int64_t n_walks,
                             int64_t walk_id_offset, int64_t walk_id_block, int64_t walk_id_stride,
                             double p, double q, int walk_length, int64_t seed,
                             int64_t* out, int64_t out_row_stride, void* stream);
/* Adds the triangle Blooms (a 32-bit Bloom of the common neighbours of every edge's endpoints,
 * kept in the edge records) to a graph prepared without them; one pass, quadratic in `cap`, the longest
 * "shorter row" it works out exactly (<= 0: the library default).  No-op when already present.  It changes what the
 * handle holds: do not call it while another thread launches walks through the same handle, and order walks on other
 * streams after `stream`. */
int trw_csr_graph_add_blooms(trw_csr_graph* graph, const void* row_ptr, const void* col_idx, int64_t cap, void* stream);
/* 64-bit position-sensitive checksum of (row_ptr[n_nodes+1], col_idx[nnz]) into *out_device (device memory):
 * one streaming pass on `stream`.  Equal sizes and checksums identify the graph a kept preparation belongs
 * to; the reference, being stateless (csrc/cuda/rw_cuda.cu:186-248), needs no such thing. */
int trw_csr_checksum(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                     uint64_t* out_device, int device, void* stream);
/* The same sum over arrays in HOST memory, on n_threads host threads (<= 0: as many as the host path uses); simd = 0
 * keeps to the scalar loop, otherwise AVX-512 where the CPU has it (same value).  A host caller can compare it with
 * trw_csr_checksum of a device replica -- what trw_walk_csr_host does on every call to its kept replica. */
int trw_csr_checksum_host(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                          int n_threads, int simd, uint64_t* out);
void trw_csr_graph_destroy(trw_csr_graph* graph);

/* ---------------------------------------------------------------------------------------
 * CSR arrays with 32-bit elements.  The reference's accessors are int64-only
 * (csrc/cuda/rw_cuda.cu:206-209: an int32 tensor raises); these entries take row_ptr and col_idx
 * with 4- or 8-byte elements each (signed), which halves the graph in HBM and on PCIe.  Start
 * nodes and walks stay int64, and the walks are bit-identical to the int64 call on the same
 * values.  A graph prepared from typed arrays remembers their widths: trw_walk_csr_prepared_at,
 * trw_csr_graph_add_blooms and trw_csr_graph_view take pointers to arrays of those same widths.
 * ------------------------------------------------------------------------------------- */
int trw_walk_csr_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                       int64_t n_nodes, int64_t nnz,
                       const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                       double p, double q, int walk_length, int64_t seed,
                       int64_t* out, int64_t out_row_stride,
                       void* workspace, size_t workspace_bytes, int device, void* stream);
int trw_csr_graph_prepare_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                                int64_t n_nodes, int64_t nnz, void* workspace, size_t workspace_bytes,
                                int device, void* stream, int64_t bloom_cap, trw_csr_graph** out_graph);
int trw_csr_checksum_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                           int64_t n_nodes, int64_t nnz, uint64_t* out_device, int device, void* stream);
/* What a prepared graph holds, after waiting for `stream` (the stream it was prepared on):
 * out[0] membership table, out[1] edge records, out[2] bits of the L2-resident edge filter (0: none),
 * out[3] triangle Blooms computed, out[4] graph symmetric (1 yes, 0 no, -1 not checked),
 * out[5] table build overflowed (the walk then scans).  n_out >= 6. */
int trw_csr_graph_info(const trw_csr_graph* graph, void* stream, int64_t* out, int n_out);

/* Same computation with every buffer in HOST memory (pinned or pageable): stages the graph and
 * the start nodes to the device, walks in chunks and streams finished chunks back while the
 * next ones run.  Allocates its own device memory (kept between calls, grow-only, until
 * trw_release_cached_buffers) and returns after `out` is complete.  This is the end-to-end
 * path a caller holding CPU tensors uses.  A caller that comes back with the same host arrays finds
 * the device replica of the graph kept: the walk starts on it at once while the copy engine (re-reading pinned
 * arrays) and the host threads verify, by checksum, that the arrays still hold what was uploaded (a mismatch uploads
 * afresh and walks again). */
int trw_walk_csr_host(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                      const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                      double p, double q, int walk_length, int64_t seed,
                      int64_t* out, int device);

/* The download half on its own, for a graph that is already on the device and prepared (each rank of a multi-GPU
 * job runs this on its shard): start nodes from host memory, walks into host memory, chunks walked and copied back
 * in a pipeline, uint32 on the wire when the ids fit and the host has the threads.  Global walk ids as in
 * trw_walk_csr_prepared_at.  Returns after `out` is complete. */
typedef struct trw_csr_graph_view {
    const trw_csr_graph* graph; /* from trw_csr_graph_prepare                                              */
    const void* row_ptr;        /* device arrays holding the prepared content, in the element widths the   */
    const void* col_idx;        /* graph was prepared with (NULL: those of the handle)                     */
    void* ready_stream;         /* stream the preparation was enqueued on (NULL: known to be complete)     */
} trw_csr_graph_view;
int trw_walk_csr_to_host(const trw_csr_graph_view* view, const int64_t* targets, int64_t n_walks,
                         int64_t walk_id_offset, int64_t walk_id_block, int64_t walk_id_stride,
                         double p, double q, int walk_length, int64_t seed, int64_t* out);

/* The replica trw_walk_csr_host keeps on `device`: out[0] one is held, out[1] its preparation level (-1 none, 0 one
 * call's needs, 1 the full kept preparation, 2 with triangle Blooms), out[2] validated hits so far, out[3] how the last
 * call got its graph (0 no call yet, 1 kept replica validated, 2 fresh upload, 3 kept replica found changed and
 * uploaded afresh).  n_out >= 4; with n_out >= 6 also out[4], out[5] = the bytes the last host-path call on the device
 * (trw_walk_csr_host or trw_walk_csr_to_host) moved over PCIe host->device (start nodes, graph upload, the copy engine's
 * share of the content check) and device->host (walks, 4 or 8 bytes per entry). */
int trw_host_replica_info(int device, int64_t* out, int n_out);

/* Frees the device buffers, streams and events trw_walk_csr_host keeps between calls. */
void trw_release_cached_buffers(void);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU helpers over NCCL (the reference has no distributed code).
 * Walks are independent: one broadcast replicates the CSR at set-up, every rank then walks its shard of
 * the start nodes with global walk ids (walk_id_offset / walk_id_block / walk_id_stride above) and no
 * traffic, and the shards may be gathered afterwards.  `nccl_comm` is the caller's ncclComm_t; NCCL is
 * resolved at run time (from the process, else libnccl.so.2), so the library itself does not link it.
 *
 * trw_replicate_csr: broadcasts row_ptr[n_nodes+1] and col_idx[nnz] (4- or 8-byte elements) from `root`
 *   into the same-sized device buffers of every other rank, on `stream`.
 * trw_gather_walks: out[sum(rows_per_rank), row_len] on every rank = the ranks' contiguous shards in rank
 *   order (`local` holds rows_per_rank[rank] rows; rows_per_rank is a host array of `world` entries).
 * trw_nccl_available: 1 when the NCCL entry points were found.
 * ------------------------------------------------------------------------------------- */
int trw_nccl_available(void);
int trw_replicate_csr(void* nccl_comm, int root, void* row_ptr, int row_ptr_bytes, int64_t n_nodes,
                      void* col_idx, int col_idx_bytes, int64_t nnz, void* stream);
int trw_gather_walks(void* nccl_comm, int rank, int world, const int64_t* local, int64_t row_len,
                     int64_t* out, const int64_t* rows_per_rank, void* stream);

/* ---------------------------------------------------------------------------------------
 * Edge-list walks.   Replaces walk_edge_list() -> walk_edge_list_gpu():
 * csrc/rw_init.cpp:27-45, csrc/cuda/rw_cuda_edge_list.cu:243-308 (kernels :42-96, :126-240).
 *
 *   edge_list[n_edges,2] sorted by head, node_edge_index[n_index_rows,2] = inclusive
 *   [first,last] edge row per node or [-1,-1]  ->  out[n_walks, walk_length+1].
 *   Dead end -> padding_idx, then the start node (restart != 0) or padding for ever.
 *   The second-order path keeps the reference's acceptance rule verbatim (half-open
 *   neighbour scan, fall-through after a rejected return).
 * ------------------------------------------------------------------------------------- */
int trw_walk_edge_list(const int64_t* edge_list, int64_t n_edges,
                       const int64_t* node_edge_index, int64_t n_index_rows,
                       const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                       double p, double q, int walk_length, int64_t seed,
                       int64_t padding_idx, int restart,
                       int64_t* out, int64_t out_row_stride, int device, void* stream);

/* The same walk with a workspace: the second-order path then answers "x is a neighbour of t" from
 * the hashed membership table of the CSR walk (built per call over a CSR view of the edge list)
 * instead of the reference's O(deg) scan (rw_cuda_edge_list.cu:98-123), with identical results.
 * workspace: trw_walk_edge_list_workspace_bytes() bytes of 256-byte aligned device memory (0 for
 * p == q == 1, which needs none); NULL selects the scan. */
size_t trw_walk_edge_list_workspace_bytes(int64_t n_edges, int64_t n_index_rows, double p, double q);
int trw_walk_edge_list_ws(const int64_t* edge_list, int64_t n_edges,
                          const int64_t* node_edge_index, int64_t n_index_rows,
                          const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                          double p, double q, int walk_length, int64_t seed,
                          int64_t padding_idx, int restart,
                          int64_t* out, int64_t out_row_stride,
                          void* workspace, size_t workspace_bytes, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Knowledge-graph triple walks.   Replaces walk_triples() -> triples::walk_triples_gpu():
 * csrc/rw_init.cpp:47-75, csrc/cuda/rw_cuda_triples.cu:103-169 (kernel :49-96).
 *
 *   triples[n_triples,3] sorted by head, relation_tail_index[n_index_rows,2]
 *   -> out[n_walks, 2*walk_length+1] = head, rel, tail, rel, tail, ...
 *   `restart` is accepted and ignored, as in the reference.  Unlike the reference's CUDA
 *   launcher (which discards a non-zero seed, rw_cuda_triples.cu:143-148) the seed is honoured.
 * ------------------------------------------------------------------------------------- */
int trw_walk_triples(const int64_t* triples, int64_t n_triples,
                     const int64_t* relation_tail_index, int64_t n_index_rows,
                     const int64_t* targets, int64_t n_walks, int64_t walk_id_offset,
                     int walk_length, int64_t padding_idx, int restart, int64_t seed,
                     int64_t* out, int64_t out_row_stride, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Window generation (one gather kernel per call).  walks[n_walks, walk_cols] contiguous.
 * Positives/targets are bit-exact with the reference; negatives are Philox draws with the
 * reference's support (uniform node id / uniform row of `triples`).
 *
 * trw_windows            replaces to_windows()      csrc/rw_init.cpp:77-88,
 *                                                   csrc/cuda/windows_cuda.cu:67-119 (:7-65)
 *   -> target[K], pos[K,W-1], neg[K,W-1],  K = n_walks*(walk_cols-W+1)
 * trw_windows_cbow       replaces to_windows_cbow() csrc/rw_init.cpp:90-101,
 *                                                   csrc/cuda/windows_cuda.cu:187-239 (:122-185)
 *   -> pos_nodes[K], neg_nodes[K] (!= pos_nodes[k], up to 101 redraws), windows[K,W-1]
 * trw_windows_triples    replaces to_windows_triples()  csrc/rw_init.cpp:103-116,
 *                                                   csrc/cuda/windows_cuda.cu:375-435 (:241-373)
 *   -> target[K,3], pos[K,2W,3], neg[K,2W,3],  K = n_walks*((walk_cols-1)/2)
 * trw_windows_triples_cbow replaces to_windows_triples_cbow() csrc/rw_init.cpp:118-131,
 *                                                   csrc/cuda/windows_cuda.cu:584-644 (:438-582)
 *   -> pos_triples[K,3], neg_triples[K,3], pos_windows[K,2W,3]
 * ------------------------------------------------------------------------------------- */
int trw_windows(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                int64_t num_nodes, int64_t seed,
                int64_t* target, int64_t* pos, int64_t* neg, int device, void* stream);

/* Negatives of the node windows from a caller-supplied distribution instead of the uniform one (SURVEY.md section 8 f3,
 * optional; no counterpart in the reference, whose negatives are rand() % num_nodes, csrc/cuda/windows_cuda.cu:57-62):
 * trw_alias_table_build (HOST code, O(n)) turns weights[n] >= 0 into an alias table for P(v) ~ weights[v]^power --
 * word2vec's unigram^0.75 with weights = degrees, power = 0.75 -- one 8-byte cell per node, threshold | alias << 32;
 * copy it to the device and pass it to trw_windows_alias / trw_windows_cbow_alias, which are trw_windows /
 * trw_windows_cbow with every negative drawn from it (n == num_nodes < 2^32; the CBOW form still redraws a negative
 * equal to its positive node, up to 101 times).  Targets and positives are unchanged. */
int trw_alias_table_build(const double* weights, int64_t n, double power, uint64_t* table_out);
int trw_windows_alias(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                      int64_t num_nodes, int64_t seed, const uint64_t* alias_table,
                      int64_t* target, int64_t* pos, int64_t* neg, int device, void* stream);
int trw_windows_cbow_alias(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                           int64_t num_nodes, int64_t seed, const uint64_t* alias_table,
                           int64_t* pos_nodes, int64_t* neg_nodes, int64_t* windows, int device, void* stream);

int trw_windows_cbow(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                     int64_t num_nodes, int64_t seed,
                     int64_t* pos_nodes, int64_t* neg_nodes, int64_t* windows,
                     int device, void* stream);

int trw_windows_triples(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                        int64_t num_nodes, int64_t padding_idx,
                        const int64_t* triples, int64_t n_triples, int64_t seed,
                        int64_t* target, int64_t* pos, int64_t* neg, int device, void* stream);

int trw_windows_triples_cbow(const int64_t* walks, int64_t n_walks, int64_t walk_cols,
                             int window_size, int64_t num_nodes, int64_t padding_idx,
                             const int64_t* triples, int64_t n_triples, int64_t seed,
                             int64_t* pos_triples, int64_t* neg_triples, int64_t* pos_windows,
                             int device, void* stream);

/* The triple-window calls with a workspace (trw_windows_triples_workspace_bytes(n_triples) bytes of 256-byte aligned
 * device memory, contents undefined before and after): the negative rows are then gathered from a 16-byte uint32
 * copy of `triples` made inside the call (one 128-bit load per row instead of three 64-bit ones) whenever every id
 * fits 32 bits; identical outputs. */
size_t trw_windows_triples_workspace_bytes(int64_t n_triples);
int trw_windows_triples_ws(const int64_t* walks, int64_t n_walks, int64_t walk_cols, int window_size,
                           int64_t num_nodes, int64_t padding_idx,
                           const int64_t* triples, int64_t n_triples, int64_t seed,
                           int64_t* target, int64_t* pos, int64_t* neg,
                           void* workspace, size_t workspace_bytes, int device, void* stream);
int trw_windows_triples_cbow_ws(const int64_t* walks, int64_t n_walks, int64_t walk_cols,
                                int window_size, int64_t num_nodes, int64_t padding_idx,
                                const int64_t* triples, int64_t n_triples, int64_t seed,
                                int64_t* pos_triples, int64_t* neg_triples, int64_t* pos_windows,
                                void* workspace, size_t workspace_bytes, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Measurement helpers (bench.py / profiles only; not part of the reference's surface).
 * trw_calib_gather: every thread performs `loads_per_thread` dependent random loads of
 * `bytes_per_load` (8, 32, 64 or 128; the wide ones read adjacent sectors) from table[0, table_elems) -- the attainable random-sector rate
 * the walk kernels are compared against.  sink[>=1] receives a checksum.
 * ------------------------------------------------------------------------------------- */
int trw_calib_gather(const int64_t* table, int64_t table_elems, int64_t n_threads,
                     int loads_per_thread, int bytes_per_load, int64_t seed, int64_t* sink,
                     int device, void* stream);

/* A/B of the fused walk -> window pipeline (measurement only): node2vec walks
 * on a kept graph (plain-rejection laws) whose skip-gram windows of width 5 are written straight to
 * window_target[n_walks*(L-3)] and window_pos[n_walks*(L-3), 4] -- bit-identical to trw_windows on the walks of
 * trw_walk_csr_prepared -- without the walks ever being stored. */
int trw_walk_csr_prepared_windows5(const trw_csr_graph* graph, const int64_t* targets, int64_t n_walks,
                                   int64_t walk_id_offset, double p, double q, int walk_length, int64_t seed,
                                   int64_t* window_target, int64_t* window_pos, void* stream);

/* With option "time_kernels" = 1 the CSR walk brackets its table build (memset + build kernel)
 * and its walk kernel with CUDA events on the launch stream; this waits for the last such call
 * and returns both durations in milliseconds (0 when that phase did not run). */
int trw_last_kernel_ms(float* build_ms, float* walk_ms);

/* out[0..5] = SM count, L2 bytes, max persisting-L2 bytes, max access-policy window bytes,
 * current L2 fetch granularity, max opt-in shared memory per block. */
int trw_device_info(int device, int64_t* out, int n_out);

/* Tuning knobs for experiments ("name" -> integer); returns TRW_ERR_ARG for unknown names.
 * Defaults are the shipped configuration; bench.py records any override it applies.  A knob belongs to the thread
 * that sets it: other threads keep the defaults, and concurrent callers do not see each other's settings. */
int trw_set_option(const char* name, int64_t value);
int64_t trw_get_option(const char* name);
/* Puts every option of the calling thread back to the shipped default. */
void trw_reset_options(void);

#ifdef __cplusplus
}
#endif
#endif /* TRW_B200_H */
