// Tuning knobs of the calling thread (trw_set_option / trw_get_option; thread_local).  Defaults = shipped configuration.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trw {

struct Options {
    int64_t stage_output = 1;     // 1: CSR walk output staged in shared memory and written as whole lines by the warp (LineStager); 0: plain 8-byte stores
    int64_t n2v_table = 1;        // 1: hashed adjacency membership (needs workspace); 0: linear scan of adj(t)
    int64_t n2v_speculate = -1;   // fetch row_ptr[x] before the membership answer is known: 1 yes, 0 no, -1 by (p,q)
    int64_t n2v_fold = 1;         // 1: fold the return edge out of the rejection envelope when 1/p > max(1, 1/q)
    int64_t n2v_slots = 8;        // A/B: 16 selects whole-line store pieces for the plain-rejection kernel with records (default 64-byte pieces)
    int64_t n2v_mix = 1;          // 1: two-sided mixture sampling when q > 1 and p <= q (duplicate-free rows; see node2vec_walk_kernel)
    int64_t n2v_warp = 0;         // A/B: 1 runs node2vec walks on a kept graph with the warp-per-walk exact-CDF kernel (node2vec_warp_walk_kernel)
    int64_t n2v_min_ctas = -1;    // __launch_bounds__ min CTAs/SM of the node2vec kernel (4, 5 or 6; -1: 5 with edge records, else 4)
    int64_t row32 = 1;            // 1: re-encode row_ptr as uint32 offsets (needs workspace; the edge records depend on it)
    int64_t el_table = 1;         // 1: edge-list node2vec walks test membership through the hashed table (needs workspace); 0: the reference's scan
    int64_t records = -1;         // 16-byte edge records (neighbour id + its row span; the walk then needs no row-index loads): 1 always, 0 never,
                                  // -1 kept graphs always, one-shot calls when the walk is long enough to repay one pass over col_idx (csr_one_shot_needs)
    int64_t edge_filter_mb = 32;  // L2-resident hub-pair filter in front of the membership table (member_table.cuh; filled by the triangle-Bloom pass
                                  // for edges between rows longer than the Bloom cap): size cap in MB, 0 = none
                                  // (default: measured -8 % on the c3 walk alone, but nothing on top of the triangle Blooms, and 2.9 ms per build)
    int64_t edge_bloom_cap = 256; // kept graphs: triangle Blooms in the edge records (member_table.cuh) for pairs whose shorter row has at most
                                  // this many entries (the pass is quadratic in it); 0 = none
    int64_t build_mode = 2;       // table build: 2 assembled in shared memory (tiles + hub segments); 0 global CAS (A/B baseline)
    int64_t persist_row_ptr = 0;  // 1: L2 access-policy window (persisting) over row_ptr during walk kernels
    int64_t persist_l2_mb = 64;   // persisting-L2 carve-out requested when persist_row_ptr is on
    int64_t host_chunk_walks = 1 << 20;  // walks per pipelined chunk in trw_walk_csr_host
    int64_t host_compress = 1;    // wire format of the host path when ids fit 32 bits: 1 col_idx goes up as uint32 (>= 12 host threads) and walks come back
                                  // as uint32 for as many chunks as the host threads keep up with (host_packed_share), 2 every other chunk packed, 0 plain int64 copies
    int64_t host_threads = 0;     // host threads of the wire compression (0: the machine's, divided by LOCAL_WORLD_SIZE)
    int64_t host_up_chunk = 1 << 25;  // col_idx entries per compressed upload chunk
    int64_t host_packed_share = -1; // of every 8 download chunks, how many travel as uint32 (-1: decided chunk by chunk by how the host threads keep up)
    int64_t host_check_dma = 4;   // of the kept replica's content check, up to how many eighths the copy engine may re-read from pinned host arrays
                                  // (summed on the device; it takes pieces from the front of the list, the host threads from the back; 0: host threads only)
    int64_t host_sum_piece = 1 << 22;  // entries per piece of the content check's work list (32 MB; tests shrink it to exercise the list on small graphs)
    int64_t host_keep_graph = 1;  // 1: trw_walk_csr_host keeps the device replica of the graph between calls (same host arrays, content
                                  // checked by checksum on every call); needs host_cache_buffers
    int64_t host_cache_buffers = 1;  // 1: trw_walk_csr_host keeps its device buffers between calls
    int64_t win_table16 = 1;      // 1: the triple window kernels gather their negative rows from a 16-byte uint32 copy of `triples` when the caller
                                  // passes a workspace (trw_windows_triples_ws) and every id fits; 0: always the int64 table (A/B)
    int64_t win_direct_pos = 1;   // 1: triple window kernels write the positive windows straight from the walk tile in 16-byte pieces (window_size <= 10);
                                  // 0: through the per-warp stage like the other rows
    int64_t win_bulk = 0;         // 1: the warps' stages of the triple window kernels leave as bulk stores (cp.async.bulk, SASS UBLKCP) instead of 16-byte
                                  // stores.  A/B only: measured equal for to_windows_triples (2.58 vs 2.60 ms) and slower for the CBOW form, whose
                                  // occupancy the second stage costs (profiles/r02_summary.md) -- the kernels are bound by index arithmetic and
                                  // the L2 gathers of the negative rows, not by their stores
    int64_t store_mode = 0;       // output-store L2 policy experiment: 0 evict_first, 1 normal, 2 evict_last, 3 no stores
    int64_t smem_carveout_kb = 0; // > 0: preferred shared-memory carve-out (KB per SM) of the CSR walk kernels
    int64_t calib_aux_mb = 0;     // calibration gather: > 0 adds one dependent lookup per step into the first N MiB of the table (L2-resident)
    int64_t calib_mode = 0;       // load flavour of the 8-byte calibration gather (see calib_load64)
    int64_t time_kernels = 0;     // 1: bracket the CSR table build and walk kernel with CUDA events (trw_last_kernel_ms)
};

#define TRW_OPTION_LIST                                                                      \
    TRW_OPT(stage_output) TRW_OPT(n2v_table) TRW_OPT(n2v_speculate) TRW_OPT(persist_row_ptr) \
    TRW_OPT(persist_l2_mb) TRW_OPT(host_chunk_walks) TRW_OPT(time_kernels) TRW_OPT(n2v_min_ctas) TRW_OPT(row32)       \
    TRW_OPT(build_mode) TRW_OPT(calib_mode) TRW_OPT(n2v_fold) TRW_OPT(host_cache_buffers) TRW_OPT(host_compress) TRW_OPT(host_threads) TRW_OPT(host_up_chunk) TRW_OPT(store_mode) TRW_OPT(records) TRW_OPT(el_table) TRW_OPT(n2v_mix) TRW_OPT(n2v_slots) TRW_OPT(calib_aux_mb) TRW_OPT(smem_carveout_kb) TRW_OPT(edge_filter_mb) TRW_OPT(edge_bloom_cap) TRW_OPT(host_keep_graph) TRW_OPT(host_packed_share) TRW_OPT(win_bulk) TRW_OPT(win_direct_pos) TRW_OPT(n2v_warp) TRW_OPT(win_table16) TRW_OPT(host_check_dma) TRW_OPT(host_sum_piece)

Options& options();
void count_launch(int n);
// Validates `device` (or the current device when < 0) as an sm_100 part; returns its ordinal or -1.
int resolve_device(int device);

// Event pairs around the last CSR table build (slot 0) and walk kernel (slot 1) when time_kernels is on.
void timing_begin(int slot, cudaStream_t st);
void timing_end(int slot, cudaStream_t st);

}  // namespace trw
