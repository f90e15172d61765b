// Multi-GPU helpers behind the C ABI: CSR replication and walk gathering over NCCL.
//
// The reference has no distributed code.  Walks are independent, so a multi-GPU job needs two
// collectives only: ONE broadcast of the CSR at set-up (every rank then walks its shard of the start nodes with global
// walk ids and no traffic), and optionally a gather of the walk shards afterwards.  The Python package does both
// through torch.distributed (dist.py); these two entry points give a binder of include/trw_b200.h that has no Python
// the same thing over its own ncclComm_t.
//
// libtrw_b200.so does not link NCCL: the symbols are resolved at run time, first from what the process has already
// loaded (a host application that uses NCCL), then from libnccl.so.2 on the loader path.  Without NCCL the two
// functions return TRW_ERR_DEVICE with a message and nothing else in the library is affected.
#include <dlfcn.h>

#include <cstdlib>

#include <mutex>

#include "trw_common.cuh"

namespace trw {

typedef int (*nccl_broadcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_error_fn)(int);
typedef int (*nccl_group_fn)(void);
constexpr int kNcclUint8 = 1;  // ncclUint8 of nccl.h (stable across NCCL 2.x)

struct NcclApi {
    nccl_broadcast_fn broadcast = nullptr;
    nccl_error_fn error_string = nullptr;
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    bool tried = false;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (api.tried) return api;
    api.tried = true;
    // The functions must come from the library that made the caller's communicator: prefer what the process already
    // has (global symbols, then a loaded libnccl.so.2 whatever its path -- PyTorch loads its bundled copy privately);
    // only then load one, privately, so that a later import with its own NCCL is not disturbed.  TRW_NCCL_LIBRARY names
    // a specific file.
    void* handle = RTLD_DEFAULT;
    const char* named = getenv("TRW_NCCL_LIBRARY");
    if (named && *named) {
        handle = dlopen(named, RTLD_NOW | RTLD_LOCAL);
        if (!handle) return api;
    } else if (!dlsym(RTLD_DEFAULT, "ncclBroadcast")) {
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!handle) return api;
    }
    api.broadcast = (nccl_broadcast_fn)dlsym(handle, "ncclBroadcast");
    api.error_string = (nccl_error_fn)dlsym(handle, "ncclGetErrorString");
    api.group_start = (nccl_group_fn)dlsym(handle, "ncclGroupStart");
    api.group_end = (nccl_group_fn)dlsym(handle, "ncclGroupEnd");
    return api;
}

static int nccl_check(const NcclApi& api, int rc, const char* what) {
    if (rc == 0) return TRW_OK;
    set_error("%s: NCCL error %d (%s)", what, rc, api.error_string ? api.error_string(rc) : "?");
    return TRW_ERR_CUDA;
}

// NCCL counts are size_t elements; bytes are sent as ncclUint8 in pieces below 2^31 to stay clear of any 32-bit count
static int broadcast_bytes(const NcclApi& api, void* comm, int root, void* buf, size_t bytes, cudaStream_t st, const char* what) {
    const size_t piece = (size_t)1 << 30;
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t n = bytes - off < piece ? bytes - off : piece;
        const int rc = nccl_check(api, api.broadcast((char*)buf + off, (char*)buf + off, n, kNcclUint8, root, comm, st), what);
        if (rc) return rc;
    }
    return TRW_OK;
}

}  // namespace trw

using namespace trw;

extern "C" int trw_nccl_available(void) { return nccl_api().broadcast != nullptr ? 1 : 0; }

extern "C" int trw_replicate_csr(void* nccl_comm, int root, void* row_ptr, int row_ptr_bytes, int64_t n_nodes, void* col_idx,
                                 int col_idx_bytes, int64_t nnz, void* stream) {
    if (!nccl_comm || n_nodes < 0 || nnz < 0 || !row_ptr || (nnz > 0 && !col_idx) || (row_ptr_bytes != 4 && row_ptr_bytes != 8) ||
        (col_idx_bytes != 4 && col_idx_bytes != 8)) {
        set_error("trw_replicate_csr: bad argument");
        return TRW_ERR_ARG;
    }
    const NcclApi& api = nccl_api();
    if (!api.broadcast) { set_error("trw_replicate_csr: NCCL is not loaded in this process and libnccl.so.2 was not found"); return TRW_ERR_DEVICE; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = broadcast_bytes(api, nccl_comm, root, row_ptr, (size_t)(n_nodes + 1) * (size_t)row_ptr_bytes, st, "trw_replicate_csr(row_ptr)");
    if (rc) return rc;
    return broadcast_bytes(api, nccl_comm, root, col_idx, (size_t)nnz * (size_t)col_idx_bytes, st, "trw_replicate_csr(col_idx)");
}

extern "C" int trw_gather_walks(void* nccl_comm, int rank, int world, const int64_t* local, int64_t row_len, int64_t* out,
                                const int64_t* rows_per_rank, void* stream) {
    if (!nccl_comm || world < 1 || rank < 0 || rank >= world || row_len < 0 || !rows_per_rank || !out) {
        set_error("trw_gather_walks: bad argument");
        return TRW_ERR_ARG;
    }
    const NcclApi& api = nccl_api();
    if (!api.broadcast) { set_error("trw_gather_walks: NCCL is not loaded in this process and libnccl.so.2 was not found"); return TRW_ERR_DEVICE; }
    cudaStream_t st = (cudaStream_t)stream;
    // this rank's shard goes into its place first; every rank's slice of `out` is then broadcast from its owner
    int64_t first = 0;
    for (int r = 0; r < rank; ++r) first += rows_per_rank[r];
    if (rows_per_rank[rank] > 0) {
        if (!local) { set_error("trw_gather_walks: null local shard"); return TRW_ERR_ARG; }
        int rc = check_cuda(cudaMemcpyAsync(out + first * row_len, local, (size_t)rows_per_rank[rank] * row_len * 8, cudaMemcpyDeviceToDevice, st),
                            "trw_gather_walks(local copy)");
        if (rc) return rc;
    }
    int64_t row0 = 0;
    for (int r = 0; r < world; ++r) {
        const size_t bytes = (size_t)rows_per_rank[r] * (size_t)row_len * 8;
        if (bytes) {
            int rc = broadcast_bytes(api, nccl_comm, r, out + row0 * row_len, bytes, st, "trw_gather_walks");
            if (rc) return rc;
        }
        row0 += rows_per_rank[r];
    }
    return TRW_OK;
}
