// Content checksum of a CSR graph: what lets a kept (prepared) graph be reused safely.
//
// The reference's launcher is stateless (csrc/cuda/rw_cuda.cu:186-248): it looks at row_ptr and col_idx
// afresh on every call.  A binding that keeps the graph-side preparation between calls must therefore
// know that the arrays still hold what was prepared -- tensor identity and version counters do not see
// writes through raw pointers, `.data`, DLPack or another library.  One streaming pass over both arrays
// (HBM speed: 0.7 ms for the 4.3 GB of the c3 graph) gives a 64-bit position-sensitive checksum; equal
// sizes and equal checksums are taken as equal graphs (collision probability 2^-64 per comparison).
#include "trw_common.cuh"
#include "trw_options.h"

namespace trw {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

// sum over i of mix64(value[i] + golden * (i + salt)): order-independent to accumulate, position-sensitive in value
__global__ void __launch_bounds__(256) csr_checksum_kernel(IdxPtr row_ptr, int64_t n_row, IdxPtr col_idx, int64_t nnz,
                                                           unsigned long long* __restrict__ out) {
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    // two elements per 16-byte load where the alignment allows it
    const bool vec = col_idx.wide() && (((uintptr_t)col_idx.base) & 15) == 0;
    if (vec) {
        const int64_t n2 = nnz >> 1;
        const longlong2* c2 = reinterpret_cast<const longlong2*>(col_idx.base);
        for (int64_t i = gtid; i < n2; i += gsz) {
            longlong2 v;
            asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(c2 + i));
            acc += mix64((uint64_t)v.x + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * i + 1));
            acc += mix64((uint64_t)v.y + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * i + 2));
        }
        if (gtid == 0 && (nnz & 1)) acc += mix64((uint64_t)ldg64_stream(col_idx + (nnz - 1)) + 0x9E3779B97F4A7C15ull * (uint64_t)nnz);
    } else {
        for (int64_t i = gtid; i < nnz; i += gsz)
            acc += mix64((uint64_t)ldg64_stream(col_idx + i) + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    }
    for (int64_t i = gtid; i < n_row; i += gsz)
        acc += mix64((uint64_t)ldg64_stream(row_ptr + i) + 0xD6E8FEB86659FD93ull * (uint64_t)(i + 1));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(out, s);
    }
}

}  // namespace trw

using namespace trw;

extern "C" int trw_csr_checksum(const int64_t* row_ptr, const int64_t* col_idx, int64_t n_nodes, int64_t nnz,
                                uint64_t* out_device, int device, void* stream) {
    return trw_csr_checksum_typed(row_ptr, 8, col_idx, 8, n_nodes, nnz, out_device, device, stream);
}

extern "C" int trw_csr_checksum_typed(const void* row_ptr, int row_ptr_bytes, const void* col_idx, int col_idx_bytes,
                                      int64_t n_nodes, int64_t nnz, uint64_t* out_device, int device, void* stream) {
    if ((row_ptr_bytes != 4 && row_ptr_bytes != 8) || (col_idx_bytes != 4 && col_idx_bytes != 8)) {
        set_error("trw_csr_checksum: CSR elements must be 4 or 8 bytes wide");
        return TRW_ERR_ARG;
    }
    if (n_nodes < 0 || nnz < 0 || !out_device || (n_nodes > 0 && !row_ptr) || (nnz > 0 && !col_idx)) {
        set_error("trw_csr_checksum: bad argument");
        return TRW_ERR_ARG;
    }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    DeviceGuard guard(d);
    if (!guard.ok) { set_error("trw_csr_checksum: cudaSetDevice(%d) failed", d); return TRW_ERR_DEVICE; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(out_device, 0, sizeof(uint64_t), st), "checksum memset");
    if (rc) return rc;
    const int64_t n_row = row_ptr ? n_nodes + 1 : 0;
    const int64_t work = (nnz >> 1) + n_row + 1;
    const int64_t want = (work + 255) / 256;
    const unsigned grid = (unsigned)(want < (int64_t)sm_count(d) * 16 ? (want < 1 ? 1 : want) : (int64_t)sm_count(d) * 16);
    csr_checksum_kernel<<<grid, 256, 0, st>>>(IdxPtr(row_ptr, row_ptr_bytes), n_row, IdxPtr(col_idx, col_idx_bytes), nnz,
                                              (unsigned long long*)out_device);
    count_launch(1);
    return check_cuda(cudaGetLastError(), "csr_checksum launch");
}
