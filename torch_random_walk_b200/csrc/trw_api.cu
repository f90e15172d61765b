// Error plumbing, options, device checks and the calibration gather of libtrw_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "trw_common.cuh"
#include "trw_options.h"

namespace trw {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return TRW_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return TRW_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count(int device) {
    static int cached[64];
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cached[device] = v;
    }
    return cached[device];
}

// The knobs belong to the calling thread: a thread that never sets one runs on the shipped defaults whatever other
// threads experiment with, and concurrent callers cannot change each other's kernels mid-call.
Options& options() {
    thread_local Options o;
    return o;
}

// Event pairs of option time_kernels: per device (an event belongs to the device it was created on), guarded by a
// mutex; slot 0 = graph preparation, slot 1 = walk kernel of the last CSR walk on that device.
struct TimingSlot {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used = false;
};
static TimingSlot g_timing[64][2];
static std::mutex g_timing_mu;
static int g_timing_last_device = 0;

static TimingSlot* timing_slot(int slot) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) { cudaGetLastError(); return nullptr; }
    g_timing_last_device = d;
    return &g_timing[d][slot];
}

void timing_begin(int slot, cudaStream_t st) {
    if (!options().time_kernels) return;
    std::lock_guard<std::mutex> lock(g_timing_mu);
    TimingSlot* t = timing_slot(slot);
    if (!t) return;
    for (int k = 0; k < 2; ++k)
        if (!t->ev[k] && cudaEventCreate(&t->ev[k]) != cudaSuccess) { cudaGetLastError(); t->ev[k] = nullptr; return; }
    if (cudaEventRecord(t->ev[0], st) != cudaSuccess) cudaGetLastError();  // a failed measurement must not fail the walk
}

void timing_end(int slot, cudaStream_t st) {
    if (!options().time_kernels) return;
    std::lock_guard<std::mutex> lock(g_timing_mu);
    TimingSlot* t = timing_slot(slot);
    if (!t || !t->ev[0] || !t->ev[1]) return;
    if (cudaEventRecord(t->ev[1], st) != cudaSuccess) { cudaGetLastError(); return; }
    t->used = true;
}

int resolve_device(int device) {
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) {
            set_error("no current CUDA device (this library has no CPU fallback)");
            return -1;
        }
    }
    static int checked[64];
    if (device < 64 && checked[device] == 1) return device;
    int major = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) {
        set_error("CUDA device %d unavailable: %s (this library has no CPU fallback)", device, cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
    }
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libtrw_b200 carries sm_100a code only", device, major);
        return -1;
    }
    if (device < 64) checked[device] = 1;
    return device;
}

// ------------------------------------------------------------------ calibration gather
// Each thread chases `loads` dependent pseudo-random locations: the address of load k+1 mixes
// the value returned by load k, like a walk step.  BYTES = 8 reads one int64; 32/64/128 read
// 1/2/4 adjacent sectors.  MODE selects the load flavour for 8-byte loads (which cache operators
// and L2 policies change what a miss costs in DRAM traffic).
template <int MODE>
__device__ __forceinline__ uint64_t calib_load64(const int64_t* p, uint64_t pol) {
    uint64_t v;
    if (MODE == 1) asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.cv.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    else if (MODE == 5) asm volatile("ld.global.lu.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.cs.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <int BYTES, int MODE>
__global__ void __launch_bounds__(256) calib_gather_kernel(const int64_t* __restrict__ table, uint64_t n_units,
                                                           int64_t n_threads, int loads, uint2 key,
                                                           int64_t* __restrict__ sink, int aux_mb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_threads) return;
    const uint64_t pol = make_policy_evict_first();
    const uint64_t pol_keep = make_policy_evict_last();
    uint64_t acc = 0;
    uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0, 0), key);
    uint64_t state = ((uint64_t)r.x << 32) | r.y;
    for (int k = 0; k < loads; ++k) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        if (aux_mb > 0) {  // experiment: one dependent L2-resident lookup per step, like the walk's row index
            const uint64_t a_unit = __umul64hi(state ^ (state >> 31), (uint64_t)aux_mb << 18);
            state += (uint64_t)ldg32_keep(reinterpret_cast<const uint32_t*>(table) + a_unit, pol_keep) & 0xFFu;
        }
        uint64_t unit = __umul64hi(state ^ (state >> 29), n_units);
        if (BYTES == 8) {
            uint64_t v = calib_load64<MODE>(table + unit, pol);
            acc += v;
            state += v;
        } else {
            // BYTES/32 adjacent sectors of one aligned BYTES-wide granule
            uint64_t mix = 0;
#pragma unroll
            for (int j = 0; j < BYTES / 32; ++j) {
                Sector64 s = ldg_sector(table + unit * (BYTES / 8) + 4 * j);
                mix ^= s.a ^ s.b ^ s.c ^ s.d;
            }
            acc += mix;
            state += mix;
        }
    }
    if (acc == 0x9E3779B97F4A7C15ull) sink[0] = (int64_t)acc;  // keeps the loads alive
}

}  // namespace trw

using namespace trw;

extern "C" {

int trw_abi_version(void) { return TRW_ABI_VERSION; }
const char* trw_last_error(void) { return g_err; }
int64_t trw_launch_count(void) { return g_launches.load(); }
void trw_reset_launch_count(void) { g_launches.store(0); }

int trw_device_check(int device) {
    int d = resolve_device(device);
    return d < 0 ? TRW_ERR_DEVICE : TRW_OK;
}

int trw_calib_gather(const int64_t* table, int64_t table_elems, int64_t n_threads, int loads_per_thread,
                     int bytes_per_load, int64_t seed, int64_t* sink, int device, void* stream) {
    if (!table || !sink || table_elems < 4 || n_threads < 0 || loads_per_thread < 0 ||
        (bytes_per_load != 8 && bytes_per_load != 32 && bytes_per_load != 64 && bytes_per_load != 128)) {
        set_error("trw_calib_gather: bad argument");
        return TRW_ERR_ARG;
    }
    if (bytes_per_load >= 32 && ((uintptr_t)table & 127)) {
        set_error("trw_calib_gather: table must be 128-byte aligned for sector loads");
        return TRW_ERR_ARG;
    }
    int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    DeviceGuard g(d);
    if (n_threads == 0) return TRW_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((n_threads + 255) / 256);
    uint2 key = philox_key(seed, 0x43414C49u);
    const int64_t units8 = table_elems;
    const int mode = (int)options().calib_mode;
    const int aux_mb = (int)options().calib_aux_mb;
#define TRW_CALIB8(M) calib_gather_kernel<8, M><<<grid, 256, 0, st>>>(table, (uint64_t)units8, n_threads, loads_per_thread, key, sink, aux_mb)
    if (bytes_per_load == 8) {
        switch (mode) {
            case 1: TRW_CALIB8(1); break;
            case 2: TRW_CALIB8(2); break;
            case 3: TRW_CALIB8(3); break;
            case 4: TRW_CALIB8(4); break;
            case 5: TRW_CALIB8(5); break;
            case 6: TRW_CALIB8(6); break;
            default: TRW_CALIB8(0); break;
        }
    } else if (bytes_per_load == 32)
        calib_gather_kernel<32, 0><<<grid, 256, 0, st>>>(table, (uint64_t)(table_elems / 4), n_threads, loads_per_thread, key, sink, aux_mb);
    else if (bytes_per_load == 64)
        calib_gather_kernel<64, 0><<<grid, 256, 0, st>>>(table, (uint64_t)(table_elems / 8), n_threads, loads_per_thread, key, sink, aux_mb);
    else
        calib_gather_kernel<128, 0><<<grid, 256, 0, st>>>(table, (uint64_t)(table_elems / 16), n_threads, loads_per_thread, key, sink, aux_mb);
#undef TRW_CALIB8
    count_launch(1);
    return check_cuda(cudaGetLastError(), "calib_gather launch");
}

int trw_last_kernel_ms(float* build_ms, float* walk_ms) {
    float* dst[2] = {build_ms, walk_ms};
    cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool used[2] = {false, false};
    int d = 0;
    {
        std::lock_guard<std::mutex> lock(g_timing_mu);
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) { cudaGetLastError(); d = g_timing_last_device; }
        for (int slot = 0; slot < 2; ++slot) {
            used[slot] = g_timing[d][slot].used;
            ev[slot][0] = g_timing[d][slot].ev[0];
            ev[slot][1] = g_timing[d][slot].ev[1];
        }
        g_timing[d][0].used = false;  // a walk on a kept graph leaves the preparation slot unused
    }
    for (int slot = 0; slot < 2; ++slot) {
        if (!dst[slot]) continue;
        *dst[slot] = 0.0f;
        if (!used[slot]) continue;
        int rc = check_cuda(cudaEventSynchronize(ev[slot][1]), "trw_last_kernel_ms");
        if (rc) return rc;
        rc = check_cuda(cudaEventElapsedTime(dst[slot], ev[slot][0], ev[slot][1]), "trw_last_kernel_ms");
        if (rc) return rc;
    }
    return TRW_OK;
}

int trw_device_info(int device, int64_t* out, int n_out) {
    if (!out || n_out < 6) { set_error("trw_device_info: need room for 6 values"); return TRW_ERR_ARG; }
    const int d = resolve_device(device);
    if (d < 0) return TRW_ERR_DEVICE;
    DeviceGuard g(d);
    int v = 0;
    size_t lim = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); out[0] = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, d); out[1] = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, d); out[2] = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, d); out[3] = v;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); out[4] = (int64_t)lim;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, d); out[5] = v;
    cudaGetLastError();
    return TRW_OK;
}

int trw_set_option(const char* name, int64_t value) {
    if (!name) return TRW_ERR_ARG;
    Options& o = options();
    if (!strcmp(name, "l2_fetch_granularity")) {  // device-wide hint: 32, 64 or 128 bytes fetched per L2 miss
        return check_cuda(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value), "cudaLimitMaxL2FetchGranularity");
    }
#define TRW_OPT(field) if (!strcmp(name, #field)) { o.field = value; return TRW_OK; }
    TRW_OPTION_LIST
#undef TRW_OPT
    set_error("trw_set_option: unknown option '%s'", name);
    return TRW_ERR_ARG;
}

void trw_reset_options(void) { options() = Options{}; }

int64_t trw_get_option(const char* name) {
    if (!name) return -1;
    Options& o = options();
#define TRW_OPT(field) if (!strcmp(name, #field)) return o.field;
    TRW_OPTION_LIST
#undef TRW_OPT
    return -1;
}

}  // extern "C"
