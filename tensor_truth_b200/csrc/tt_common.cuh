// Shared device/host helpers for libtt_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tt_b200.h"

namespace tt {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char* fmt, ...);
int sm_count(int device);
int current_device();

#define TT_CHECK_ARG(cond, ...)        \
    do {                               \
        if (!(cond)) {                 \
            tt::set_error(__VA_ARGS__); \
            return TT_ERR_INVALID;     \
        }                              \
    } while (0)

#define TT_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            tt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return TT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#ifdef __CUDACC__
// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute when `dependent` (and not switched off: TT_NO_PDL)
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool dependent,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define TT_LAUNCH_OK(what)                                                                     \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            tt::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));            \
            return TT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------ ordering key packing
// One 64-bit word per (key, id): descending u64 order == key descending, id ascending.
//   hi 32 bits: order-preserving image of the fp32 key (NaN -> lowest), lo 32 bits: ~id.
// 0 is "empty" and sorts below every real entry.
__host__ __device__ __forceinline__ uint32_t okey_of(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    if (f != f) return 1u;  // NaN: just above "empty"
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return u < 2u ? 2u : u;
}
__host__ __device__ __forceinline__ float key_of_okey(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(u);
#else
    memcpy(&f, &u, 4);
#endif
    return f;
}
__host__ __device__ __forceinline__ uint64_t pack_entry(float key, uint32_t id) {
    return (uint64_t(okey_of(key)) << 32) | uint64_t(~id);
}
__host__ __device__ __forceinline__ float entry_key(uint64_t e) { return key_of_okey(uint32_t(e >> 32)); }
__host__ __device__ __forceinline__ uint32_t entry_id(uint64_t e) { return ~uint32_t(e); }

#ifdef __CUDACC__
// ------------------------------------------------------------------ device-side status (bounded waits never trap)
// A wait that runs out of time (an mbarrier of a TMA/MMA ring, a peer's exchange flag) records a TT_STATUS_* code
// and lets the kernel run to its end with whatever it has: the CUDA context -- and with it every other index and
// every peer rank's mapping of this GPU -- stays alive, and the host turns the code into a TTError
// (tt_status_configure / tt_status_read in the header).  The words are per translation unit (no relocatable
// device code in this build) and per device, like every __device__ variable; api.cu fans configure/read out.
struct StatusCfg {
    unsigned* mapped;          // optional host-mapped (pinned) word the host can poll without a device read
    long long timeout_cycles;  // bound on one wait
};
static __device__ unsigned g_status_word = 0u;
static __device__ StatusCfg g_status_cfg = {nullptr, 8000000000ll};

__device__ __forceinline__ void status_report(unsigned code) {
    atomicOr(&g_status_word, code);
    unsigned* m = g_status_cfg.mapped;
    if (m) {
        atomicOr_system(m, code);
        __threadfence_system();
    }
}
__device__ __forceinline__ bool status_raised() { return *reinterpret_cast<volatile unsigned*>(&g_status_word) != 0u; }
__device__ __forceinline__ long long status_timeout_cycles() { return g_status_cfg.timeout_cycles; }

static inline int status_configure_tu(unsigned* mapped, long long cycles) {
    StatusCfg c{mapped, cycles};
    return cudaMemcpyToSymbol(g_status_cfg, &c, sizeof(c)) == cudaSuccess ? 0 : -1;
}
static inline int status_read_tu(unsigned* out, bool clear) {
    unsigned v = 0u, z = 0u;
    if (cudaMemcpyFromSymbol(&v, g_status_word, sizeof(v)) != cudaSuccess) return -1;
    if (clear && v && cudaMemcpyToSymbol(g_status_word, &z, sizeof(z)) != cudaSuccess) return -1;
    *out |= v;
    return 0;
}

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// The three / four kernels of a query form a strict chain (prepare -> scan -> stage 2 [-> merge]).  Launched with the
// programmatic-stream-serialization attribute, a kernel may become resident while its predecessor still runs; it then
// blocks in pdl_wait() -- which returns once the predecessor grid has completed and its writes are visible -- right
// before it first touches the predecessor's output.  What that buys: the dependent's launch latency, and for the scan
// everything that does not need the prepared queries (barrier init, TMEM allocation), disappear from the latency chain.
// Both instructions are no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// every translation unit whose kernels can wait defines its pair of hooks with this (api.cu calls them all)
#define TT_DEFINE_STATUS_HOOKS(tu)                                                                        \
    int tu##_status_configure(unsigned* mapped, long long cycles) { return status_configure_tu(mapped, cycles); } \
    int tu##_status_read(unsigned* out, bool clear) { return status_read_tu(out, clear); }

// ------------------------------------------------------------------ warp helpers
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, uint32_t(v), m);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, uint32_t(v >> 32), m);
    return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
    uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
    return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        uint64_t o = shfl_xor_u64(v, m);
        v = o < v ? o : v;
    }
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ------------------------------------------------------------------ warp-distributed top-K' list
// K' = 32*E packed entries, E per lane in registers; `worst` (warp-uniform) is the smallest one.
// insert() takes a warp-uniform candidate.  Rows not kept are all <= the final `worst`.
template <int E>
struct WarpList {
    uint64_t e[E];
    uint64_t worst;

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < E; ++i) e[i] = 0ull;
        worst = 0ull;
    }
    __device__ __forceinline__ void refresh() {
        uint64_t m = e[0];
#pragma unroll
        for (int i = 1; i < E; ++i) m = e[i] < m ? e[i] : m;
        worst = warp_min_u64(m);
    }
    // x is warp-uniform
    __device__ __forceinline__ void insert(uint64_t x) {
        if (x <= worst) return;
        // replace exactly one copy of `worst` (the lowest lane / slot holding it)
        bool mine = false;
        int slot = -1;
#pragma unroll
        for (int i = E - 1; i >= 0; --i)
            if (e[i] == worst) { mine = true; slot = i; }
        unsigned who = __ballot_sync(0xffffffffu, mine);
        int lane = threadIdx.x & 31;
        if (lane == __ffs(who) - 1) {
#pragma unroll
            for (int i = 0; i < E; ++i)
                if (i == slot) e[i] = x;
        }
        refresh();
    }
};

// ------------------------------------------------------------------ block bitonic sort (descending) in smem
// n must be a power of two; all threads of the block participate.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* s, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = s[i], b = s[ixj];
                    bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        s[i] = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// exclusive prefix sum of one int per thread over the whole block (blockDim.x <= 1024, multiple of 32).
// warp_sums: smem int[33].  Returns the exclusive prefix; *total receives the block total.
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();  // protect warp_sums from the previous call
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = lane < nw ? warp_sums[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += o;
        }
        if (lane < nw) warp_sums[lane] = winc - ws;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[w] + inc - v;
}

// 8 bf16 packed in a uint4 -> 8 floats (exact)
__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x << 16);
    f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16);
    f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16);
    f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16);
    f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
#endif  // __CUDACC__

}  // namespace tt
