// tcgen05 / TMA / mbarrier PTX wrappers and the shortlist helpers shared by the stage-1 tensor-core kernels
// (scan_tc.cu: one CTA per SM; scan_tc2.cu: CTA pairs, cta_group::2).  sm_100a only.
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include "tt_common.cuh"

namespace tt {
namespace tc {

constexpr int TILE_ROWS = 128;           // MMA M
constexpr int CHUNK_COLS = 64;           // bf16 elements per 128-byte swizzle row
constexpr int CHUNK_BYTES = TILE_ROWS * 128;  // one [128 x 64] bf16 sub-tile
constexpr int EPI_THREADS = 128;
constexpr int THREADS = 192;
constexpr int SMEM_LIMIT = 232448;       // 227 KB

constexpr uint64_t POLICY_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t POLICY_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t POLICY_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on the phase parity.  A wait that lasts ~seconds can only be a protocol bug or a lost TMA: it records
// TT_STATUS_RING_TIMEOUT and gives up (status_report, tt_common.cuh) -- the kernel then runs to its end on garbage,
// every later wait of the launch gives up after its first slow-path check, and the host raises TTError.  Neither a
// hung device nor a trapped context.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spins & 0xfffu) == 0xfffu) {
            if (status_raised()) return;
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > status_timeout_cycles()) {
                status_report(TT_STATUS_RING_TIMEOUT);
                return;
            }
        }
    }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3ffffu) >> 4);  // start address
    d |= uint64_t(1) << 16;                      // leading byte offset (unused with swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
    d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                      // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128, N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(TILE_ROWS >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// named barrier 1 over the EPI epilogue threads of the CTA
template <int EPI = EPI_THREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory"); }
// barrier + OR-reduction of a predicate over the 128 epilogue threads
template <int EPI = EPI_THREADS>
__device__ __forceinline__ bool epi_bar_or(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, %2, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(uint32_t(pred)), "n"(EPI)
        : "memory");
    return r != 0;
}

// Cut one candidate list of n entries back to its kp best: rank by counting, each lane owns entries lane + 32u
// (u < NE), the list is read once per rank step as a shared-memory broadcast.  Whole warp.
template <int NE>
__device__ __forceinline__ void cut_one(uint64_t* L, int n, int kp, float* thresh, int lane) {
    uint64_t e[NE];
    int r[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        const int x = lane + 32 * u;
        e[u] = x < n ? L[x] : 0ull;
        r[u] = 0;
    }
    for (int m = 0; m < n; ++m) {
        const uint64_t x = L[m];
#pragma unroll
        for (int u = 0; u < NE; ++u) r[u] += x > e[u] ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NE; ++u) {
        if (lane + 32 * u < n && r[u] < kp) {
            L[r[u]] = e[u];
            if (r[u] == kp - 1) *thresh = entry_key(e[u]);
        }
    }
}

// Cut candidate lists back to their K' best, one warp per query, warps stride over the queries.
// all = false: only lists that are full (count >= cap);  all = true: every list longer than K'.
// Sets thresh[q] to the K'-th key kept.  Callers put a barrier of the epilogue threads on both sides.
template <int EPI = EPI_THREADS>
__device__ __forceinline__ void cut_lists(uint64_t* lists, int* cnt_s, float* thresh_s, int nq, int kp, int cap,
                                          int warp, int lane, bool all) {
    for (int j = warp; j < nq; j += EPI / 32) {
        const int raw = cnt_s[j];
        const int n = min(raw, cap);
        if (all ? n <= kp : raw < cap) continue;
        uint64_t* L = lists + size_t(j) * cap;
        if (n <= 64) cut_one<2>(L, n, kp, thresh_s + j, lane);
        else if (n <= 128) cut_one<4>(L, n, kp, thresh_s + j, lane);
        else cut_one<8>(L, n, kp, thresh_s + j, lane);
        if (lane == 0) cnt_s[j] = kp;
    }
}

// The rare path, kept out of line so that the compiler does not speculate its arithmetic into the filter loop.
// Returns true when the list was full (the candidate stays pending).
static __device__ __noinline__ bool push_candidate(uint64_t* list, int* cnt, float s, uint32_t row, int cap, int* full) {
    const int slot = atomicAdd(cnt, 1);
    if (slot >= cap - 1) *full = 1;
    if (slot < cap) {
        list[slot] = pack_entry(s, row);
        return false;
    }
    return true;
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int NQ, bool HILO, int NACC, int EPI = EPI_THREADS>
__device__ __forceinline__ void filter_and_push(const float (&acc)[NACC], float inv, bool row_ok, uint32_t row, int nq,
                                                uint64_t* lists, int* cnt_s, float* thresh_s, int kp, int cap,
                                                int epi_warp, int lane, int q_off = 0) {
    static_assert(NQ % 4 == 0 && NQ <= 64, "NQ");
    constexpr int NW = (NQ + 31) / 32;
    uint32_t hit[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) hit[w] = 0u;
    const uint32_t th_addr = smem_u32(thresh_s + q_off);
#pragma unroll
    for (int j4 = 0; j4 < NQ / 4; ++j4) {
        const float4 th = lds_f4(th_addr + j4 * 16);
        const float tv[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j4 * 4 + u;
            const float s = (HILO ? acc[j] + acc[NQ + j] : acc[j]) * inv;
            if (s > tv[u]) hit[j >> 5] |= 1u << (j & 31);
        }
    }
    if (!row_ok) {
#pragma unroll
        for (int w = 0; w < NW; ++w) hit[w] = 0u;
    }
    for (;;) {
        int full = 0;
        bool any = false;
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= hit[w] != 0u;
        if (any) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                if (hit[j >> 5] & (1u << (j & 31))) {
                    const float s = (HILO ? acc[j] + acc[NQ + j] : acc[j]) * inv;
                    bool keep = false;
                    if (s > thresh_s[q_off + j])  // re-check: the threshold may have risen since the bit was set
                        keep = push_candidate(lists + size_t(q_off + j) * cap, cnt_s + q_off + j, s, row, cap, &full);
                    if (!keep) hit[j >> 5] &= ~(1u << (j & 31));
                }
            }
        }
        if (!epi_bar_or<EPI>(full != 0)) break;  // no list filled up: the tile is done (one barrier per tile)
        cut_lists<EPI>(lists, cnt_s, thresh_s, nq, kp, cap, epi_warp, lane, false);
        epi_bar_sync<EPI>();
    }
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Remote arrive with the default (.release.cta) semantics: a cluster-scope release would put a GPU-wide MEMBAR on
// the producer's / epilogue's critical path.  No generic-proxy data is published through these barriers (TMA bytes
// are tracked by complete_tx, TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Publishing DATA to the peer CTA through a barrier (the dynamic tile scheduler of the cluster kernels: one tile id per
// work item, microseconds apart -- here the cluster-scope release / acquire pair is affordable and required).
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// mbar_wait with acquire at cluster scope (pairs with mbar_arrive_cluster_release)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if ((spins & 0xfffu) == 0xfffu) {
            if (status_raised()) return;
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > status_timeout_cycles()) {
                status_report(TT_STATUS_RING_TIMEOUT);
                return;
            }
        }
    }
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA (cta_group::2)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                 uint32_t bar_cluster_addr, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster_addr), "l"(policy)
        : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(256 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(uint16_t(3))
        : "memory");
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, dim] bf16 row-major seen as (64 cols, rows, dim/64 chunks); box = (64, box_rows, box_chunks)
inline int make_map(CUtensorMap* m, const void* base, int64_t rows, int dim, int64_t stride_elems, int box_rows,
                    int box_chunks) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return TT_ERR_CUDA;
    }
    cuuint64_t dims[3] = {cuuint64_t(CHUNK_COLS), cuuint64_t(rows), cuuint64_t(dim / CHUNK_COLS)};
    cuuint64_t strides[2] = {cuuint64_t(stride_elems) * 2, cuuint64_t(CHUNK_COLS) * 2};
    cuuint32_t box[3] = {cuuint32_t(CHUNK_COLS), cuuint32_t(box_rows), cuuint32_t(box_chunks)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for base=%p rows=%lld dim=%d stride=%lld", int(r), base,
                  (long long)rows, dim, (long long)stride_elems);
        return TT_ERR_CUDA;
    }
    return TT_OK;
}

}  // namespace tc
}  // namespace tt
