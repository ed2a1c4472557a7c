// Stage 1 for every batch whose queries travel as bf16 hi halves (more than 32 concurrent queries, up to tens of
// thousands: the tensor-bound regime).  A GEMM-shaped kernel -- corpus AND query tiles both streamed through the TMA
// ring, 256 x N output tiles per CTA pair (tcgen05.mma cta_group::2, M = 256; N = 256 with the accumulators
// double-buffered in all 512 TMEM columns, N = 128 / 64 when the whole batch fits one narrower block) -- whose
// epilogue never writes the score matrix: every accumulator element is compared with its query's running threshold
// tau[q] and only the survivors are appended to that query's candidate buffer in HBM.
//
// The thresholds come from the data itself.  The corpus is visited in PHASES over a pseudo-random permutation
// of its 256-row super-tiles: phase 0 is a small sample scanned with tau = -inf (everything is kept), after each
// phase `gemm_cut_kernel` radix-selects the K' best of every query's buffer and sets tau[q] to the K'-th approximate
// score seen so far.  Each phase visits `growth` times the rows seen before it, so it appends about growth x K'
// candidates per query and the whole scan writes O(K' log(n_rows)) entries per query instead of n_rows.
// A row that is dropped anywhere has a(r) <= tau at that moment <= the final K'-th score, which is the
// `out_thresh` contract of tt_scan_topk_bf16 with a single list per query -- stage 2 (rescore.cu) and the
// certificate are unchanged.  A query whose buffer overflows reports thresh = +inf (certificate fails, the caller's
// repair ladder answers it).
//
// Queries travel as bf16 hi halves (one MMA column per query).
//
// Replaces the vector-store query behind `index.as_retriever(similarity_top_k=k)` at
// /root/reference/src/tensortruth/rag_engine.py:639 for a large batch of concurrent queries (BASELINE config C4).
#include "tc_ptx.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(scan_gemm)
namespace tc3 {

using namespace tc;

// NQB = query columns per output tile (MMA N): 256 in the tensor-bound regime; 128 / 64 for batches that fit one
// narrower block -- those are HBM-bound, and a narrower query stage leaves more of the ring to corpus bytes in flight.
constexpr int NQB_MAX = 256;
constexpr int EPI3 = 256;                      // 8 epilogue warps
constexpr int THREADS3 = EPI3 + 64;

struct Params {
    const float* inv_norm;
    int64_t n_rows;
    int n_super;          // 256-row super-tiles in the corpus
    int n_chunks;         // dim / 64
    int stages;
    int n_qb;             // NQB-query blocks
    int i0, i1;           // this phase: permuted super-tile indices [i0, i1)
    uint32_t perm_mul;    // super = (i * perm_mul) % n_super, gcd(perm_mul, n_super) == 1
    int* sched;           // {next work item} counter, zero at launch: work items are claimed dynamically (clusters run at
                          // different speeds -- HBM channels, L2 slices -- and a static interleave leaves the fast ones idle)
    const float* tau;     // [>= n_qb * NQB] running thresholds (+inf in the padding columns)
    unsigned long long* buf;  // [n_q, cap] packed (approx key, local row) entries
    int* cnt;             // [>= n_qb * NQB] entries appended so far (may exceed cap: overflow)
    int cap;
    int direct_min;       // survivors in a 32 x 32 chunk from which they bypass the staging (DIRECT_MIN; tuning knob)
};

constexpr int STG_CAP = 64;  // staged survivors per epilogue warp (16 B each)
constexpr int DIRECT_MIN = 24;  // survivors in a 32 x 32 chunk from which they bypass the staging (one reservation per column)
constexpr int SCHED_SLOTS = 4;  // work-item ring between CTA 0's producer and every other role of the pair

// v[j] for a warp-uniform j: a select tree instead of dynamic register indexing
__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
    float a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[16 + i] : v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[8 + i] : a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[4 + i] : b[i];
    const float d0 = (j & 2) ? c[2] : c[0], d1 = (j & 2) ? c[3] : c[1];
    return (j & 1) ? d1 : d0;
}

// Append the n staged (entry, query) records of this warp to their queries' buffers in HBM.  All 32 lanes' atomics
// are in flight together (one round trip per 32 survivors instead of one each); lanes appending to the same query
// share one atomic.  Out of line: the hot loop only pays for it when a stage fills up.
static __device__ __noinline__ void flush_staged(const uint4* stg, int n, unsigned long long* buf, int* cnt, int cap) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += 32) {
        const bool active = base + lane < n;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (active) x = stg[base + lane];
        const int q = active ? int(x.z) : -1 - lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, q);
        const int leader = __ffs(peers) - 1;
        int slot0 = 0;
        if (active && lane == leader) slot0 = atomicAdd(cnt + q, __popc(peers));
        slot0 = __shfl_sync(0xffffffffu, slot0, leader);
        const int slot = slot0 + __popc(peers & ((1u << lane) - 1u));
        if (active && slot < cap) buf[size_t(q) * cap + slot] = (uint64_t(x.y) << 32) | x.x;
    }
    __syncwarp();
}

// Dynamic shared memory of each CTA (base rounded up to 1024 B):
//   [ ring: stages x (A 16 KB | B 16 KB) ][ tau: 2 x 256 f32 ][ staging: 8 warps x 64 x 16 B ]
//   [ barriers: full[stages] (CTA 0), empty[stages], tmem_full[2], tmem_empty[2] (CTA 0) ][ tmem base ]
// Ten warps: 0-7 epilogue (warps w and w+4 share the TMEM lane quarter w%4 and split the 256 query columns in
// halves), 8 TMA producer, 9 TMEM alloc + MMA issue (leader CTA).
// HILO (NQB = 64 only): up to 32 queries, each as TWO MMA columns -- column j is the bf16 hi half of query j, column
// 32 + j its lo half (CTA 0 of the pair stages the hi rows, CTA 1 the lo rows: map_q / map_q2) -- and the epilogue
// adds the two accumulators before it filters.  16 mantissa bits of the query reach the tensor cores, so the tight
// certificate bound of the hi+lo scan holds, at the bytes-in-flight of the streaming GEMM pipeline.
template <int NQB, int CH, bool HILO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS3, 1)
scan_gemm_kernel(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_q,
                 const __grid_constant__ CUtensorMap map_q2, const Params p) {
    static_assert(!HILO || NQB == 64, "the hi+lo variant is the 64-column pass");
    constexpr int NH = NQB / 2;                     // query rows (B operand) held by each CTA of the pair
    constexpr int STAGE_A = CH * CHUNK_BYTES;       // 128 corpus rows x CH 64-column chunks (256 B contiguous per row at CH = 2)
    constexpr int STAGE_B = CH * NH * 128;          // NH query rows x CH chunks
    constexpr int STAGE_BYTES = STAGE_A + STAGE_B;  // 32 KB at NQB = 256, CH = 1
    constexpr int TMEM_COLS = 2 * NQB;              // 128 / 256 / 512: powers of two
    static_assert(NQB == 64 || NQB == 128 || NQB == 256, "NQB");
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

    unsigned char* ring = smem;
    float* tau_s = reinterpret_cast<float*>(ring + size_t(p.stages) * STAGE_BYTES);
    uint4* stage_s = reinterpret_cast<uint4*>(tau_s + 2 * NQB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_s + (EPI3 / 32) * STG_CAP);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* tmem_full = bars + 2 * p.stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* sched_full = tmem_empty + 2;               // SCHED_SLOTS: a work item id has been published in this CTA's ring
    uint64_t* sched_empty = sched_full + SCHED_SLOTS;    // SCHED_SLOTS (used in CTA 0): every consumer of both CTAs has read it
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sched_empty + SCHED_SLOTS);
    volatile int* work_ring = reinterpret_cast<volatile int*>(tmem_base_s + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = int(blockIdx.x) >> 1, n_clusters = int(gridDim.x) >> 1;
    const int n_work = int(min(int64_t(p.i1 - p.i0) * p.n_qb, int64_t(0x7fffffff)));  // (super-tile, query block) pairs, query block fastest
    // Work items travel from CTA 0's producer (the only thread that claims them) to every other role of both CTAs
    // through a SCHED_SLOTS-deep ring: id written into both CTAs' rings, sched_full raised in both, sched_empty (CTA 0)
    // collects one arrival per consumer.  -1 ends the stream.
    auto sched_take = [&](int it) -> int {  // consumers: the it-th work item (blocking)
        const int slot = it & (SCHED_SLOTS - 1);
        const uint32_t par = uint32_t(it / SCHED_SLOTS) & 1u;
        if (rank == 0) mbar_wait(smem_u32(sched_full + slot), par);
        else mbar_wait_cluster(smem_u32(sched_full + slot), par);
        return work_ring[slot];
    };
    auto sched_release = [&](int it) {      // one thread per consumer, after its last read of the ring slot
        mbar_arrive_cluster(mapa(smem_u32(sched_empty + (it & (SCHED_SLOTS - 1))), 0));
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 2);   // one arrive per CTA's producer; the bytes of both land here (CTA 0)
            mbar_init(smem_u32(empty_bar + s), 1);  // multicast tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full + a), 1);
            mbar_init(smem_u32(tmem_empty + a), 2 * (EPI3 / 32));  // one lane per epilogue warp of both CTAs
        }
        for (int r = 0; r < SCHED_SLOTS; ++r) {
            mbar_init(smem_u32(sched_full + r), 1);
            mbar_init(smem_u32(sched_empty + r), 2 * (EPI3 / 32) + 2);  // epilogue warps of both CTAs + MMA thread + CTA 1's producer
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                     "r"(uint32_t(TMEM_COLS))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 8) {
        // ===================================================== TMA producer (both CTAs: own corpus rows + own query half)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            auto load_item = [&](int w) {
                const int qb = w % p.n_qb;
                const int super = int((uint64_t(p.i0 + w / p.n_qb) * p.perm_mul) % uint32_t(p.n_super));
                const int row0 = super * 256 + int(rank) * TILE_ROWS;
                const int qrow0 = HILO ? 0 : qb * NQB + int(rank) * NH;  // HILO: one block of <= 32 queries, hi rows / lo rows
                for (int c = 0; c < p.n_chunks; c += CH) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
                    const uint32_t fb = mapa(smem_u32(full_bar + stage), 0);
                    if (rank == 0) mbar_expect_tx(smem_u32(full_bar + stage), 2 * STAGE_BYTES);
                    else mbar_arrive_cluster(fb);
                    const uint32_t dst = smem_u32(ring + size_t(stage) * STAGE_BYTES);
                    tma_load_3d_pair(dst, &map_c, 0, row0, c, fb, POLICY_EVICT_NORMAL);  // re-read from L2 by every query block
                    tma_load_3d_pair(dst + STAGE_A, (HILO && rank == 1) ? &map_q2 : &map_q, 0, qrow0, c, fb, POLICY_EVICT_LAST);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            };
            if (rank == 0) {
                // the publisher: item `it + 1` is claimed and published BEFORE the loads of item `it` are issued, so every
                // consumer (the epilogue prefetches one item ahead) finds its next id waiting
                auto publish = [&](int it, int w) {
                    const int slot = it & (SCHED_SLOTS - 1);
                    mbar_wait(smem_u32(sched_empty + slot), (uint32_t(it / SCHED_SLOTS) & 1u) ^ 1u);
                    work_ring[slot] = w;
                    st_shared_cluster_u32(mapa(smem_u32(const_cast<int*>(work_ring) + slot), 1), uint32_t(w));
                    mbar_arrive(smem_u32(sched_full + slot));
                    mbar_arrive_cluster_release(mapa(smem_u32(sched_full + slot), 1));
                };
                int w = cluster_id < n_work ? cluster_id : -1;
                publish(0, w);
                for (int it = 0; w >= 0; ++it) {
                    int nx = p.sched ? n_clusters + atomicAdd(p.sched, 1) : w + n_clusters;
                    if (nx >= n_work) nx = -1;
                    publish(it + 1, nx);
                    load_item(w);
                    w = nx;
                }
            } else {
                for (int it = 0;; ++it) {
                    const int w = sched_take(it);
                    sched_release(it);
                    if (w < 0) break;
                    load_item(w);
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== MMA issuer: one thread of the leader CTA
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_m256(NQB);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0;; ++it) {
                const int w = sched_take(it);
                sched_release(it);
                if (w < 0) break;
                const int a = it & 1;
                mbar_wait(smem_u32(tmem_empty + a), (uint32_t(it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(a * NQB);
                for (int c = 0; c < p.n_chunks; c += CH) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(ring + size_t(stage) * STAGE_BYTES);
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
                        for (int k = 0; k < CHUNK_COLS / 16; ++k)
                            umma_bf16_pair(d_tmem, umma_desc_sw128(a_base + ch * CHUNK_BYTES + k * 32),
                                           umma_desc_sw128(a_base + STAGE_A + ch * NH * 128 + k * 32), idesc,
                                           uint32_t((c | ch | k) != 0));
                    }
                    umma_commit_pair(smem_u32(empty_bar + stage));  // frees this ring slot in both CTAs
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(smem_u32(tmem_full + a));  // both CTAs' accumulator halves are complete
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: thread = (corpus row, 128 of the 256 query columns)
        const int quarter = warp & 3, half = warp >> 2;
        const uint32_t te0 = mapa(smem_u32(tmem_empty), 0), te1 = mapa(smem_u32(tmem_empty + 1), 0);
        uint4* stg = stage_s + warp * STG_CAP;
        int n_stg = 0;  // warp-uniform
        const int64_t row_in_super = int64_t(rank) * TILE_ROWS + quarter * 32 + lane;
        // thresholds and inv_norm of a tile are fetched one tile ahead (their latency hides behind the current tile)
        float tau_n = INFINITY, inv_n = 1.f;
        int w = sched_take(0);
        if (w >= 0) {
            const int super = int((uint64_t(p.i0 + w / p.n_qb) * p.perm_mul) % uint32_t(p.n_super));
            const int64_t row = int64_t(super) * 256 + row_in_super;
            if (int(threadIdx.x) < NQB) tau_n = __ldcg(p.tau + size_t(w % p.n_qb) * NQB + threadIdx.x);
            if (p.inv_norm && row < p.n_rows) inv_n = __ldg(p.inv_norm + row);
        }
        for (int it = 0; w >= 0; ++it) {
            const int a = it & 1;
            const int qb = w % p.n_qb;
            const int super = int((uint64_t(p.i0 + w / p.n_qb) * p.perm_mul) % uint32_t(p.n_super));
            const int64_t row = int64_t(super) * 256 + row_in_super;
            const bool row_ok = row < p.n_rows;
            const float inv = inv_n;

            // this tile's thresholds (double-buffered by `a`: one barrier per tile keeps readers and writers apart)
            float* th = tau_s + a * NQB;
            if (int(threadIdx.x) < NQB) th[threadIdx.x] = tau_n;
            epi_bar_sync<EPI3>();
            const int w2 = sched_take(it + 1);  // published one item ahead of the loads: normally waiting already
            if (w2 >= 0) {
                const int super2 = int((uint64_t(p.i0 + w2 / p.n_qb) * p.perm_mul) % uint32_t(p.n_super));
                const int64_t row2 = int64_t(super2) * 256 + row_in_super;
                if (int(threadIdx.x) < NQB) tau_n = __ldcg(p.tau + size_t(w2 % p.n_qb) * NQB + threadIdx.x);
                inv_n = (p.inv_norm && row2 < p.n_rows) ? __ldg(p.inv_norm + row2) : 1.f;
            }

            mbar_wait(smem_u32(tmem_full + a), uint32_t(it >> 1) & 1u);
            tcgen05_fence_after();
            // HILO: this warp's 16 queries are columns [16 half, 16 half + 16) (hi) and the same + 32 (lo)
            const int col0 = HILO ? half * 16 : half * (NQB / 2);
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(a * NQB + col0);
            const uint32_t th_addr = smem_u32(th + col0);
#pragma unroll 1
            for (int c32 = 0; c32 < (HILO ? 32 : NQB / 2); c32 += 32) {
                float v[32];
                if (HILO) {
                    float lo[16];
                    tmem_ld_x16(taddr, v);
                    tmem_ld_x16(taddr + 32, lo);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] += lo[i];
                        v[16 + i] = -INFINITY;  // columns this warp does not own: never pass (their thresholds are some other query's)
                    }
                } else {
                    tmem_ld_x32(taddr + c32, v);
                    tmem_ld_wait();
                }
                // fast filter: does any of the 32 columns beat its threshold?  fma(v, inv, -tau) > 0 is implied by
                // the exact test (v * inv rounded) > tau used below, so nothing is missed.
                float m = -INFINITY;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 t4 = lds_f4(th_addr + (c32 + j4 * 4) * 4);
                    m = fmaxf(m, fmaf(v[j4 * 4 + 0], inv, -t4.x));
                    m = fmaxf(m, fmaf(v[j4 * 4 + 1], inv, -t4.y));
                    m = fmaxf(m, fmaf(v[j4 * 4 + 2], inv, -t4.z));
                    m = fmaxf(m, fmaf(v[j4 * 4 + 3], inv, -t4.w));
                }
                if (__any_sync(0xffffffffu, row_ok && m > 0.f)) {
                    // survivors of this chunk: branch-free bit mask per lane, then one warp-uniform round per column
                    // that has any; survivors are staged in shared memory and appended in groups of >= 32.
                    uint32_t hm = 0u;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 t4 = lds_f4(th_addr + (c32 + j4 * 4) * 4);
                        hm |= (v[j4 * 4 + 0] * inv > t4.x ? 1u : 0u) << (j4 * 4 + 0);
                        hm |= (v[j4 * 4 + 1] * inv > t4.y ? 1u : 0u) << (j4 * 4 + 1);
                        hm |= (v[j4 * 4 + 2] * inv > t4.z ? 1u : 0u) << (j4 * 4 + 2);
                        hm |= (v[j4 * 4 + 3] * inv > t4.w ? 1u : 0u) << (j4 * 4 + 3);
                    }
                    if (!row_ok) hm = 0u;
                    uint32_t any = __reduce_or_sync(0xffffffffu, hm);
                    const int q0 = qb * NQB + col0 + c32;
                    if (__reduce_add_sync(0xffffffffu, __popc(hm)) >= p.direct_min) {
                        // a dense chunk (the early phases, whose thresholds are still weak): ONE reservation per query
                        // column -- lane j reserves for column j, all 32 atomics in flight together -- and the entries go
                        // straight to the buffers; the staged path would pay an atomic round trip per 32 survivors.
                        uint32_t mine = 0u;
                        for (uint32_t rest = any; rest; rest &= rest - 1u) {
                            const int j = __ffs(rest) - 1;
                            const uint32_t who = __ballot_sync(0xffffffffu, (hm >> j) & 1u);
                            if (lane == j) mine = who;
                        }
                        int base = 0;
                        if (mine) base = atomicAdd(p.cnt + q0 + lane, __popc(mine));
                        for (uint32_t rest = any; rest; rest &= rest - 1u) {
                            const int j = __ffs(rest) - 1;
                            const uint32_t who = __shfl_sync(0xffffffffu, mine, j);
                            const int b0 = __shfl_sync(0xffffffffu, base, j);
                            const float val = pick32(v, j) * inv;
                            if ((hm >> j) & 1u) {
                                const int slot = b0 + __popc(who & ((1u << lane) - 1u));
                                if (slot < p.cap) p.buf[size_t(q0 + j) * p.cap + slot] = pack_entry(val, uint32_t(row));
                            }
                        }
                        any = 0u;
                    }
                    while (any) {
                        const int j = __ffs(any) - 1;
                        any &= any - 1u;
                        const bool hit = (hm >> j) & 1u;
                        const uint32_t who = __ballot_sync(0xffffffffu, hit);
                        if (hit) {
                            const uint64_t e = pack_entry(pick32(v, j) * inv, uint32_t(row));
                            stg[n_stg + __popc(who & ((1u << lane) - 1u))] = make_uint4(uint32_t(e), uint32_t(e >> 32), uint32_t(q0 + j), 0u);
                        }
                        n_stg += __popc(who);
                        __syncwarp();
                        if (n_stg >= 32) {
                            flush_staged(stg, n_stg, p.buf, p.cnt, p.cap);
                            n_stg = 0;
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(a ? te1 : te0);  // this warp is done with the accumulator (leader's barrier)
                sched_release(it);                   // ... and with ring slot `it` (slot it + 1 stays held until the next round)
            }
            w = w2;
            if (w < 0 && lane == 0) sched_release(it + 1);  // the terminator's slot
        }
        if (n_stg > 0) flush_staged(stg, n_stg, p.buf, p.cnt, p.cap);
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody frees TMEM or exits while the peer may still signal / read
    if (warp == 9) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS))
                     : "memory");
    }
}

// ------------------------------------------------------------------ per-query cut between phases
__global__ void gemm_init_kernel(int* cnt, float* tau, int* ovf, int n_q, int n_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) {
        cnt[i] = 0;
        ovf[i] = 0;
        tau[i] = i < n_q ? -INFINITY : INFINITY;  // padding columns never pass
    }
}

// One block per query: keep the K' best entries (key desc, id asc) at the front of the buffer, tau := K'-th key.
// The K'-th largest packed entry is found by an MSB-first radix select (8 bits per pass, stops as soon as the
// current digit bin holds exactly the entries still wanted -- normally after the 4 passes over the score bits);
// nothing is sorted between phases.  final: the K' survivors are sorted and emitted in the stage-1 output format.
constexpr int CUT_THREADS = 512;
__global__ void __launch_bounds__(CUT_THREADS)
gemm_cut_kernel(unsigned long long* buf, int* cnt, float* tau, int* ovf, int cap, int kp, int final_pass, int64_t id_base,
                int64_t* out_ids, float* out_approx, float* out_thresh) {
    extern __shared__ uint64_t cut_s[];  // [cap] entries, then [kp] survivors
    __shared__ int hist[256];
    __shared__ int sh_bin, sh_want, sh_n_sel;
    __shared__ unsigned long long sh_min;
    uint64_t* ent = cut_s;
    uint64_t* sel = cut_s + cap;
    const int q = blockIdx.x, t = threadIdx.x;
    const int raw = cnt[q];
    const int n = min(raw, cap);
    unsigned long long* B = buf + size_t(q) * cap;
    const bool over = raw > cap || ovf[q] != 0;
    if (t == 0 && over) ovf[q] = 1;
    if (n <= kp && !final_pass) return;  // nothing to drop yet

    for (int i = t; i < n; i += CUT_THREADS) ent[i] = B[i];
    if (t == 0) {
        sh_n_sel = 0;
        sh_min = ~0ull;
    }
    __syncthreads();
    int n_keep = n;
    if (n > kp) {
        uint64_t prefix = 0ull, mask = 0ull;
        int want = kp;  // still to take among the entries matching `prefix` under `mask`
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (t < 256) hist[t] = 0;
            __syncthreads();
            for (int i = t; i < n; i += CUT_THREADS) {
                const uint64_t e = ent[i];
                if ((e & mask) == prefix) atomicAdd(&hist[int(e >> shift) & 255], 1);
            }
            __syncthreads();
            if (t < 32) {  // bins from 255 down: lane l owns bins 255-8l .. 248-8l
                int local[8], sum = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    local[u] = hist[255 - 8 * t - u];
                    sum += local[u];
                }
                int inc = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (t >= d) inc += o;
                }
                int before = inc - sum;  // entries in bins above this lane's
                if (before < want && want <= inc) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (before < want && want <= before + local[u]) {
                            sh_bin = 255 - 8 * t - u;
                            sh_want = want - before;
                        }
                        before += local[u];
                    }
                }
            }
            __syncthreads();
            const int bin = sh_bin;
            want = sh_want;
            prefix |= uint64_t(bin) << shift;
            mask |= 0xffull << shift;
            if (hist[bin] == want) break;  // every entry matching the prefix is wanted: selection = {e & mask >= prefix}
            __syncthreads();
        }
        for (int i = t; i < n; i += CUT_THREADS) {
            const uint64_t e = ent[i];
            if ((e & mask) >= prefix) {
                const int slot = atomicAdd(&sh_n_sel, 1);
                if (slot < kp) sel[slot] = e;  // (entries are distinct, so exactly kp match; the guard is for safety)
                atomicMin(&sh_min, static_cast<unsigned long long>(e));
            }
        }
        __syncthreads();
        n_keep = kp;  // == sh_n_sel: entries are distinct
        if (!final_pass)
            for (int i = t; i < kp; i += CUT_THREADS) B[i] = sel[i];
        if (t == 0) {
            cnt[q] = kp;
            tau[q] = entry_key(sh_min);
        }
    } else {
        for (int i = t; i < n; i += CUT_THREADS) sel[i] = ent[i];
    }
    if (final_pass) {
        int n2 = 1;
        while (n2 < n_keep) n2 <<= 1;
        for (int i = n_keep + t; i < n2; i += CUT_THREADS) sel[i] = 0ull;
        __syncthreads();
        if (n2 > 1) block_bitonic_sort_desc(sel, n2);
        for (int i = t; i < kp; i += CUT_THREADS) {
            const uint64_t e = i < n_keep ? sel[i] : 0ull;
            out_ids[size_t(q) * kp + i] = e ? int64_t(id_base + entry_id(e)) : int64_t(-1);
            out_approx[size_t(q) * kp + i] = e ? entry_key(e) : -INFINITY;
        }
        if (t == 0) out_thresh[q] = over ? INFINITY : (n > kp ? entry_key(sh_min) : tau[q]);
    }
}

static uint32_t pick_perm_mul(uint32_t n) {
    if (n <= 2) return 1;
    uint32_t m = uint32_t(double(n) * 0.6180339887);
    if (m < 1) m = 1;
    auto gcd = [](uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; };
    while (gcd(m, n) != 1) ++m;
    return m % n ? m % n : 1;
}

}  // namespace tc3

// Entries a query's survivor buffer holds.  A phase appends about growth x K' entries per query (the rows that beat the
// K'-th score of the rows seen before it), so 16 K' leaves a 4x margin at growth 4.  Small batches -- the HBM-bound
// regime, where every extra phase costs a launch + ramp + cut (~15 us of a ~2.9 ms pass at 10M rows) -- get twice the
// buffer and with it a dense first phase four to eight times as long and growth 8 at the same 4x margin: 5 phases
// instead of 7 at 10M rows.  Measured at 10M rows: 32 queries hi+lo 2.94 -> 2.91 ms, 64 queries 3.01 -> 2.98 ms (top-10),
// 3.21 -> 3.07 ms (top-100).  Not beyond 64 queries: a dense tile appends 256 x n_q entries through the staging
// buffers, and at 128 queries the longer dense phase costs more than the phases saved (3.20 -> 3.46 ms).
static inline bool gemm_small_batch(int n_q, int kprime) {
    static const int max_q = [] {
        const char* e = getenv("TT_GEMM_SMALL_MAX");  // tuning knob
        return e ? atoi(e) : 64;
    }();
    return n_q <= max_q && kprime <= 256;
}
static inline int gemm_cap(int n_q, int kprime) { return (gemm_small_batch(n_q, kprime) ? 32 : 16) * kprime; }

size_t scan_gemm_workspace_bytes(int n_q, int kprime) {
    const size_t n_pad = size_t((n_q + tc3::NQB_MAX - 1) / tc3::NQB_MAX) * tc3::NQB_MAX;
    return 3 * n_pad * 4 + 16 /* work-item counter */ + size_t(n_q) * size_t(gemm_cap(n_q, kprime)) * 8;
}

bool scan_gemm_supported(int dim, int kprime, int n_lists) {
    return n_lists >= 2 && dim % 64 == 0 && dim >= 64 && (kprime == 128 || kprime == 256 || kprime == 512);
}

namespace tc3 {

template <int NQB, int CH, bool HILO = false>
static int run_phases(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                      int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx, float* out_thresh,
                      void* ws, int n_sms, cudaStream_t st, const void* q_lo = nullptr) {
    constexpr int NH = NQB / 2;
    constexpr int STAGE_BYTES = CH * (CHUNK_BYTES + NH * 128);
    const int n_qb = (n_q + NQB - 1) / NQB;
    const int n_pad = (n_q + NQB_MAX - 1) / NQB_MAX * NQB_MAX;  // the layout scan_gemm_workspace_bytes() sized
    const bool small = gemm_small_batch(n_q, kprime) && !getenv("TT_GEMM_SMALL_OFF");
    const int cap = gemm_cap(n_q, kprime);
    int* cnt = reinterpret_cast<int*>(ws);
    float* tau = reinterpret_cast<float*>(cnt + n_pad);
    int* ovf = reinterpret_cast<int*>(tau + n_pad);
    int* sched = ovf + n_pad;  // 16 bytes: the work-item counter of the dynamic scheduler (zeroed before every phase)
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(sched + 4);

    gemm_init_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(cnt, tau, ovf, n_q, n_pad);
    TT_LAUNCH_OK("gemm_init_kernel");

    int stages = (tc::SMEM_LIMIT - 1024 - 2 * NQB * 4 - (EPI3 / 32) * STG_CAP * 16 - 384) / (STAGE_BYTES + 16);
    if (stages > 12) stages = 12;
    if (const char* e = getenv("TT_GEMM_STAGES")) {
        const int want = atoi(e);
        if (want >= 2 && want < stages) stages = want;
    }
    const size_t smem = 1024 + size_t(stages) * STAGE_BYTES + 2 * NQB * 4 + (EPI3 / 32) * STG_CAP * 16 +
                        (2 * size_t(stages) + 4 + 2 * SCHED_SLOTS) * 8 + 16 + SCHED_SLOTS * 4;

    Params p;
    p.inv_norm = inv_norm;
    p.n_rows = n_rows;
    p.n_super = int((n_rows + 255) / 256);
    p.n_chunks = dim / tc::CHUNK_COLS;
    p.stages = stages;
    p.n_qb = n_qb;
    p.perm_mul = pick_perm_mul(uint32_t(p.n_super));
    // Static interleave by default.  Measured at 10M rows (profiles/r02_gemm_sweep_*.log): claiming items dynamically
    // moves 64 queries from 2.97 to 2.95 ms and 4096 x top-100 from 63.4 to 63.1 ms (noise) but costs 128 queries 5 %
    // (3.22 -> 3.40 ms: the publisher's cluster-scope release sits on the TMA issue path once per item) -- the
    // clusters of this kernel do not drift apart the way the single-CTA scan's did.  TT_GEMM_DYNAMIC=1 turns it on.
    p.sched = getenv("TT_GEMM_DYNAMIC") ? sched : nullptr;
    p.tau = tau;
    p.buf = buf;
    p.cnt = cnt;
    p.cap = cap;
    p.direct_min = DIRECT_MIN;
    if (const char* e = getenv("TT_GEMM_DIRECT_MIN")) p.direct_min = atoi(e) > 0 ? atoi(e) : DIRECT_MIN;

    CUtensorMap map_c, map_q, map_q2;
    if (n_rows > 0) {
        int rc = tc::make_map(&map_c, corpus, n_rows, dim, stride, tc::TILE_ROWS, CH);
        if (rc) return rc;
        rc = tc::make_map(&map_q, q_hi, n_q, dim, dim, NH, CH);
        if (rc) return rc;
        rc = tc::make_map(&map_q2, HILO ? q_lo : q_hi, n_q, dim, dim, NH, CH);
        if (rc) return rc;
    }
    auto kern = scan_gemm_kernel<NQB, CH, HILO>;
    TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
    TT_CUDA_OK(cudaFuncSetAttribute(gemm_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (cap + kprime) * 8));

    // phases: the first visits ~4 K' rows (small batches: as many rows as the buffer holds -- it is dense, every row of
    // it is appended), each later one `growth` times the rows visited before it
    int growth = small ? 8 : 4;
    if (const char* e = getenv("TT_GEMM_GROWTH")) {
        const int g = atoi(e);
        if (g >= 1 && g <= 64) growth = g;
    }
    const int grid = n_sms & ~1;
    int seen = 0;
    int next = small ? cap / 256 : (4 * kprime + 255) / 256;
    bool done = p.n_super == 0;
    if (done) {
        gemm_cut_kernel<<<n_q, CUT_THREADS, (cap + kprime) * 8, st>>>(buf, cnt, tau, ovf, cap, kprime, 1, id_base, out_ids,
                                                                      out_approx, out_thresh);
        TT_LAUNCH_OK("gemm_cut_kernel");
    }
    while (!done) {
        int upto = seen + next;
        if (upto >= p.n_super || p.n_super - upto < next / 2) upto = p.n_super;  // fold a short tail into this phase
        p.i0 = seen;
        p.i1 = upto;
        if (p.sched) TT_CUDA_OK(cudaMemsetAsync(p.sched, 0, 16, st));
        kern<<<grid, THREADS3, smem, st>>>(map_c, map_q, map_q2, p);
        TT_LAUNCH_OK("scan_gemm_kernel");
        done = upto == p.n_super;
        gemm_cut_kernel<<<n_q, CUT_THREADS, (cap + kprime) * 8, st>>>(buf, cnt, tau, ovf, cap, kprime, done ? 1 : 0, id_base,
                                                                      out_ids, out_approx, out_thresh);
        TT_LAUNCH_OK("gemm_cut_kernel");
        seen = upto;
        next = seen * growth;
    }
    return TT_OK;
}

}  // namespace tc3

int scan_gemm_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                     const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx,
                     float* out_thresh, void* ws, int n_sms, cudaStream_t st) {
    if (q_lo) {  // hi+lo: <= 32 queries as 64 MMA columns, two 64-column chunks per ring stage when dim allows
        if (dim % 128 == 0)
            return tc3::run_phases<64, 2, true>(corpus, n_rows, dim, stride, inv_norm, q_hi, n_q, kprime, id_base, out_ids, out_approx,
                                                out_thresh, ws, n_sms, st, q_lo);
        return tc3::run_phases<64, 1, true>(corpus, n_rows, dim, stride, inv_norm, q_hi, n_q, kprime, id_base, out_ids, out_approx,
                                            out_thresh, ws, n_sms, st, q_lo);
    }
    int width = n_q <= 64 ? 64 : n_q <= 128 ? 128 : 256;  // one narrower query block when the batch fits it
    if (const char* e = getenv("TT_GEMM_WIDTH")) {         // tuning knob
        const int w = atoi(e);
        if (w == 64 || w == 128 || w == 256) width = w;
    }
    // two 64-column chunks per ring stage (256 B contiguous per corpus / query row per TMA box) whenever dim allows:
    // measured at 10M rows, 128 queries 3.89 -> 3.24 ms, 4096 queries (width 256, 3 stages of 64 KB) 69.4 -> 64.5 ms
    int ch = dim % 128 == 0 ? 2 : 1;
    if (const char* e = getenv("TT_GEMM_CH")) {
        const int c = atoi(e);
        if (c == 1 || (c == 2 && dim % 128 == 0)) ch = c;
    }
#define TT_GEMM(W, C)                                                                                                    \
    return tc3::run_phases<W, C>(corpus, n_rows, dim, stride, inv_norm, q_hi, n_q, kprime, id_base, out_ids, out_approx, \
                                 out_thresh, ws, n_sms, st)
    if (width == 64 && ch == 2) TT_GEMM(64, 2);
    if (width == 64) TT_GEMM(64, 1);
    if (width == 128 && ch == 2) TT_GEMM(128, 2);
    if (width == 128) TT_GEMM(128, 1);
    if (ch == 2) TT_GEMM(256, 2);
    TT_GEMM(256, 1);
#undef TT_GEMM
}

}  // namespace tt
