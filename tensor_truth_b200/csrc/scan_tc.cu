// Stage 1, tensor-core variant: TMA-staged corpus tiles -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
// -> fused on-chip top-K' shortlist.  The score matrix never leaves the SM.
//
// Replaces the vector-store query issued through `index.as_retriever(similarity_top_k=k)` at
// /root/reference/src/tensortruth/rag_engine.py:639 (ChromaVectorStore.query -> collection.query);
// the exact result is restored by stage 2 (rescore.cu).
//
// Shape of one launch ("pass"): up to NQ = N/2 queries, each occupying two MMA columns (hi and lo
// bf16 halves of the fp32 query), against every 128-row tile of the corpus:
//
//      D[128 rows, N] (TMEM, fp32)  =  C_tile[128, dim] (smem via TMA, K-major SW128)  x  Q^T[dim, N] (smem, resident)
//
// Persistent grid, one CTA per SM, tiles interleaved (tile t -> CTA t % grid).  Six warps:
//   warps 0-3  epilogue: tcgen05.ld the accumulator row of "their" corpus row, apply inv_norm, filter
//              against the per-query threshold, push survivors into a shared-memory candidate list;
//              rank-select the list back to K' when it fills (threshold := K'-th approximate score)
//   warp  4    TMA producer: corpus tile ring (STAGES x CH x 16 KB) + the query block once
//   warp  5    TMEM allocation + single-thread tcgen05.mma issue, double-buffered accumulators
#include "tc_ptx.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(scan_tc)

namespace tc {

constexpr int MAX_SEGMENTS = 16;

struct Params {
    const float* inv_norm;
    int64_t n_rows;
    int64_t id_base;
    int n_tiles;
    int n_chunks;    // dim / 64
    int stages;      // ring depth
    int q0;          // first query of this pass
    int nq_here;     // queries in this pass (<= N/2)
    int n_q;         // total queries (output row count)
    int kprime;
    int cap;         // candidate list capacity per query: kprime + spare, <= 256
    int has_lo;
    int64_t* out_ids;
    float* out_approx;
    float* out_thresh;
    int* sched;      // {next-tile counter, finished-CTA counter}, both zero between launches; NULL = static interleave
    // Segmented corpus (several indexes concatenated): rows [seg_end[s-1], seg_end[s]) form segment s and
    // every (segment, query) pair keeps its own shortlist -- a "virtual query" v = s * n_q + q in the outputs.
    int n_seg;       // >= 1
    int seg_end[MAX_SEGMENTS];
};

constexpr int SCHED_SLOTS = 4;  // tile-id ring between the producer and the MMA / epilogue roles

// Dynamic shared memory (base rounded up to 1024 B):
//   [ Q: n_chunks x N x 128 B ][ ring: stages x CH x 16 KB ][ lists: NQ x cap x 8 B ][ thresh NQ f32 ][ cnt NQ i32 ]
//   [ barriers: full[stages], empty[stages], q_full, tmem_full[2], tmem_empty[2], sched_full[4], sched_empty[4] ]
//   [ tmem base ][ tile ring int[4] ]
// Tiles are handed out by the producer: tile = blockIdx.x first, then (dynamic) gridDim.x + atomicAdd(counter) or
// (static) +gridDim.x; the id travels to the other roles through a 4-slot ring, -1 ends the stream.
template <int N, int CH, bool HILO>
__global__ void __launch_bounds__(THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_qhi,
               const __grid_constant__ CUtensorMap map_qlo, const Params p) {
    constexpr int NQ = HILO ? N / 2 : N;  // queries per pass: two MMA columns each (hi, lo) or one (hi only)
    constexpr int STAGE_BYTES = CH * CHUNK_BYTES;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

    const int q_bytes = p.n_chunks * N * 128;
    unsigned char* q_s = smem;
    unsigned char* ring = q_s + q_bytes;
    const int NL = p.n_seg * NQ;  // shortlists: one per (segment, query column)
    uint64_t* lists = reinterpret_cast<uint64_t*>(ring + size_t(p.stages) * STAGE_BYTES);
    float* thresh_s = reinterpret_cast<float*>(lists + size_t(NL) * p.cap);
    int* cnt_s = reinterpret_cast<int*>(thresh_s + NL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(cnt_s + NL);  // 8-byte aligned: NQ is a multiple of 8
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* q_full = bars + 2 * p.stages;
    uint64_t* tmem_full = q_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* sched_full = tmem_empty + 2;
    uint64_t* sched_empty = sched_full + SCHED_SLOTS;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(sched_empty + SCHED_SLOTS);
    volatile int* tile_ring = reinterpret_cast<volatile int*>(tmem_base_s + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int steps_per_tile = p.n_chunks / CH;
    constexpr int TMEM_COLS = (2 * N < 32) ? 32 : 2 * N;  // power of two for N in {16, 32, 64}
    pdl_launch_dependents();  // stage 2 may become resident next to this CTA (tt_common.cuh, PDL); it waits for this grid to end

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);
            mbar_init(smem_u32(empty_bar + s), 1);
        }
        mbar_init(smem_u32(q_full), 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full + a), 1);
            mbar_init(smem_u32(tmem_empty + a), EPI_THREADS);
        }
        for (int r = 0; r < SCHED_SLOTS; ++r) {
            mbar_init(smem_u32(sched_full + r), 1);
            mbar_init(smem_u32(sched_empty + r), 1 + EPI_THREADS / 32);  // MMA thread + one lane per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int l = threadIdx.x; l < NL; l += THREADS) {
        thresh_s[l] = (l % NQ) < p.nq_here ? -INFINITY : INFINITY;  // unused columns never pass
        cnt_s[l] = 0;
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                     "r"(uint32_t(TMEM_COLS))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 4) {
        // ===================================================== TMA producer
        if (lane == 0) {
            pdl_wait();  // launched as a programmatic dependent of prepare_queries: q_hi / q_lo are complete from here on
            mbar_expect_tx(smem_u32(q_full), uint32_t(p.n_chunks * N * 128));
            for (int c = 0; c < p.n_chunks; ++c) {
                tma_load_3d(smem_u32(q_s + size_t(c) * N * 128), &map_qhi, 0, p.q0, c, smem_u32(q_full), POLICY_EVICT_LAST);
                if (HILO)
                    tma_load_3d(smem_u32(q_s + size_t(c) * N * 128 + NQ * 128), &map_qlo, 0, p.q0, c, smem_u32(q_full),
                                POLICY_EVICT_LAST);
            }
            int stage = 0;
            uint32_t phase = 0;
            int tile = int(blockIdx.x);
            for (int it = 0;; ++it) {
                if (tile >= p.n_tiles) tile = -1;
                const int slot = it & (SCHED_SLOTS - 1);
                mbar_wait(smem_u32(sched_empty + slot), (uint32_t(it / SCHED_SLOTS) & 1u) ^ 1u);
                tile_ring[slot] = tile;
                mbar_arrive(smem_u32(sched_full + slot));  // release: publishes the tile id
                if (tile < 0) break;
                // claim the next tile now; the atomic's latency hides behind this tile's loads
                const int next = p.sched ? int(gridDim.x) + atomicAdd(p.sched, 1) : tile + int(gridDim.x);
                for (int s = 0; s < steps_per_tile; ++s) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
                    mbar_expect_tx(smem_u32(full_bar + stage), STAGE_BYTES);
                    tma_load_3d(smem_u32(ring + size_t(stage) * STAGE_BYTES), &map_c, 0, tile * TILE_ROWS, s * CH,
                                smem_u32(full_bar + stage), POLICY_EVICT_FIRST);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                tile = next;
            }
            if (p.sched) {  // the last CTA to get here re-arms the counters for the next launch
                __threadfence();
                if (atomicAdd(p.sched + 1, 1) == int(gridDim.x) - 1) {
                    p.sched[0] = 0;
                    p.sched[1] = 0;
                    __threadfence();
                }
            }
        }
    } else if (warp == 5) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(N);
            mbar_wait(smem_u32(q_full), 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0;; ++i) {
                const int slot = i & (SCHED_SLOTS - 1);
                mbar_wait(smem_u32(sched_full + slot), uint32_t(i / SCHED_SLOTS) & 1u);
                const int tile_id = tile_ring[slot];
                mbar_arrive(smem_u32(sched_empty + slot));
                if (tile_id < 0) break;
                const int a = i & 1;
                mbar_wait(smem_u32(tmem_empty + a), (uint32_t(i >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(a * N);
                for (int s = 0; s < steps_per_tile; ++s) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(ring + size_t(stage) * STAGE_BYTES);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const uint32_t b_base = smem_u32(q_s + size_t(s * CH + c) * N * 128);
#pragma unroll
                        for (int k = 0; k < CHUNK_COLS / 16; ++k) {
                            umma_bf16(d_tmem, umma_desc_sw128(a_base + c * CHUNK_BYTES + k * 32),
                                      umma_desc_sw128(b_base + k * 32), idesc, uint32_t((s | c | k) != 0));
                        }
                    }
                    umma_commit(smem_u32(empty_bar + stage));  // frees the smem slot once these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(smem_u32(tmem_full + a));  // accumulator of this tile complete
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: 128 threads, thread = corpus row
        const int t = threadIdx.x;
        const int nq = p.nq_here;
        const int kp = p.kprime, cap = p.cap;
        for (int i = 0;; ++i) {
            const int slot = i & (SCHED_SLOTS - 1);
            mbar_wait(smem_u32(sched_full + slot), uint32_t(i / SCHED_SLOTS) & 1u);
            const int tile = tile_ring[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(sched_empty + slot));
            if (tile < 0) break;
            const int a = i & 1;
            const int64_t row = int64_t(tile) * TILE_ROWS + t;
            const bool row_ok = row < p.n_rows;
            float inv = 1.f;
            if (p.inv_norm && row_ok) inv = __ldg(p.inv_norm + row);

            mbar_wait(smem_u32(tmem_full + a), uint32_t(i >> 1) & 1u);
            tcgen05_fence_after();
            float acc[N];
            const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(a * N);
#pragma unroll
            for (int c = 0; c < N; c += 16) tmem_ld_x16(taddr + c, acc + c);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(smem_u32(tmem_empty + a));  // accumulator is in registers: hand TMEM back

            int seg = 0;  // the row's segment picks the block of lists it competes in (uniform per tile except at a boundary)
            while (seg < p.n_seg - 1 && row >= p.seg_end[seg]) ++seg;
            filter_and_push<NQ, HILO, N>(acc, inv, row_ok, uint32_t(row), NL, lists, cnt_s, thresh_s, kp, cap, warp, lane,
                                         seg * NQ);
        }

        // ---- final cut of every list to its K' best, then emit this CTA's shortlists
        epi_bar_sync();
        cut_lists(lists, cnt_s, thresh_s, NL, kp, cap, warp, lane, true);
        epi_bar_sync();
        for (int sg = 0; sg < p.n_seg; ++sg) {
            for (int j = 0; j < nq; ++j) {
                const int l = sg * NQ + j;
                const int n = cnt_s[l];
                const size_t v = size_t(sg) * p.n_q + size_t(p.q0 + j);  // virtual query
                const size_t o = (v * gridDim.x + blockIdx.x) * kp;
                for (int s = t; s < kp; s += EPI_THREADS) {
                    const uint64_t e = s < n ? lists[size_t(l) * cap + s] : 0ull;
                    p.out_ids[o + s] = e ? int64_t(p.id_base + entry_id(e)) : int64_t(-1);
                    p.out_approx[o + s] = e ? entry_key(e) : -INFINITY;
                }
                if (t == 0) p.out_thresh[v * gridDim.x + blockIdx.x] = thresh_s[l];
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS))
                     : "memory");
    }
}

// ------------------------------------------------------------------ host side
// Shared-memory budget: the resident query block, then the candidate lists, the rest is the TMA ring.
// Lists hold K' + spare entries; spare = 128 never needs a retry (a tile pushes at most 128 rows per query);
// with many queries per pass the spare shrinks so that the ring keeps enough bytes in flight.
// CO_RESIDENT_RESERVE bytes of the SM's shared memory are left to other kernels when that costs no ring stage worth
// having: in a two-lane pipeline the tail of step i (re-score/select, merge/auto-merge: <= 18 KB, 512 threads) has to
// become resident NEXT TO the persistent scan CTAs of step i + 1, or it would wait for the whole scan to drain.
constexpr size_t CO_RESIDENT_RESERVE = 20 * 1024;

template <int N, int CH, bool HILO>
static bool plan_within(size_t limit, int dim, int kprime, int* spare_out, int* stages_out, size_t* smem_out, int n_seg);

template <int N, int CH, bool HILO>
static bool plan(int dim, int kprime, int* spare_out, int* stages_out, size_t* smem_out, int n_seg = 1) {
    int sp_full = 0, st_full = 0;
    size_t sm_full = 0;
    if (!plan_within<N, CH, HILO>(size_t(SMEM_LIMIT), dim, kprime, &sp_full, &st_full, &sm_full, n_seg)) return false;
    int sp = 0, st_n = 0;
    size_t sm = 0;
    // keep the reserve if the ring stays >= 128 KB, or loses nothing at all
    if (!getenv("TT_SCAN_NO_RESERVE") &&
        plan_within<N, CH, HILO>(size_t(SMEM_LIMIT) - CO_RESIDENT_RESERVE, dim, kprime, &sp, &st_n, &sm, n_seg) &&
        (st_n == st_full || size_t(st_n) * CH * CHUNK_BYTES >= 128 * 1024) && sp == sp_full) {
        *spare_out = sp;
        *stages_out = st_n;
        *smem_out = sm;
        return true;
    }
    *spare_out = sp_full;
    *stages_out = st_full;
    *smem_out = sm_full;
    return true;
}

template <int N, int CH, bool HILO>
static bool plan_within(size_t limit, int dim, int kprime, int* spare_out, int* stages_out, size_t* smem_out, int n_seg) {
    const int NQ = (HILO ? N / 2 : N) * n_seg;  // shortlists per CTA
    const size_t q_bytes = size_t(dim / CHUNK_COLS) * N * 128;
    const size_t base = 1024 /*align slack*/ + q_bytes + size_t(NQ) * 8 + 160;
    const size_t per_stage = size_t(CH) * CHUNK_BYTES + 16;
    int spare = 0, stages = 0;
    auto try_spare = [&](int sp, size_t want_ring) {
        if (kprime + sp > 256) return false;
        const size_t fixed = base + size_t(NQ) * (kprime + sp) * 8;
        if (fixed + 2 * per_stage > limit) return false;
        const int st_n = int((limit - fixed) / per_stage);
        if (size_t(st_n) * CH * CHUNK_BYTES < want_ring) return false;
        spare = sp;
        stages = st_n;
        return true;
    };
    // bytes in flight are what buys HBM bandwidth: the largest spare that still leaves >= 128 KB of ring, then
    // >= 96 KB, else a small spare and whatever ring is left
    if (!try_spare(128, 128 * 1024) && !try_spare(64, 128 * 1024) && !try_spare(32, 128 * 1024) &&
        !try_spare(128, 96 * 1024) && !try_spare(64, 96 * 1024) && !try_spare(32, 96 * 1024) && !try_spare(32, 0) &&
        !try_spare(16, 0))
        return false;
    if (stages > 24) stages = 24;
    if (const char* e = getenv("TT_SCAN_STAGES")) {  // tuning knob
        const int want = atoi(e);
        if (want >= 2 && want < stages) stages = want;
    }
    *spare_out = spare;
    *stages_out = stages;
    *smem_out = base + size_t(NQ) * (kprime + spare) * 8 + size_t(stages) * per_stage;
    return true;
}

template <int N, int CH, bool HILO>
static int launch(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                  const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx,
                  float* out_thresh, int n_lists, int* sched, cudaStream_t st, int n_seg = 1, const int64_t* seg_end = nullptr) {
    constexpr int NQ = HILO ? N / 2 : N;
    int spare = 0, stages = 0;
    size_t smem = 0;
    if (!plan<N, CH, HILO>(dim, kprime, &spare, &stages, &smem, n_seg)) {
        set_error("scan_tc: dim=%d kprime=%d does not fit shared memory with N=%d", dim, kprime, N);
        return TT_ERR_UNSUPPORTED;
    }
    Params p;
    p.sched = sched;
    p.inv_norm = inv_norm;
    p.n_rows = n_rows;
    p.id_base = id_base;
    p.n_tiles = int((n_rows + TILE_ROWS - 1) / TILE_ROWS);
    p.n_chunks = dim / CHUNK_COLS;
    p.n_q = n_q;
    p.kprime = kprime;
    p.has_lo = HILO;
    p.out_ids = out_ids;
    p.out_approx = out_approx;
    p.out_thresh = out_thresh;
    p.cap = kprime + spare;
    p.stages = stages;
    p.n_seg = n_seg;
    for (int i = 0; i < MAX_SEGMENTS; ++i) p.seg_end[i] = (seg_end && i < n_seg) ? int(seg_end[i]) : int(n_rows);

    CUtensorMap map_c, map_qhi, map_qlo;
    int rc = make_map(&map_c, corpus, n_rows, dim, stride, TILE_ROWS, CH);
    if (rc) return rc;
    rc = make_map(&map_qhi, q_hi, n_q, dim, dim, NQ, 1);
    if (rc) return rc;
    rc = make_map(&map_qlo, HILO ? q_lo : q_hi, n_q, dim, dim, NQ, 1);
    if (rc) return rc;

    auto kern = scan_tc_kernel<N, CH, HILO>;
    TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    for (int q0 = 0; q0 < n_q; q0 += NQ) {
        p.q0 = q0;
        p.nq_here = (n_q - q0 < NQ) ? n_q - q0 : NQ;
        // only the first pass may overlap its predecessor (prepare_queries): later passes share the scheduler counters
        TT_CUDA_OK(launch_kernel(kern, dim3(n_lists), dim3(THREADS), smem, st, q0 == 0, map_c, map_qhi, map_qlo, p));
        TT_LAUNCH_OK("scan_tc_kernel");
    }
    return TT_OK;
}

}  // namespace tc

bool scan_tc_supported(int64_t n_rows, int dim, int64_t stride, int kprime, const void* corpus) {
    return dim % 128 == 0 && dim >= 128 && dim <= 2048 && stride % 8 == 0 && kprime >= 1 && kprime <= 128 &&
           (reinterpret_cast<uintptr_t>(corpus) % 16) == 0 && n_rows > 0 && n_rows < (int64_t(1) << 31) - 256;
}

int scan_tc_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                   const void* q_hi, const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids,
                   float* out_approx, float* out_thresh, int n_lists, int* sched, cudaStream_t st) {
    if (!scan_tc_supported(n_rows, dim, stride, kprime, corpus)) {
        set_error("scan_tc: unsupported shape (n_rows=%lld dim=%d stride=%lld kprime=%d)", (long long)n_rows, dim,
                  (long long)stride, kprime);
        return TT_ERR_UNSUPPORTED;
    }
    if (const char* e = getenv("TT_SCAN_STATIC")) { if (atoi(e)) sched = nullptr; }  // tuning knob
#define TT_TC(NN, CC, HL)                                                                                              \
    return tc::launch<NN, CC, HL>(corpus, n_rows, dim, stride, inv_norm, q_hi, q_lo, n_q, kprime, id_base, out_ids, \
                                  out_approx, out_thresh, n_lists, sched, st)
    int sp, sg;
    size_t sm;
    if (q_lo) {  // two MMA columns per query: 8 / 16 / 32 queries per pass (the widest pass that fits shared memory)
        if (n_q <= 8) TT_TC(16, 2, true);
        if (n_q <= 16 || !tc::plan<64, 1, true>(dim, kprime, &sp, &sg, &sm)) TT_TC(32, 2, true);
        TT_TC(64, 1, true);
    }
    // hi only: 16 / 32 / 64 queries per pass, wider certificate
    if (n_q <= 16) TT_TC(16, 2, false);
    if (n_q <= 32 || !tc::plan<64, 1, false>(dim, kprime, &sp, &sg, &sm)) TT_TC(32, 2, false);
    TT_TC(64, 1, false);
#undef TT_TC
}

// Segmented corpus: n_seg <= 16 row ranges, up to 8 queries per pass (hi+lo), one shortlist per (segment, query).
int scan_tc_segmented(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                      const void* q_lo, int n_q, int kprime, int64_t id_base, const int64_t* seg_end, int n_seg,
                      int64_t* out_ids, float* out_approx, float* out_thresh, int n_lists, int* sched, cudaStream_t st) {
    if (!scan_tc_supported(n_rows, dim, stride, kprime, corpus) || !q_lo || n_seg < 1 || n_seg > tc::MAX_SEGMENTS) {
        set_error("scan_tc_segmented: unsupported shape (n_rows=%lld dim=%d stride=%lld kprime=%d n_seg=%d)",
                  (long long)n_rows, dim, (long long)stride, kprime, n_seg);
        return TT_ERR_UNSUPPORTED;
    }
    return tc::launch<16, 2, true>(corpus, n_rows, dim, stride, inv_norm, q_hi, q_lo, n_q, kprime, id_base, out_ids,
                                   out_approx, out_thresh, n_lists, sched, st, n_seg, seg_end);
}

}  // namespace tt
