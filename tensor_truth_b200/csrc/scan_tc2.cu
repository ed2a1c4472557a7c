// Stage 1, tensor-core variant for wide query batches: CTA PAIRS (tcgen05 cta_group::2).
//
// scan_tc.cu keeps the whole query block resident in every CTA's shared memory; at 64 queries that is
// 128 KB and leaves too little room for the TMA ring that has to keep ~7 TB/s of corpus in flight.  Here two
// CTAs on neighbouring SMs work as one MMA unit:
//
//      D[256 rows, N] (TMEM of both CTAs, 128 rows each) = C_tile[256, dim] x Q^T[dim, N]
//
// each CTA streams ITS OWN 128 corpus rows (A operand) but holds only HALF of the query block (N/2 rows of
// the B operand, 64 KB at N = 64) -- the tensor cores of the pair read both halves.  Everything after the
// MMA is per CTA and identical to scan_tc.cu: tcgen05.ld of the CTA's 128 accumulator rows, inv_norm, the
// threshold filter and the shared-memory candidate lists (tc_ptx.cuh).
//
// Queries travel as bf16 hi halves only (one MMA column per query): 64 queries per corpus pass, with the
// wider certificate bound (index.py EPS_HI_ONLY); queries it cannot prove are re-run through scan_tc.cu hi+lo.
//
// Replaces the vector-store query behind `index.as_retriever(similarity_top_k=k)` at
// /root/reference/src/tensortruth/rag_engine.py:639 for a batch of concurrent queries.
#include "tc_ptx.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(scan_tc2)
namespace tc2 {

using namespace tc;

struct Params {
    const float* inv_norm;
    int64_t n_rows;
    int64_t id_base;
    int n_super;     // 256-row super-tiles
    int n_chunks;    // dim / 64
    int stages;
    int q0, nq_here, n_q;
    int kprime, cap;
    int64_t* out_ids;
    float* out_approx;
    float* out_thresh;
};

// Dynamic shared memory of each CTA (base rounded up to 1024 B):
//   [ Q half: n_chunks x N/2 x 128 B ][ ring: stages x 16 KB ][ lists: N x cap x 8 B ][ thresh N f32 ][ cnt N i32 ]
//   [ barriers: full[stages] (used in CTA 0), empty[stages], q_full (CTA 0), tmem_full[2], tmem_empty[2] (CTA 0) ][ tmem base ]
// Ten warps: 0-7 epilogue (warps w and w+4 share the TMEM lane quarter w%4 and split the N query columns in
// halves -- two warps per scheduler hide each other's latencies), 8 TMA producer, 9 TMEM alloc + MMA issue.
constexpr int EPI2 = 256;
constexpr int THREADS2 = EPI2 + 64;

template <int N, int CH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_q, const Params p) {
    constexpr int NQ = N;      // hi only: one MMA column per query
    constexpr int NH = N / 2;  // query rows held by each CTA
    constexpr int STAGE_BYTES = CH * CHUNK_BYTES;  // CH 64-column chunks per ring stage (256 B contiguous per row at CH = 2)
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

    const int q_bytes = p.n_chunks * NH * 128;
    unsigned char* q_s = smem;
    unsigned char* ring = q_s + q_bytes;
    uint64_t* lists = reinterpret_cast<uint64_t*>(ring + size_t(p.stages) * STAGE_BYTES);
    float* thresh_s = reinterpret_cast<float*>(lists + size_t(NQ) * p.cap);
    int* cnt_s = reinterpret_cast<int*>(thresh_s + NQ);
    uint64_t* bars = reinterpret_cast<uint64_t*>(cnt_s + NQ);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* q_full = bars + 2 * p.stages;
    uint64_t* tmem_full = q_full + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = int(blockIdx.x) >> 1, n_clusters = int(gridDim.x) >> 1;
    const int n_my = (p.n_super > cluster_id) ? (p.n_super - 1 - cluster_id) / n_clusters + 1 : 0;
    const int steps_per_tile = p.n_chunks / CH;
    constexpr int TMEM_COLS = (2 * N < 32) ? 32 : 2 * N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(full_bar + s), 2);   // one arrive per CTA's producer; bytes of both land here (CTA 0)
            mbar_init(smem_u32(empty_bar + s), 1);  // multicast tcgen05.commit
        }
        mbar_init(smem_u32(q_full), 2);
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full + a), 1);
            mbar_init(smem_u32(tmem_empty + a), 2 * (EPI2 / 32));  // one lane per epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < NQ) {
        thresh_s[threadIdx.x] = int(threadIdx.x) < p.nq_here ? -INFINITY : INFINITY;  // unused columns never pass
        cnt_s[threadIdx.x] = 0;
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
        // ===================================================== TMA producer (both CTAs, own rows + own query half)
        if (lane == 0 && n_my > 0) {
            const uint32_t qf = mapa(smem_u32(q_full), 0);
            if (rank == 0) mbar_expect_tx(smem_u32(q_full), uint32_t(2 * q_bytes));
            else mbar_arrive_cluster(qf);
            for (int c = 0; c < p.n_chunks; ++c)
                tma_load_3d_pair(smem_u32(q_s + size_t(c) * NH * 128), &map_q, 0, p.q0 + int(rank) * NH, c, qf,
                                 POLICY_EVICT_LAST);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_my; ++i) {
                const int super = cluster_id + i * n_clusters;
                const int row0 = super * 256 + int(rank) * TILE_ROWS;
                for (int s = 0; s < steps_per_tile; ++s) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
                    const uint32_t fb = mapa(smem_u32(full_bar + stage), 0);
                    if (rank == 0) mbar_expect_tx(smem_u32(full_bar + stage), 2 * STAGE_BYTES);
                    else mbar_arrive_cluster(fb);
                    tma_load_3d_pair(smem_u32(ring + size_t(stage) * STAGE_BYTES), &map_c, 0, row0, s * CH, fb,
                                     POLICY_EVICT_FIRST);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== MMA issuer: one thread of the leader CTA
        if (lane == 0 && rank == 0 && n_my > 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_m256(N);
            mbar_wait(smem_u32(q_full), 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_my; ++i) {
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
                        const uint32_t b_base = smem_u32(q_s + size_t(s * CH + c) * NH * 128);
#pragma unroll
                        for (int k = 0; k < CHUNK_COLS / 16; ++k)
                            umma_bf16_pair(d_tmem, umma_desc_sw128(a_base + c * CHUNK_BYTES + k * 32),
                                           umma_desc_sw128(b_base + k * 32), idesc, uint32_t((s | c | k) != 0));
                    }
                    umma_commit_pair(smem_u32(empty_bar + stage));  // frees this ring slot in both CTAs
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(smem_u32(tmem_full + a));  // both CTAs' accumulator halves are complete
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: 8 warps per CTA; thread = (corpus row, half of the queries)
        const int quarter = warp & 3, half = warp >> 2;
        constexpr int NW = N / 2;  // query columns per warp
        const int t = threadIdx.x;
        const int nq = p.nq_here;
        const int kp = p.kprime, cap = p.cap;
        const uint32_t te0 = mapa(smem_u32(tmem_empty), 0), te1 = mapa(smem_u32(tmem_empty + 1), 0);
        for (int i = 0; i < n_my; ++i) {
            const int a = i & 1;
            const int super = cluster_id + i * n_clusters;
            const int64_t row = int64_t(super) * 256 + int64_t(rank) * TILE_ROWS + quarter * 32 + lane;
            const bool row_ok = row < p.n_rows;
            float inv = 1.f;
            if (p.inv_norm && row_ok) inv = __ldg(p.inv_norm + row);

            mbar_wait(smem_u32(tmem_full + a), uint32_t(i >> 1) & 1u);
            tcgen05_fence_after();
            float sc[NW];
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(a * N + half * NW);
#pragma unroll
            for (int c = 0; c < NW; c += 16) tmem_ld_x16(taddr + c, sc + c);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a ? te1 : te0);  // accumulator is in registers: hand TMEM back (leader's barrier)
            filter_and_push<NW, false, NW, EPI2>(sc, inv, row_ok, uint32_t(row), nq, lists, cnt_s, thresh_s, kp, cap, warp, lane,
                                                 half * NW);
        }

        // ---- final cut of every list to its K' best, then emit this CTA's shortlist
        epi_bar_sync<EPI2>();
        cut_lists<EPI2>(lists, cnt_s, thresh_s, nq, kp, cap, warp, lane, true);
        epi_bar_sync<EPI2>();
        for (int j = 0; j < nq; ++j) {
            const int n = cnt_s[j];
            const size_t o = (size_t(p.q0 + j) * gridDim.x + blockIdx.x) * kp;
            for (int s = t; s < kp; s += EPI2) {
                const uint64_t e = s < n ? lists[size_t(j) * cap + s] : 0ull;
                p.out_ids[o + s] = e ? int64_t(p.id_base + entry_id(e)) : int64_t(-1);
                p.out_approx[o + s] = e ? entry_key(e) : -INFINITY;
            }
            if (t == 0) p.out_thresh[size_t(p.q0 + j) * gridDim.x + blockIdx.x] = thresh_s[j];
        }
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

template <int N, int CH>
static bool plan(int dim, int kprime, int* spare_out, int* stages_out, size_t* smem_out) {
    const size_t q_bytes = size_t(dim / CHUNK_COLS) * (N / 2) * 128;
    const size_t base = 1024 + q_bytes + size_t(N) * 8 + 160;
    const size_t per_stage = size_t(CH) * CHUNK_BYTES + 16;
    const int spares[4] = {128, 64, 32, 16};
    // prefer >= 128 KB of ring (bytes in flight are what buys HBM bandwidth), then >= 96 KB, then whatever fits
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = 0; i < 4; ++i) {
            if (kprime + spares[i] > 256) continue;
            const size_t fixed = base + size_t(N) * (kprime + spares[i]) * 8;
            if (fixed + 2 * per_stage > size_t(SMEM_LIMIT)) continue;
            int st_n = int((size_t(SMEM_LIMIT) - fixed) / per_stage);
            if (pass == 0 && size_t(st_n) * CH * CHUNK_BYTES < 128 * 1024) continue;
            if (pass == 1 && size_t(st_n) * CH * CHUNK_BYTES < 96 * 1024) continue;
            if (st_n > 24) st_n = 24;
            if (const char* e = getenv("TT_SCAN_STAGES")) {
                const int want = atoi(e);
                if (want >= 2 && want < st_n) st_n = want;
            }
            *spare_out = spares[i];
            *stages_out = st_n;
            *smem_out = fixed + size_t(st_n) * per_stage;
            return true;
        }
    }
    return false;
}

}  // namespace tc2

template <int N, int CH>
static int launch2(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                   int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx, float* out_thresh,
                   int n_lists, cudaStream_t st) {
    int spare = 0, stages = 0;
    size_t smem = 0;
    if (n_lists % 2 != 0 || !tc2::plan<N, CH>(dim, kprime, &spare, &stages, &smem)) {
        set_error("scan_tc2: dim=%d kprime=%d n_lists=%d unsupported", dim, kprime, n_lists);
        return TT_ERR_UNSUPPORTED;
    }
    tc2::Params p;
    p.inv_norm = inv_norm;
    p.n_rows = n_rows;
    p.id_base = id_base;
    p.n_super = int((n_rows + 255) / 256);
    p.n_chunks = dim / tc::CHUNK_COLS;
    p.stages = stages;
    p.n_q = n_q;
    p.kprime = kprime;
    p.cap = kprime + spare;
    p.out_ids = out_ids;
    p.out_approx = out_approx;
    p.out_thresh = out_thresh;

    CUtensorMap map_c, map_q;
    int rc = tc::make_map(&map_c, corpus, n_rows, dim, stride, tc::TILE_ROWS, CH);
    if (rc) return rc;
    rc = tc::make_map(&map_q, q_hi, n_q, dim, dim, N / 2, 1);
    if (rc) return rc;

    auto kern = tc2::scan_tc2_kernel<N, CH>;
    TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
    for (int q0 = 0; q0 < n_q; q0 += N) {
        p.q0 = q0;
        p.nq_here = (n_q - q0 < N) ? n_q - q0 : N;
        kern<<<n_lists, tc2::THREADS2, smem, st>>>(map_c, map_q, p);
        TT_LAUNCH_OK("scan_tc2_kernel");
    }
    return TT_OK;
}

bool scan_tc2_supported(int dim, int kprime, int n_lists) {
    int sp, sg;
    size_t sm;
    return n_lists % 2 == 0 && dim % 128 == 0 && tc2::plan<64, 2>(dim, kprime, &sp, &sg, &sm);
}

// hi-only scan of n_q queries in passes of 64 with CTA pairs
int scan_tc2_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                    const void* q_hi, int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx,
                    float* out_thresh, int n_lists, cudaStream_t st) {
    if (const char* e = getenv("TT_SCAN_CH")) {  // tuning knob
        if (atoi(e) == 1)
            return launch2<64, 1>(corpus, n_rows, dim, stride, inv_norm, q_hi, n_q, kprime, id_base, out_ids, out_approx,
                                  out_thresh, n_lists, st);
    }
    return launch2<64, 2>(corpus, n_rows, dim, stride, inv_norm, q_hi, n_q, kprime, id_base, out_ids, out_approx,
                          out_thresh, n_lists, st);
}

}  // namespace tt
