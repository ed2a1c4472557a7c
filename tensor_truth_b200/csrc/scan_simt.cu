// CUDA-core streaming scan kernels.
//
//  * scan_simt_kernel<.., float, false>  -- stage-1 shortlist in fp32 (any dim % 8 == 0); the
//    bring-up / odd-shape variant of the tcgen05 scan (scan_tc.cu), same outputs.
//  * scan_simt_kernel<.., double, true>  -- exact fp64 scoring of every row: certificate-failure
//    fallback and the on-GPU secondary oracle (tt_scan_exact_f64).
//
// Replaces the vector-store query issued through `index.as_retriever(similarity_top_k=k)` at
// /root/reference/src/tensortruth/rag_engine.py:639 (ChromaVectorStore.query -> collection.query).
//
// Layout: one warp per corpus row (32 lanes x 16 B coalesced, rows interleaved over all warps of a
// persistent grid of one CTA per SM), two rows in flight per warp; queries staged in shared memory;
// a register-resident warp top-K' list per query; CTA-level bitonic merge at the end.
#include "tt_common.cuh"

namespace tt {

constexpr int SIMT_THREADS = 512;
constexpr int SIMT_WARPS = SIMT_THREADS / 32;

// Row access is split in two: `fetch` only issues the 16-byte global loads of one 8-element chunk, `unpack` turns them
// into floats.  The scan loop fetches several chunks of both rows before it unpacks the first one, so that every warp
// keeps a few KB in flight (ncu had the exact scan latency-bound -- long-scoreboard stalls, DRAM at 34 % -- with one chunk
// per row in flight).
template <typename CT>
struct RowLoader;
template <>
struct RowLoader<__nv_bfloat16> {
    struct Raw { uint4 v; };
    static constexpr int DEPTH = 4;  // chunks per row fetched ahead: 2 rows x 4 x 512 B = 4 KB per warp
    __device__ static __forceinline__ Raw fetch(const __nv_bfloat16* row, int chunk) {
        return Raw{ldg_stream_u4(reinterpret_cast<const uint4*>(row) + chunk)};
    }
    __device__ static __forceinline__ void unpack(const Raw& r, float* f) { unpack_bf16x8(r.v, f); }
};
template <>
struct RowLoader<float> {
    struct Raw { uint4 a, b; };
    static constexpr int DEPTH = 2;  // 2 rows x 2 x 1 KB = 4 KB per warp
    __device__ static __forceinline__ Raw fetch(const float* row, int chunk) {
        return Raw{ldg_stream_u4(reinterpret_cast<const uint4*>(row) + 2 * chunk),
                   ldg_stream_u4(reinterpret_cast<const uint4*>(row) + 2 * chunk + 1)};
    }
    __device__ static __forceinline__ void unpack(const Raw& r, float* f) {
        f[0] = __uint_as_float(r.a.x); f[1] = __uint_as_float(r.a.y); f[2] = __uint_as_float(r.a.z); f[3] = __uint_as_float(r.a.w);
        f[4] = __uint_as_float(r.b.x); f[5] = __uint_as_float(r.b.y); f[6] = __uint_as_float(r.b.z); f[7] = __uint_as_float(r.b.w);
    }
};

// Query staging layout in shared memory, per query: [planes][chunks][VE] with VE = 16 B / sizeof(Acc),
// so that a warp's 16-byte loads (lane = chunk) are bank-conflict free.
template <typename Acc>
struct QLayout {
    static constexpr int VE = 16 / sizeof(Acc);
    static constexpr int PLANES = 8 / VE;
    __device__ static __forceinline__ int index(int d, int chunks) {
        int c = d >> 3, e = d & 7;
        return ((e / VE) * chunks + c) * VE + (e % VE);
    }
    __device__ static __forceinline__ void load8(const Acc* q, int c, int chunks, Acc* out) {
#pragma unroll
        for (int p = 0; p < PLANES; ++p) {
            const uint4 v = *reinterpret_cast<const uint4*>(q + (p * chunks + c) * VE);
            *reinterpret_cast<uint4*>(out + p * VE) = v;
        }
    }
};

__device__ __forceinline__ float exact_key(double dot, double qq, double nn, int mode) {
    if (mode == TT_SCORE_COSINE) {
        double den = sqrt(qq) * sqrt(nn);
        return den > 0.0 ? (float)(dot / den) : 0.0f;
    }
    return -(float)(qq + nn - 2.0 * dot);
}

// Dynamic smem: Acc q_s[QT][dim]  |  (after the scan) uint64 merge[SIMT_WARPS*KP]
template <int E, int QT, typename CT, typename Acc, bool EXACT>
__global__ void __launch_bounds__(SIMT_THREADS, 1)
scan_simt_kernel(const CT* __restrict__ corpus, int64_t n_rows, int dim, int64_t stride,
                 const float* __restrict__ inv_norm,
                 const __nv_bfloat16* __restrict__ q_hi, const __nv_bfloat16* __restrict__ q_lo,  // approx
                 const float* __restrict__ q_f32,                                                   // exact
                 int n_q, int q0, int64_t id_base, int mode,
                 int64_t* __restrict__ out_ids, float* __restrict__ out_approx, float* __restrict__ out_thresh,
                 uint64_t* __restrict__ out_packed /* exact: [n_q, gridDim.x*KP] */) {
    constexpr int KP = 32 * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Acc* q_s = reinterpret_cast<Acc*>(smem_raw);
    __shared__ double qq_s[QT];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nq_here = min(QT, n_q - q0);

    for (int i = threadIdx.x; i < QT * dim; i += SIMT_THREADS) {
        int qi = i / dim, d = i - qi * dim;
        Acc v = Acc(0);
        if (qi < nq_here) {
            size_t o = size_t(q0 + qi) * dim + d;
            if (EXACT) v = Acc(q_f32[o]);
            else v = Acc(__bfloat162float(q_hi[o]) + (q_lo ? __bfloat162float(q_lo[o]) : 0.f));
        }
        q_s[qi * dim + QLayout<Acc>::index(d, dim >> 3)] = v;
    }
    __syncthreads();
    if (EXACT && warp < QT) {
        double a = 0.0;
        for (int d = lane; d < dim; d += 32) { double v = double(q_s[warp * dim + QLayout<Acc>::index(d, dim >> 3)]); a += v * v; }
        a = warp_sum_f64(a);
        if (lane == 0) qq_s[warp] = a;
    }
    __syncthreads();

    WarpList<E> lists[QT];
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) lists[qi].init();

    const int chunks = dim >> 3;
    const int64_t n_warps = int64_t(gridDim.x) * SIMT_WARPS;
    const int64_t gw = int64_t(blockIdx.x) * SIMT_WARPS + warp;

    // The rows of a pair are consumed in fetch groups of DEPTH chunks per lane.  The loop is software-pipelined over the
    // sequence of groups (across row pairs): the loads of group i + 1 are issued before group i is unpacked, so a warp
    // has its next 4 KB in flight while it does the arithmetic of the current ones.
    constexpr int DEPTH = RowLoader<CT>::DEPTH;
    using Raw = typename RowLoader<CT>::Raw;
    const int groups = (chunks + 32 * DEPTH - 1) / (32 * DEPTH);
    auto fetch_group = [&](int64_t ra, int g, Raw (&x0)[DEPTH], Raw (&x1)[DEPTH]) {
        const int64_t rb = ra + n_warps;
        const CT* row0 = corpus + ra * stride;
        const CT* row1 = corpus + (rb < n_rows ? rb : ra) * stride;
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            const int c = (g * DEPTH + u) * 32 + lane;
            if (c < chunks) {
                x0[u] = RowLoader<CT>::fetch(row0, c);
                x1[u] = RowLoader<CT>::fetch(row1, c);
            }
        }
    };
    Raw nxt0[DEPTH] = {}, nxt1[DEPTH] = {};
    Acc dot0[QT], dot1[QT];
    Acc nn0 = Acc(0), nn1 = Acc(0);
    int64_t r0 = gw;
    int g = 0;
    if (r0 < n_rows) fetch_group(r0, 0, nxt0, nxt1);
    while (r0 < n_rows) {
        Raw cur0[DEPTH], cur1[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) { cur0[u] = nxt0[u]; cur1[u] = nxt1[u]; }
        int g2 = g + 1;
        int64_t r2 = r0;
        if (g2 == groups) { g2 = 0; r2 += 2 * n_warps; }
        if (r2 < n_rows) fetch_group(r2, g2, nxt0, nxt1);
        if (g == 0) {
            nn0 = Acc(0);
            nn1 = Acc(0);
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) { dot0[qi] = Acc(0); dot1[qi] = Acc(0); }
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            const int c = (g * DEPTH + u) * 32 + lane;
            if (c >= chunks) break;
            float f0[8], f1[8];
            RowLoader<CT>::unpack(cur0[u], f0);
            RowLoader<CT>::unpack(cur1[u], f1);
            Acc a0[8], a1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { a0[e] = Acc(f0[e]); a1[e] = Acc(f1[e]); }
            if (EXACT) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { nn0 += a0[e] * a0[e]; nn1 += a1[e] * a1[e]; }
            }
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                Acc qv[8];
                QLayout<Acc>::load8(q_s + qi * dim, c, chunks, qv);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dot0[qi] += qv[e] * a0[e];
                    dot1[qi] += qv[e] * a1[e];
                }
            }
        }
        if (g == groups - 1) {
            const int64_t r1 = r0 + n_warps;
            const bool has1 = r1 < n_rows;
            // inv_norm doubles as the ROW GATE: a row whose entry is NaN does not exist for this search (metadata
            // filters).  The approximate scan inherits that from the arithmetic (NaN scores are never inserted); the exact
            // scan, which computes its own norms, reads the array only for the gate.
            float in0 = 1.f, in1 = 1.f;
            if (inv_norm) { in0 = inv_norm[r0]; in1 = inv_norm[has1 ? r1 : r0]; }
            const bool ok0 = in0 == in0, ok1 = has1 && in1 == in1;
            if (EXACT) { nn0 = Acc(warp_sum_f64(double(nn0))); nn1 = Acc(warp_sum_f64(double(nn1))); }
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                float k0, k1;
                if (EXACT) {
                    double d0 = warp_sum_f64(double(dot0[qi])), d1 = warp_sum_f64(double(dot1[qi]));
                    k0 = exact_key(d0, qq_s[qi], double(nn0), mode);
                    k1 = exact_key(d1, qq_s[qi], double(nn1), mode);
                } else {
                    k0 = warp_sum_f32(float(dot0[qi])) * in0;
                    k1 = warp_sum_f32(float(dot1[qi])) * in1;
                }
                if (qi < nq_here) {
                    if (ok0) lists[qi].insert(pack_entry(k0, uint32_t(r0)));
                    if (ok1) lists[qi].insert(pack_entry(k1, uint32_t(r1)));
                }
            }
        }
        g = g2;
        r0 = r2;
    }

    // ---- CTA merge: SIMT_WARPS lists of KP -> top KP, per query
    __syncthreads();  // q_s is dead from here on; reuse smem
    uint64_t* merge = reinterpret_cast<uint64_t*>(smem_raw);
    constexpr int NM = SIMT_WARPS * KP;  // power of two
    for (int qi = 0; qi < nq_here; ++qi) {
#pragma unroll
        for (int i = 0; i < E; ++i) merge[warp * KP + i * 32 + lane] = lists[qi].e[i];
        __syncthreads();
        block_bitonic_sort_desc(merge, NM);
        const int q = q0 + qi;
        if (EXACT) {
            for (int i = threadIdx.x; i < KP; i += SIMT_THREADS) {
                uint64_t e = merge[i];
                // re-pack with the GLOBAL id so the select kernel can order across CTAs
                out_packed[(size_t(q) * gridDim.x + blockIdx.x) * KP + i] =
                    e ? pack_entry(entry_key(e), uint32_t(id_base + entry_id(e))) : 0ull;
            }
        } else {
            for (int i = threadIdx.x; i < KP; i += SIMT_THREADS) {
                uint64_t e = merge[i];
                size_t o = (size_t(q) * gridDim.x + blockIdx.x) * KP + i;
                out_ids[o] = e ? int64_t(id_base + entry_id(e)) : int64_t(-1);
                out_approx[o] = e ? entry_key(e) : -INFINITY;
            }
            if (threadIdx.x == 0) {
                uint64_t last = merge[KP - 1];
                out_thresh[size_t(q) * gridDim.x + blockIdx.x] = last ? entry_key(last) : -INFINITY;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host launchers
template <int E, int QT, typename CT, typename Acc, bool EXACT>
static int launch_simt(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                       const void* q_hi, const void* q_lo, const float* q_f32, int n_q, int64_t id_base, int mode,
                       int64_t* out_ids, float* out_approx, float* out_thresh, uint64_t* out_packed,
                       int n_lists, cudaStream_t st) {
    constexpr int KP = 32 * E;
    size_t smem = size_t(QT) * dim * sizeof(Acc);
    size_t merge_bytes = size_t(SIMT_WARPS) * KP * sizeof(uint64_t);
    if (merge_bytes > smem) smem = merge_bytes;
    auto kern = scan_simt_kernel<E, QT, CT, Acc, EXACT>;
    if (smem > 48 * 1024) TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    for (int q0 = 0; q0 < n_q; q0 += QT) {
        kern<<<n_lists, SIMT_THREADS, smem, st>>>(reinterpret_cast<const CT*>(corpus), n_rows, dim, stride, inv_norm,
                                                  reinterpret_cast<const __nv_bfloat16*>(q_hi),
                                                  reinterpret_cast<const __nv_bfloat16*>(q_lo), q_f32, n_q, q0,
                                                  id_base, mode, out_ids, out_approx, out_thresh, out_packed);
        TT_LAUNCH_OK("scan_simt_kernel");
    }
    return TT_OK;
}

int kprime_to_E(int kprime) {
    if (kprime <= 32) return 1;
    if (kprime <= 64) return 2;
    if (kprime <= 128) return 4;
    if (kprime <= 256) return 8;
    return -1;
}

int scan_simt_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                     const void* q_hi, const void* q_lo, int n_q, int kprime, int64_t id_base,
                     int64_t* out_ids, float* out_approx, float* out_thresh, int n_lists, cudaStream_t st) {
    const int E = kprime_to_E(kprime);
    const bool wide = n_q >= 4;
#define TT_SIMT_A(EE, QQ)                                                                                          \
    return launch_simt<EE, QQ, __nv_bfloat16, float, false>(corpus, n_rows, dim, stride, inv_norm, q_hi, q_lo,       \
                                                            nullptr, n_q, id_base, 0, out_ids, out_approx, out_thresh, \
                                                            nullptr, n_lists, st)
    if (E == 1) { if (wide) TT_SIMT_A(1, 4); else TT_SIMT_A(1, 1); }
    if (E == 2) { if (wide) TT_SIMT_A(2, 4); else TT_SIMT_A(2, 1); }
    if (E == 4) TT_SIMT_A(4, 1);
    if (E == 8) TT_SIMT_A(8, 1);
#undef TT_SIMT_A
    set_error("kprime %d unsupported", kprime);
    return TT_ERR_UNSUPPORTED;
}

int scan_simt_exact(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t stride,
                    const float* q_f32, int n_q, int kprime, int64_t id_base, int mode,
                    uint64_t* out_packed, int n_lists, cudaStream_t st, const float* row_gate) {
    const int E = kprime_to_E(kprime);
    // several queries per corpus pass when there are several (a batch of failed certificates, the parity checks): the
    // rows are loaded, widened and squared once for all of them
#define TT_SIMT_X(EE, QQ, CT)                                                                                    \
    return launch_simt<EE, QQ, CT, double, true>(corpus, n_rows, dim, stride, row_gate, nullptr, nullptr, q_f32, n_q, \
                                                 id_base, mode, nullptr, nullptr, nullptr, out_packed, n_lists, st)
    const int qt = n_q >= 3 ? 4 : n_q == 2 ? 2 : 1;
    if (corpus_dtype == TT_DTYPE_BF16) {
        if (E == 1) { if (qt == 4) TT_SIMT_X(1, 4, __nv_bfloat16); if (qt == 2) TT_SIMT_X(1, 2, __nv_bfloat16); TT_SIMT_X(1, 1, __nv_bfloat16); }
        if (E == 2) { if (qt >= 2) TT_SIMT_X(2, 2, __nv_bfloat16); TT_SIMT_X(2, 1, __nv_bfloat16); }
        if (E == 4) TT_SIMT_X(4, 1, __nv_bfloat16);
        if (E == 8) TT_SIMT_X(8, 1, __nv_bfloat16);
    } else {
        if (E == 1) { if (qt == 4) TT_SIMT_X(1, 4, float); if (qt == 2) TT_SIMT_X(1, 2, float); TT_SIMT_X(1, 1, float); }
        if (E == 2) { if (qt >= 2) TT_SIMT_X(2, 2, float); TT_SIMT_X(2, 1, float); }
        if (E == 4) TT_SIMT_X(4, 1, float);
        if (E == 8) TT_SIMT_X(8, 1, float);
    }
#undef TT_SIMT_X
    set_error("k %d unsupported by the exact scan", kprime);
    return TT_ERR_UNSUPPORTED;
}

}  // namespace tt
