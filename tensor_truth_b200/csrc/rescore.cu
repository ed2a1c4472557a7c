// Stage 2: exact fp64 re-score of the shortlist, exact top-k selection, shard merge, query prep.
//
// With stage 1 this reproduces the exact brute-force target of the vector-store query the
// reference issues at /root/reference/src/tensortruth/rag_engine.py:639 and the scoring
// ChromaVectorStore applies to it (cosine here; exp(-squared-L2) in TT_SCORE_CHROMA_L2_EXP).
#include <string.h>

#include "tt_common.cuh"

namespace tt {

// ------------------------------------------------------------------ query preparation
// one CTA per query: q_hat = q/|q| (fp32), hi = bf16(q_hat), lo = bf16(q_hat - hi)
__global__ void __launch_bounds__(256) prepare_queries_kernel(const float* __restrict__ q, int dim,
                                                              __nv_bfloat16* __restrict__ q_hi,
                                                              __nv_bfloat16* __restrict__ q_lo) {
    __shared__ double red[8];
    const float* qb = q + size_t(blockIdx.x) * dim;
    double a = 0.0;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) { double v = qb[d]; a += v * v; }
    a = warp_sum_f64(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += red[w];
    const double inv = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float qh = float(double(qb[d]) * inv);
        __nv_bfloat16 hi = __float2bfloat16_rn(qh);
        size_t o = size_t(blockIdx.x) * dim + d;
        q_hi[o] = hi;
        if (q_lo) q_lo[o] = __float2bfloat16_rn(qh - __bfloat162float(hi));
    }
}

// ------------------------------------------------------------------ re-score
// one warp per (query, candidate); fixed summation order (lane-strided, then xor-butterfly).
template <typename CT>
__global__ void __launch_bounds__(256) rescore_kernel(const CT* __restrict__ corpus, int64_t n_rows, int dim,
                                                      int64_t stride, int64_t id_base,
                                                      const float* __restrict__ q, int n_q,
                                                      const int64_t* __restrict__ cand_ids, int n_cand, int mode,
                                                      uint64_t* __restrict__ packed /* [n_q, n_cand] */) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = int64_t(n_q) * n_cand;
    const int chunks = dim >> 3;
    for (int64_t w = warp; w < total; w += n_warps) {
        const int b = int(w / n_cand);
        const int64_t id = cand_ids[w];
        const int64_t row = id - id_base;
        if (id < 0 || row < 0 || row >= n_rows) {
            if (lane == 0) packed[w] = 0ull;
            continue;
        }
        const CT* rp = corpus + row * stride;
        const float* qb = q + size_t(b) * dim;
        double dot = 0.0, nn = 0.0, qq = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            float f[8];
            if (sizeof(CT) == 2) {
                uint4 v = *(reinterpret_cast<const uint4*>(rp) + c);
                unpack_bf16x8(v, f);
            } else {
                float4 a = *(reinterpret_cast<const float4*>(rp) + 2 * c);
                float4 bb = *(reinterpret_cast<const float4*>(rp) + 2 * c + 1);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = bb.x; f[5] = bb.y; f[6] = bb.z; f[7] = bb.w;
            }
            float4 q0 = *(reinterpret_cast<const float4*>(qb) + 2 * c);
            float4 q1 = *(reinterpret_cast<const float4*>(qb) + 2 * c + 1);
            const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                double cd = double(f[e]), qd = double(qv[e]);
                dot += qd * cd;
                nn += cd * cd;
                qq += qd * qd;
            }
        }
        dot = warp_sum_f64(dot);
        nn = warp_sum_f64(nn);
        qq = warp_sum_f64(qq);
        if (lane == 0) {
            float key;
            if (mode == TT_SCORE_COSINE) {
                double den = sqrt(qq) * sqrt(nn);
                key = den > 0.0 ? float(dot / den) : 0.0f;
            } else {
                key = -float(qq + nn - 2.0 * dot);
            }
            packed[w] = pack_entry(key, uint32_t(id));
        }
    }
}

__device__ __forceinline__ float score_of_key(float key, int mode) {
    return mode == TT_SCORE_COSINE ? key : float(exp(double(key)));  // key = -d
}

// ------------------------------------------------------------------ peer exchange (row-sharded corpus)
// The per-rank exact top-k is PUSHED straight from the selecting kernel into every rank's receive region over
// NVLink peer mappings (symmetric memory), followed by a system-scope release of a per-source flag; the merging
// kernel on each rank spins on its own flags (acquire) before it reads.  No NCCL call, no extra kernel: the
// collective is fused into the two kernels on either side of it.
struct Xchg {
    int push_world;  // > 0: push this launch's output to that many peers
    int wait_world;  // > 0: wait for that many sources before reading the input lists
    int rank;
    unsigned epoch;
    unsigned long long rec_stride, ids_off;       // layout of a receive region: [source rank][keys | ids | margins]
    unsigned long long margins_off;               // 0: margins are not exchanged
    unsigned long long recv[TT_MAX_PEERS];        // peer p: base of its receive region for this slot
    unsigned long long flags[TT_MAX_PEERS];       // peer p: its flag array for this slot (element [rank] is ours)
    unsigned* ticket;                             // local counter, zero between launches
    const unsigned* wait_flags;                   // local flag array for this slot
};

// Certificate inputs for TT_SCORE_CHROMA_L2_EXP: the shortlist is ordered by cosine, so a dropped row r is only
// known to have cos(r) <= t (t = threshold + eps).  With every row norm in [nlo, nhi] that bounds its squared-L2
// key from above:  -(|q|^2 + |c|^2 - 2 cos |q||c|)  <=  U(t),  U maximised over |c| in [nlo, nhi] (a concave
// parabola in |c| with vertex at t|q|).  The top-k is proven exact iff its k-th key exceeds U.
struct L2Cert {
    const float* q;  // [n_q, dim]; NULL = no L2 certificate
    int dim;
    float nlo, nhi, eps;
};

__device__ __forceinline__ float l2_upper_bound(float t, double qq, float nlo, float nhi) {
    const double qn = sqrt(qq);
    double n = double(t) * qn;  // unconstrained maximiser of -(n^2 - 2 t qn n)
    n = n < double(nlo) ? double(nlo) : (n > double(nhi) ? double(nhi) : n);
    return float(-(qq + n * n - 2.0 * double(t) * qn * n));
}

// block-wide |q_b|^2 in fp64 (all threads call; result valid in thread 0)
__device__ __forceinline__ double block_sqnorm(const float* q, int dim, double* red /* [32] shared */) {
    double a = 0.0;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) { const double v = q[d]; a += v * v; }
    a = warp_sum_f64(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += red[w];
    return tot;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all threads of all blocks of the launch call this after their peer stores
__device__ __forceinline__ void xchg_publish(const Xchg& x) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(x.ticket, 1u);
        if (tk == gridDim.x - 1) {  // last block: every block's stores are fenced; raise our flag on every peer
            *x.ticket = 0u;
            __threadfence_system();
            for (int p = 0; p < x.push_world; ++p)
                st_release_sys(reinterpret_cast<unsigned*>(x.flags[p]) + x.rank, x.epoch);
        }
    }
}

// all threads of a block call this before reading gathered lists
__device__ __forceinline__ void xchg_wait(const Xchg& x) {
    if (int(threadIdx.x) < x.wait_world) {
        long long t0 = 0;
        for (unsigned spins = 0;; ++spins) {
            if (int(ld_acquire_sys(x.wait_flags + threadIdx.x) - x.epoch) >= 0) break;
            if ((spins & 0xfffu) == 0xfffu) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 20000000000ll) {
                    printf("tt_b200: peer exchange timed out waiting for rank %d (epoch %u)\n", int(threadIdx.x), x.epoch);
                    __trap();
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void xchg_store(const Xchg& x, size_t o, float key, int64_t id) {
    for (int p = 0; p < x.push_world; ++p) {
        const unsigned long long rec = x.recv[p] + (unsigned long long)x.rank * x.rec_stride;
        reinterpret_cast<float*>(rec)[o] = key;
        reinterpret_cast<int64_t*>(rec + x.ids_off)[o] = id;
    }
}

__device__ __forceinline__ void xchg_store_margin(const Xchg& x, int b, float margin) {
    if (!x.margins_off) return;
    for (int p = 0; p < x.push_world; ++p) {
        const unsigned long long rec = x.recv[p] + (unsigned long long)x.rank * x.rec_stride;
        reinterpret_cast<float*>(rec + x.margins_off)[b] = margin;
    }
}

// ------------------------------------------------------------------ select: one CTA per query
// Sorts the query's packed candidates in chunks of SEL_CHUNK (carrying the running top-k) and
// emits the k best.  Input either `packed` [n_q, n_in] or (keys, ids) laid out
// [n_lists, n_q, k_in] (the all-gather layout) when packed == nullptr.
constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const uint64_t* __restrict__ packed, int n_in,
                                                             const float* __restrict__ in_keys,
                                                             const int64_t* __restrict__ in_ids, int n_lists,
                                                             int64_t keys_stride, int64_t ids_stride,
                                                             int n_q, int k_in, int chunk /* pow2 */, int k,
                                                             int mode, const float* __restrict__ thresh,
                                                             int n_thresh, float* __restrict__ out_keys,
                                                             float* __restrict__ out_scores,
                                                             int64_t* __restrict__ out_ids,
                                                             float* __restrict__ out_margin, const Xchg x,
                                                             const L2Cert cert) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ float red[32];
    __shared__ double red64[32];
    const int b = blockIdx.x;
    if (x.wait_world) xchg_wait(x);
    const int total = packed ? n_in : n_lists * k_in;
    auto load = [&](int j) -> uint64_t {
        if (packed) return packed[size_t(b) * n_in + j];
        const int l = j / k_in, t = j - l * k_in;
        const size_t o = size_t(b) * k_in + t;
        const int64_t id = in_ids[size_t(l) * ids_stride + o];
        return id >= 0 ? pack_entry(in_keys[size_t(l) * keys_stride + o], uint32_t(id)) : 0ull;
    };
    int lim = chunk;  // s[0, lim) is sorted when the selection is done
    if (total > chunk) {
        // More candidates than one sort holds (k = 200 over 148 shortlists of 128): find the k-th largest packed entry
        // with an MSB-first radix select over the input (8 bits per pass, L2-resident re-reads, stops as soon as the
        // digit bin holds exactly the entries still wanted), then sort only the k survivors.
        __shared__ int hist[256];
        __shared__ int sh_bin, sh_want, sh_n;
        uint64_t prefix = 0ull, mask = 0ull;
        int want = k;
        if (threadIdx.x == 0) sh_n = 0;
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            for (int j = threadIdx.x; j < total; j += SEL_THREADS) {
                const uint64_t e = load(j);
                if ((e & mask) == prefix) atomicAdd(&hist[int(e >> shift) & 255], 1);
            }
            __syncthreads();
            if (threadIdx.x < 32) {  // bins from 255 down: lane l owns bins 255-8l .. 248-8l
                const int t = threadIdx.x;
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
                int before = inc - sum;
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
            if (hist[bin] == want) break;
            __syncthreads();
        }
        // distinct entries: exactly k pass; if the k-th largest is an empty slot (fewer than k candidates) all real ones do
        for (int j = threadIdx.x; j < total; j += SEL_THREADS) {
            const uint64_t e = load(j);
            if (e != 0ull && (e & mask) >= prefix) {
                const int slot = atomicAdd(&sh_n, 1);
                if (slot < chunk) s[slot] = e;
            }
        }
        __syncthreads();
        const int n_sel = min(sh_n, chunk);
        lim = 32;
        while (lim < n_sel || lim < k) lim <<= 1;  // <= chunk: chunk is a power of two >= 2k
        for (int i = n_sel + threadIdx.x; i < lim; i += SEL_THREADS) s[i] = 0ull;
        __syncthreads();
        block_bitonic_sort_desc(s, lim);
    } else {
        int carried = 0;  // entries [0, carried) of s hold the running top-k
        for (int base = 0; base < total || base == 0;) {
            const int room = chunk - carried;
            const int take = min(room, total - base);
            for (int i = threadIdx.x; i < room; i += SEL_THREADS) s[carried + i] = i < take ? load(base + i) : 0ull;
            __syncthreads();
            block_bitonic_sort_desc(s, chunk);
            carried = min(k, chunk);
            base += take;
            if (take == 0) break;
        }
    }
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
        const uint64_t e = i < lim ? s[i] : 0ull;
        const size_t o = size_t(b) * k + i;
        const float key = e ? entry_key(e) : -INFINITY;
        if (out_keys) out_keys[o] = key;
        if (out_scores) out_scores[o] = e ? score_of_key(key, mode) : -INFINITY;
        if (out_ids) out_ids[o] = e ? int64_t(entry_id(e)) : int64_t(-1);
        if (x.push_world) xchg_store(x, o, key, e ? int64_t(entry_id(e)) : int64_t(-1));
    }
    if (out_margin) {
        float m = -INFINITY;
        if (thresh)
            for (int i = threadIdx.x; i < n_thresh; i += SEL_THREADS) m = fmaxf(m, thresh[size_t(b) * n_thresh + i]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        const bool l2c = mode != TT_SCORE_COSINE && cert.q != nullptr;
        double qq = 0.0;
        if (l2c) qq = block_sqnorm(cert.q + size_t(b) * cert.dim, cert.dim, red64);
        if (threadIdx.x == 0) {
            for (int w = 1; w < SEL_THREADS / 32; ++w) m = fmaxf(m, red[w]);
            const uint64_t ek = (k - 1 < lim) ? s[k - 1] : 0ull;
            float margin;
            if (m == -INFINITY) margin = INFINITY;            // nothing was left out of any shortlist
            else if (!ek) margin = -INFINITY;                 // fewer than k candidates although rows were dropped
            else if (mode == TT_SCORE_COSINE) margin = entry_key(ek) - m;
            else if (l2c) margin = entry_key(ek) - l2_upper_bound(m + cert.eps, qq, cert.nlo, cert.nhi);  // > 0 proves it
            else margin = -INFINITY;                          // cosine-ordered shortlist, no norm bounds given
            out_margin[b] = margin;
            if (x.push_world) xchg_store_margin(x, b, margin);
        }
    }
    if (x.push_world) xchg_publish(x);  // after the margin: the flag covers the whole record
}

// ------------------------------------------------------------------ select, small k: k rounds of block-wide max
// For k <= SMALL_K and <= 8 entries per thread the k best are pulled out one at a time (register-resident
// entries, warp shuffles, one __syncthreads per round) instead of sorting everything: ~1 us for k = 10.
constexpr int SMALL_K = 32;
constexpr int SMALL_EPT = 8;

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        uint64_t o = shfl_xor_u64(v, m);
        v = o > v ? o : v;
    }
    return v;
}

__global__ void __launch_bounds__(SEL_THREADS) select_small_kernel(
    const uint64_t* __restrict__ packed, int n_in, const float* __restrict__ in_keys, const int64_t* __restrict__ in_ids,
    int n_lists, int64_t keys_stride, int64_t ids_stride, int n_q, int k_in, int k, int mode,
    const float* __restrict__ thresh, int n_thresh, float* __restrict__ out_keys, float* __restrict__ out_scores,
    int64_t* __restrict__ out_ids, float* __restrict__ out_margin, const Xchg x, const L2Cert cert) {
    __shared__ uint64_t part[2][32];
    __shared__ uint64_t win[SMALL_K];
    __shared__ float red[32];
    __shared__ double red64[32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (x.wait_world) xchg_wait(x);
    const int total = packed ? n_in : n_lists * k_in;
    uint64_t e[SMALL_EPT];
#pragma unroll
    for (int i = 0; i < SMALL_EPT; ++i) {
        const int j = t + i * SEL_THREADS;
        uint64_t v = 0ull;
        if (j < total) {
            if (packed) {
                v = packed[size_t(b) * n_in + j];
            } else {
                const int l = j / k_in, c = j - l * k_in;
                const size_t o = size_t(b) * k_in + c;
                const int64_t id = in_ids[size_t(l) * ids_stride + o];
                v = id >= 0 ? pack_entry(in_keys[size_t(l) * keys_stride + o], uint32_t(id)) : 0ull;
            }
        }
        e[i] = v;
    }
    for (int r = 0; r < k; ++r) {
        uint64_t m = e[0];
#pragma unroll
        for (int i = 1; i < SMALL_EPT; ++i) m = e[i] > m ? e[i] : m;
        m = warp_max_u64(m);
        if (lane == 0) part[r & 1][warp] = m;
        __syncthreads();
        uint64_t w = warp_max_u64(part[r & 1][lane]);  // every warp reduces the 32 partials: no second barrier
        if (t == 0) win[r] = w;
        if (w != 0ull) {
#pragma unroll
            for (int i = 0; i < SMALL_EPT; ++i)
                if (e[i] == w) e[i] = 0ull;  // entries are unique (the id is part of the word)
        }
    }
    __syncthreads();
    if (t < k) {
        const uint64_t w = win[t];
        const size_t o = size_t(b) * k + t;
        const float key = w ? entry_key(w) : -INFINITY;
        if (out_keys) out_keys[o] = key;
        if (out_scores) out_scores[o] = w ? score_of_key(key, mode) : -INFINITY;
        if (out_ids) out_ids[o] = w ? int64_t(entry_id(w)) : int64_t(-1);
        if (x.push_world) xchg_store(x, o, key, w ? int64_t(entry_id(w)) : int64_t(-1));
    }
    if (out_margin) {
        float m = -INFINITY;
        if (thresh)
            for (int i = t; i < n_thresh; i += SEL_THREADS) m = fmaxf(m, thresh[size_t(b) * n_thresh + i]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) red[warp] = m;
        __syncthreads();
        const bool l2c = mode != TT_SCORE_COSINE && cert.q != nullptr;
        double qq = 0.0;
        if (l2c) qq = block_sqnorm(cert.q + size_t(b) * cert.dim, cert.dim, red64);
        if (t == 0) {
            for (int w = 1; w < SEL_THREADS / 32; ++w) m = fmaxf(m, red[w]);
            const uint64_t ek = win[k - 1];
            float margin;
            if (m == -INFINITY) margin = INFINITY;
            else if (!ek) margin = -INFINITY;
            else if (mode == TT_SCORE_COSINE) margin = entry_key(ek) - m;
            else if (l2c) margin = entry_key(ek) - l2_upper_bound(m + cert.eps, qq, cert.nlo, cert.nhi);
            else margin = -INFINITY;
            out_margin[b] = margin;
            if (x.push_world) xchg_store_margin(x, b, margin);
        }
    }
    if (x.push_world) xchg_publish(x);  // after the margin: the flag covers the whole record
}

static int pow2_at_least(int n) {
    int p = 32;
    while (p < n) p <<= 1;
    return p;
}

static int fill_xchg(Xchg* x, const tt_exchange_t* h, bool push, bool wait) {
    memset(x, 0, sizeof(*x));
    if (!h) return TT_OK;
    TT_CHECK_ARG(h->world >= 1 && h->world <= TT_MAX_PEERS && h->rank >= 0 && h->rank < h->world,
                 "tt_exchange: world=%d rank=%d", h->world, h->rank);
    x->rank = h->rank;
    x->epoch = h->epoch;
    x->rec_stride = h->rec_stride_bytes;
    x->ids_off = h->ids_off_bytes;
    x->margins_off = h->margins_off_bytes;
    if (push) {
        TT_CHECK_ARG(h->ticket != nullptr, "tt_exchange: null ticket");
        x->push_world = h->world;
        x->ticket = h->ticket;
        for (int p = 0; p < h->world; ++p) {
            TT_CHECK_ARG(h->peer_recv[p] && h->peer_flags[p], "tt_exchange: null peer pointer %d", p);
            x->recv[p] = reinterpret_cast<unsigned long long>(h->peer_recv[p]);
            x->flags[p] = reinterpret_cast<unsigned long long>(h->peer_flags[p]);
        }
    }
    if (wait) {
        TT_CHECK_ARG(h->peer_flags[h->rank] != nullptr, "tt_exchange: null local flags");
        x->wait_world = h->world;
        x->wait_flags = h->peer_flags[h->rank];
    }
    return TT_OK;
}

// one block per peer copies the finished local record to it, then the flags are raised (exchange without compute:
// used when the record was repaired on the host side of the certificate check)
__global__ void __launch_bounds__(256) exchange_push_kernel(const unsigned char* __restrict__ rec, unsigned long long nbytes,
                                                            const Xchg x) {
    const int p = blockIdx.x;
    unsigned char* dst = reinterpret_cast<unsigned char*>(x.recv[p] + (unsigned long long)x.rank * x.rec_stride);
    for (unsigned long long i = threadIdx.x * 4ull; i < nbytes; i += 256 * 4ull)
        *reinterpret_cast<unsigned*>(dst + i) = *reinterpret_cast<const unsigned*>(rec + i);
    xchg_publish(x);
}

int launch_exchange_push(const void* rec, size_t nbytes, const tt_exchange_t* h, cudaStream_t st) {
    Xchg x;
    int rc = fill_xchg(&x, h, true, false);
    if (rc) return rc;
    TT_CHECK_ARG(h && nbytes % 4 == 0 && nbytes <= h->rec_stride_bytes, "tt_exchange_push: nbytes=%zu", nbytes);
    exchange_push_kernel<<<h->world, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(rec), nbytes, x);
    TT_LAUNCH_OK("exchange_push_kernel");
    return TT_OK;
}

int launch_select(const uint64_t* packed, int n_in, const float* in_keys, const int64_t* in_ids, int n_lists,
                  int64_t keys_stride, int64_t ids_stride, int n_q, int k_in, int k, int mode, const float* thresh, int n_thresh, float* out_keys,
                  float* out_scores, int64_t* out_ids, float* out_margin, cudaStream_t st,
                  const tt_exchange_t* xh = nullptr, bool push = false, bool wait = false,
                  const tt_l2_cert_t* l2 = nullptr, const float* q_f32 = nullptr, int dim = 0) {
    L2Cert cert;
    cert.q = (l2 && q_f32) ? q_f32 : nullptr;
    cert.dim = dim;
    cert.nlo = l2 ? l2->row_norm_min : 0.f;
    cert.nhi = l2 ? l2->row_norm_max : 0.f;
    cert.eps = l2 ? l2->eps : 0.f;
    Xchg x;
    {
        int rc = fill_xchg(&x, xh, push, wait);
        if (rc) return rc;
    }
    const int total = packed ? n_in : n_lists * k_in;
    if (keys_stride == 0) keys_stride = int64_t(n_q) * k_in;
    if (ids_stride == 0) ids_stride = int64_t(n_q) * k_in;
    if (k <= SMALL_K && total <= SMALL_EPT * SEL_THREADS) {
        if (n_q == 0) return TT_OK;
        select_small_kernel<<<n_q, SEL_THREADS, 0, st>>>(packed, n_in, in_keys, in_ids, n_lists, keys_stride, ids_stride, n_q,
                                                         k_in, k, mode, thresh, n_thresh, out_keys, out_scores, out_ids,
                                                         out_margin, x, cert);
        TT_LAUNCH_OK("select_small_kernel");
        return TT_OK;
    }
    constexpr int MAX_CHUNK = 8192;  // 64 KB of shared memory
    TT_CHECK_ARG(k >= 1 && k <= MAX_CHUNK / 2, "k=%d out of range [1, %d]", k, MAX_CHUNK / 2);
    int chunk = pow2_at_least(total > k ? total : k);
    if (chunk > MAX_CHUNK) chunk = MAX_CHUNK;
    if (chunk < 2 * k) chunk = pow2_at_least(2 * k);
    const size_t smem = size_t(chunk) * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        TT_CUDA_OK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        MAX_CHUNK * int(sizeof(uint64_t))));
        attr_set = true;
    }
    if (n_q == 0) return TT_OK;
    select_kernel<<<n_q, SEL_THREADS, smem, st>>>(packed, n_in, in_keys, in_ids, n_lists, keys_stride, ids_stride, n_q, k_in, chunk, k, mode,
                                                  thresh, n_thresh, out_keys, out_scores, out_ids, out_margin, x, cert);
    TT_LAUNCH_OK("select_kernel");
    return TT_OK;
}

int launch_prepare_queries(const float* q, int n_q, int dim, void* q_hi, void* q_lo, cudaStream_t st) {
    if (n_q == 0) return TT_OK;
    prepare_queries_kernel<<<n_q, 256, 0, st>>>(q, dim, reinterpret_cast<__nv_bfloat16*>(q_hi),
                                                reinterpret_cast<__nv_bfloat16*>(q_lo));
    TT_LAUNCH_OK("prepare_queries_kernel");
    return TT_OK;
}

int launch_rescore(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                   const float* q, int n_q, const int64_t* cand_ids, int n_cand, int mode, uint64_t* packed,
                   cudaStream_t st) {
    const int64_t total = int64_t(n_q) * n_cand;
    if (total == 0) return TT_OK;
    const int warps_per_block = 8;
    int64_t blocks = (total + warps_per_block - 1) / warps_per_block;
    const int64_t cap = int64_t(sm_count(current_device())) * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == TT_DTYPE_BF16)
        rescore_kernel<__nv_bfloat16><<<int(blocks), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(corpus),
                                                                  n_rows, dim, stride, id_base, q, n_q, cand_ids,
                                                                  n_cand, mode, packed);
    else
        rescore_kernel<float><<<int(blocks), 256, 0, st>>>(reinterpret_cast<const float*>(corpus), n_rows, dim,
                                                          stride, id_base, q, n_q, cand_ids, n_cand, mode, packed);
    TT_LAUNCH_OK("rescore_kernel");
    return TT_OK;
}

}  // namespace tt
