// Stage 2: exact fp64 re-score of the shortlist, exact top-k selection, shard merge, query prep.
//
// With stage 1 this reproduces the exact brute-force target of the vector-store query the
// reference issues at /root/reference/src/tensortruth/rag_engine.py:639 and the scoring
// ChromaVectorStore applies to it (cosine here; exp(-squared-L2) in TT_SCORE_CHROMA_L2_EXP).
#include <string.h>

#include "automerge.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(rescore)

// ------------------------------------------------------------------ query preparation
// one CTA per query: q_hat = q/|q| (fp32), hi = bf16(q_hat), lo = bf16(q_hat - hi)
__global__ void __launch_bounds__(256) prepare_queries_kernel(const float* __restrict__ q, int dim,
                                                              __nv_bfloat16* __restrict__ q_hi,
                                                              __nv_bfloat16* __restrict__ q_lo,
                                                              float* __restrict__ rho_out) {
    __shared__ double red[8];
    __shared__ double red2[8];
    pdl_launch_dependents();  // the scan may become resident now: all it needs from this kernel it waits for (pdl_wait)
    const float* qb = q + size_t(blockIdx.x) * dim;
    double a = 0.0;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) { double v = qb[d]; a += v * v; }
    a = warp_sum_f64(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) tot += red[w];
    const double inv = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
    double res = 0.0;  // |q/|q| - hi|^2: what a hi-only scan does not see of this query
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        const double qd = double(qb[d]) * inv;
        float qh = float(qd);
        __nv_bfloat16 hi = __float2bfloat16_rn(qh);
        size_t o = size_t(blockIdx.x) * dim + d;
        q_hi[o] = hi;
        if (q_lo) q_lo[o] = __float2bfloat16_rn(qh - __bfloat162float(hi));
        const double r = qd - double(__bfloat162float(hi));
        res += r * r;
    }
    if (rho_out) {  // block-uniform
        res = warp_sum_f64(res);
        if ((threadIdx.x & 31) == 0) red2[threadIdx.x >> 5] = res;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t2 = 0.0;
            for (int w = 0; w < int(blockDim.x >> 5); ++w) t2 += red2[w];
            rho_out[blockIdx.x] = float(sqrt(t2) * 1.001 + 2e-6);  // rounded up: it is used as an upper bound
        }
    }
}

// Per-query certificate credit for hi-only scans.  The certificate's error bound has a term for the part of the query
// the tensor cores never saw, |<q/|q| - hi, c>| / |c| <= |q/|q| - hi| =: rho (Cauchy-Schwarz), budgeted at its worst case
// 2^-8 (EPS_HI_ONLY, every element rounded by half a bf16 ulp at the bottom of its binade).  The actual rho of a query
// is known exactly (prepare_queries_kernel) and typically 2.5x smaller.  margin = s_k - max thresh is compared with the
// nominal eps everywhere (host, peers), so the slack eps_hi_only - rho is handed over by LOWERING this query's thresholds
// by it: every later consumer sees margin + credit without knowing.
__global__ void certificate_credit_kernel(float* __restrict__ thresh, int n_q, int n_lists, const float* __restrict__ rho,
                                          float eps_hi_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q * n_lists) return;
    const float credit = eps_hi_only - rho[i / n_lists];
    if (credit > 0.f) thresh[i] -= credit;  // -inf (nothing dropped) and +inf (overflow: never certify) stay what they are
}

// ------------------------------------------------------------------ re-score
// one warp per (query, candidate); fixed summation order (lane-strided, then xor-butterfly).
template <typename CT>
__device__ __forceinline__ uint64_t rescore_one(const CT* __restrict__ corpus, int64_t n_rows, int dim, int64_t stride,
                                                int64_t id_base, const float* __restrict__ qb, int64_t id, int mode,
                                                int lane) {
    const int64_t row = id - id_base;
    if (id < 0 || row < 0 || row >= n_rows) return 0ull;
    const int chunks = dim >> 3;
    const CT* rp = corpus + row * stride;
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
    float key;
    if (mode == TT_SCORE_COSINE) {
        double den = sqrt(qq) * sqrt(nn);
        key = den > 0.0 ? float(dot / den) : 0.0f;
    } else {
        key = -float(qq + nn - 2.0 * dot);
    }
    return pack_entry(key, uint32_t(id));
}

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
    for (int64_t w = warp; w < total; w += n_warps) {
        const int b = int(w / n_cand);
        const uint64_t e = rescore_one<CT>(corpus, n_rows, dim, stride, id_base, q + size_t(b) * dim, cand_ids[w], mode, lane);
        if (lane == 0) packed[w] = e;
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
//
// The epoch of an exchange -- and with it the slot of the receive ring it uses -- either comes from the host
// (tt_exchange_t.epoch) or lives in device memory (epoch_dev: "pushes completed on this lane so far").  The second
// form makes a launch sequence replayable as a CUDA graph: nothing in the kernel parameters changes from step to step.
struct Xchg {
    int push_world;  // > 0: push this launch's output to that many peers
    int wait_world;  // > 0: wait for that many sources before reading the input lists
    int rank;
    unsigned epoch;               // host-assigned epoch (epoch_dev == nullptr)
    unsigned* epoch_dev;          // device-resident epoch counter of this lane, or nullptr
    unsigned n_slots;             // ring of receive regions / flag arrays; slot = epoch % n_slots
    unsigned long long slot_stride, flag_slot_stride;  // bytes between regions, uint32 elements between flag arrays
    unsigned long long rec_stride, ids_off;       // layout of a receive region: [source rank][keys | ids | margins]
    unsigned long long margins_off;               // 0: margins are not exchanged
    unsigned long long recv[TT_MAX_PEERS];        // peer p: base of its receive region (slot 0)
    unsigned long long flags[TT_MAX_PEERS];       // peer p: its flag array (slot 0; element [rank] is ours)
    unsigned* ticket;                             // local counter, zero between launches
};

// what one launch resolves the descriptor to: its epoch and the byte / element offsets of its slot
struct XNow {
    unsigned epoch;
    unsigned long long roff;  // bytes into every receive region
    unsigned long long foff;  // uint32 elements into every flag array
};

__device__ __forceinline__ XNow xchg_now(const Xchg& x, bool pushing) {
    XNow n;
    n.epoch = x.epoch;
    // a pushing launch opens epoch (completed + 1); the waiting launch that follows it in stream order sees it completed
    if (x.epoch_dev) n.epoch = *reinterpret_cast<volatile unsigned*>(x.epoch_dev) + (pushing ? 1u : 0u);
    const unsigned slot = x.n_slots > 1u ? n.epoch % x.n_slots : 0u;
    n.roff = slot * x.slot_stride;
    n.foff = slot * x.flag_slot_stride;
    return n;
}

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

// Every thread of every one of the launch's `n_blocks` pushing blocks calls this after its peer stores: the last block
// to arrive raises our flag on every peer and closes the epoch.
__device__ __forceinline__ void xchg_publish(const Xchg& x, const XNow& now, unsigned n_blocks) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(x.ticket, 1u);
        if (tk == n_blocks - 1u) {  // last block: every block's stores are fenced; raise our flag on every peer
            *x.ticket = 0u;
            if (x.epoch_dev) *x.epoch_dev = now.epoch;
            __threadfence_system();
            for (int p = 0; p < x.push_world; ++p)
                st_release_sys(reinterpret_cast<unsigned*>(x.flags[p]) + now.foff + x.rank, now.epoch);
        }
    }
}

// Threads [0, wait_world) each spin on one source's flag.  Bounded: a source that never shows up is reported as
// TT_STATUS_EXCHANGE_TIMEOUT (its rank in bits 8..15) and the kernel goes on with what it has -- no trap.
__device__ __forceinline__ void xchg_spin(const unsigned* flags, int world, unsigned epoch) {
    if (int(threadIdx.x) < world) {
        long long t0 = 0;
        for (unsigned spins = 0;; ++spins) {
            if (int(ld_acquire_sys(flags + threadIdx.x) - epoch) >= 0) break;
            if ((spins & 0xfffu) == 0xfffu) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4 * status_timeout_cycles()) {
                    status_report(TT_STATUS_EXCHANGE_TIMEOUT | (unsigned(threadIdx.x) << 8));
                    break;
                }
            }
        }
    }
}

// all threads of a block call this before reading gathered lists
__device__ __forceinline__ void xchg_wait(const Xchg& x, const XNow& now) {
    xchg_spin(reinterpret_cast<const unsigned*>(x.flags[x.rank]) + now.foff, x.wait_world, now.epoch);
    __syncthreads();
}

__device__ __forceinline__ void xchg_store(const Xchg& x, const XNow& now, size_t o, float key, int64_t id) {
    for (int p = 0; p < x.push_world; ++p) {
        const unsigned long long rec = x.recv[p] + now.roff + (unsigned long long)x.rank * x.rec_stride;
        reinterpret_cast<float*>(rec)[o] = key;
        reinterpret_cast<int64_t*>(rec + x.ids_off)[o] = id;
    }
}

__device__ __forceinline__ void xchg_store_margin(const Xchg& x, const XNow& now, int b, float margin) {
    if (!x.margins_off) return;
    for (int p = 0; p < x.push_world; ++p) {
        const unsigned long long rec = x.recv[p] + now.roff + (unsigned long long)x.rank * x.rec_stride;
        reinterpret_cast<float*>(rec + x.margins_off)[b] = margin;
    }
}

// ------------------------------------------------------------------ select: one CTA per query
constexpr int SEL_THREADS = 1024;

// everything a selecting block needs besides its input; shared by the kernels below
struct SelArgs {
    const uint64_t* packed;  // [n_q, n_in], or nullptr: (in_keys, in_ids) lists
    int n_in;
    const float* in_keys;    // [n_lists][n_q, k_in] with list strides (the exchange layout)
    const int64_t* in_ids;
    int n_lists;
    int64_t keys_stride, ids_stride;
    int n_q, k_in, k, mode;
    int chunk;               // select_body only: entries per shared-memory sort (pow2, >= 2k)
    const float* thresh;     // [n_q, n_thresh] or nullptr
    int n_thresh;
    float* out_keys;
    float* out_scores;
    int64_t* out_ids;
    float* out_margin;
    float* out_all_margins;  // waiting launch: [wait_world, n_q] margins every source pushed, or nullptr
    Xchg x;
    L2Cert cert;
    AmArgs am;               // am.out_len != nullptr: auto-merge the selected list in the same block
};

__device__ __forceinline__ uint64_t sel_load(const SelArgs& a, const XNow& now, int b, int j) {
    if (a.packed) return __ldcg(a.packed + size_t(b) * a.n_in + j);
    const int l = j / a.k_in, t = j - l * a.k_in;
    const size_t o = size_t(b) * a.k_in + t;
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(a.in_keys) + now.roff;
    const unsigned char* ib = reinterpret_cast<const unsigned char*>(a.in_ids) + now.roff;
    const int64_t id = reinterpret_cast<const int64_t*>(ib)[size_t(l) * a.ids_stride + o];
    return id >= 0 ? pack_entry(reinterpret_cast<const float*>(kb)[size_t(l) * a.keys_stride + o], uint32_t(id)) : 0ull;
}

// entry of the merged / selected list -> outputs (+ peer stores)
__device__ __forceinline__ void sel_emit(const SelArgs& a, const XNow& now, int b, int i, uint64_t e) {
    const size_t o = size_t(b) * a.k + i;
    const float key = e ? entry_key(e) : -INFINITY;
    const int64_t id = e ? int64_t(entry_id(e)) : int64_t(-1);
    if (a.out_keys) a.out_keys[o] = key;
    if (a.out_scores) a.out_scores[o] = e ? score_of_key(key, a.mode) : -INFINITY;
    if (a.out_ids) a.out_ids[o] = id;
    if (a.x.push_world) xchg_store(a.x, now, o, key, id);
}

// certificate margin of query b (all THREADS threads call; `kth` = k-th selected entry, valid in thread 0)
template <int THREADS>
__device__ __forceinline__ void sel_margin(const SelArgs& a, const XNow& now, int b, uint64_t kth, float* red, double* red64,
                                           bool force_fail = false) {
    if (!a.out_margin) return;
    const int t = threadIdx.x;
    float m = -INFINITY;
    if (a.thresh)
        for (int i = t; i < a.n_thresh; i += THREADS) m = fmaxf(m, a.thresh[size_t(b) * a.n_thresh + i]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((t & 31) == 0) red[t >> 5] = m;
    __syncthreads();
    const bool l2c = a.mode != TT_SCORE_COSINE && a.cert.q != nullptr;
    double qq = 0.0;
    if (l2c) qq = block_sqnorm(a.cert.q + size_t(b) * a.cert.dim, a.cert.dim, red64);
    if (t == 0) {
        for (int w = 1; w < THREADS / 32; ++w) m = fmaxf(m, red[w]);
        float margin;
        if (m == -INFINITY) margin = INFINITY;            // nothing was left out of any shortlist
        else if (!kth) margin = -INFINITY;                // fewer than k candidates although rows were dropped
        else if (a.mode == TT_SCORE_COSINE) margin = entry_key(kth) - m;
        else if (l2c) margin = entry_key(kth) - l2_upper_bound(m + a.cert.eps, qq, a.cert.nlo, a.cert.nhi);  // > 0 proves it
        else margin = -INFINITY;                          // cosine-ordered shortlist, no norm bounds given
        if (force_fail) margin = -INFINITY;               // the selection itself was cut short: never certify it
        a.out_margin[b] = margin;
        if (a.x.push_world) xchg_store_margin(a.x, now, b, margin);
    }
}

// waiting launch: copy the margins every source pushed for query b out of this rank's receive region
__device__ __forceinline__ void sel_gather_margins(const SelArgs& a, const XNow& now, int b) {
    if (!a.out_all_margins || !a.x.wait_world || !a.x.margins_off) return;
    if (int(threadIdx.x) < a.x.wait_world) {
        const unsigned long long rec = a.x.recv[a.x.rank] + now.roff + (unsigned long long)threadIdx.x * a.x.rec_stride;
        a.out_all_margins[size_t(threadIdx.x) * a.n_q + b] = reinterpret_cast<const float*>(rec + a.x.margins_off)[b];
    }
}

// stage 3 on the list the block has just written (global memory written by this block, read back after a barrier)
__device__ __forceinline__ void sel_automerge(const SelArgs& a, int b, AmSmem& S) {
    if (!a.am.out_len) return;
    __syncthreads();
    automerge_block(S, a.out_ids + size_t(b) * a.k, a.out_scores + size_t(b) * a.k, a.k, b, a.am);
}

// General selection: sorts the query's packed candidates in chunks of `chunk` (carrying the running top-k) and emits
// the k best.  `s`: chunk uint64 of shared memory.  `n_pushers`: blocks of this launch that publish (exchange).
__device__ __forceinline__ void select_body(const SelArgs& a, int b, uint64_t* s, unsigned n_pushers) {
    __shared__ float red[32];
    __shared__ double red64[32];
    __shared__ AmSmem am_s;
    const int chunk = a.chunk, k = a.k;
    const XNow now = xchg_now(a.x, a.x.push_world > 0);
    if (a.x.wait_world) xchg_wait(a.x, now);
    const int total = a.packed ? a.n_in : a.n_lists * a.k_in;
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
                const uint64_t e = sel_load(a, now, b, j);
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
            const uint64_t e = sel_load(a, now, b, j);
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
            for (int i = threadIdx.x; i < room; i += SEL_THREADS) s[carried + i] = i < take ? sel_load(a, now, b, base + i) : 0ull;
            __syncthreads();
            block_bitonic_sort_desc(s, chunk);
            carried = min(k, chunk);
            base += take;
            if (take == 0) break;
        }
    }
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) sel_emit(a, now, b, i, i < lim ? s[i] : 0ull);
    sel_margin<SEL_THREADS>(a, now, b, (k - 1 < lim) ? s[k - 1] : 0ull, red, red64);
    sel_gather_margins(a, now, b);
    if (a.x.push_world) xchg_publish(a.x, now, n_pushers);  // after the margin: the flag covers the whole record
    sel_automerge(a, b, am_s);
}

// ------------------------------------------------------------------ select, small k: k rounds of block-wide max
// For k <= SMALL_K and <= EPT entries per thread the k best are pulled out one at a time (register-resident
// entries, warp shuffles, one __syncthreads per round) instead of sorting everything: ~1 us for k = 10.
// Shapes (THREADS x EPT): 1024 x 8 (8192 entries: the stand-alone selection), 512 x 10 (the fused re-score of
// 148 shortlists of 32) and 512 x 2 (the shard merge: world x k entries).  The 512-thread shapes are sized to run
// NEXT TO a resident scan CTA -- 192 threads x 128 registers and all but ~20 KB of the SM's shared memory -- so that
// in a two-lane pipeline the tail of step i overlaps the scan of step i + 1.
constexpr int SMALL_K = 32;

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        uint64_t o = shfl_xor_u64(v, m);
        v = o > v ? o : v;
    }
    return v;
}

template <int THREADS, int EPT>
__device__ __forceinline__ void select_small_body(const SelArgs& a, int b, unsigned n_pushers) {
    static_assert(THREADS >= AM_CAP && THREADS % 32 == 0 && THREADS <= 1024, "THREADS");
    __shared__ uint64_t part[2][32];
    __shared__ uint64_t win[SMALL_K];
    __shared__ float red[32];
    __shared__ double red64[32];
    __shared__ AmSmem am_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, k = a.k;
    const XNow now = xchg_now(a.x, a.x.push_world > 0);
    if (a.x.wait_world) xchg_wait(a.x, now);
    const int total = a.packed ? a.n_in : a.n_lists * a.k_in;
    uint64_t e[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int j = t + i * THREADS;
        e[i] = j < total ? sel_load(a, now, b, j) : 0ull;
    }
    for (int r = 0; r < k; ++r) {
        uint64_t m = e[0];
#pragma unroll
        for (int i = 1; i < EPT; ++i) m = e[i] > m ? e[i] : m;
        m = warp_max_u64(m);
        if (lane == 0) part[r & 1][warp] = m;
        __syncthreads();
        uint64_t w = warp_max_u64(lane < THREADS / 32 ? part[r & 1][lane] : 0ull);  // every warp reduces the partials: no second barrier
        if (t == 0) win[r] = w;
        if (w != 0ull) {
#pragma unroll
            for (int i = 0; i < EPT; ++i)
                if (e[i] == w) e[i] = 0ull;  // entries are unique (the id is part of the word)
        }
    }
    __syncthreads();
    if (t < k) sel_emit(a, now, b, t, win[t]);
    sel_margin<THREADS>(a, now, b, win[k - 1], red, red64);
    sel_gather_margins(a, now, b);
    if (a.x.push_world) xchg_publish(a.x, now, n_pushers);  // after the margin: the flag covers the whole record
    sel_automerge(a, b, am_s);
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const SelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    select_body(a, blockIdx.x, reinterpret_cast<uint64_t*>(smem_raw), gridDim.x);
}

template <int THREADS, int EPT>
__global__ void __launch_bounds__(THREADS) select_small_kernel(const SelArgs a) {
    pdl_wait();  // (a merge launched as a programmatic dependent of the pushing kernel: its local outputs are read below)
    select_small_body<THREADS, EPT>(a, blockIdx.x, gridDim.x);
}
// the 1024 x 8 shape: 40 registers is what lets a block become resident next to a scan CTA (40 K + 24 K of 64 K)
__global__ void __maxnreg__(40) select_small_kernel_1024(const SelArgs a) {
    select_small_body<SEL_THREADS, 8>(a, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------ re-score + select in ONE launch
// grid (gx, n_q): the gx blocks of query b re-score its candidates (one warp per candidate) into `packed`; the block
// that finishes last (per-query ticket in `tickets`, zero between launches) selects the query's top-k from them --
// and pushes it to the peers / auto-merges it, like the selecting kernels above.  One launch instead of two on the
// latency chain of every query.
constexpr int FUSED_SMALL_THREADS = 512, FUSED_SMALL_EPT = 10;

template <typename CT, bool SMALL>
__global__ void __launch_bounds__(SMALL ? FUSED_SMALL_THREADS : SEL_THREADS) rescore_select_kernel(const CT* __restrict__ corpus, int64_t n_rows, int dim,
                                                                     int64_t stride, int64_t id_base,
                                                                     const float* __restrict__ q,
                                                                     const int64_t* __restrict__ cand_ids,
                                                                     uint64_t* __restrict__ packed_out,
                                                                     unsigned* __restrict__ tickets, const SelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned sh_last;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n_cand = a.n_in;
    pdl_launch_dependents();
    pdl_wait();
    const float* qb = q + size_t(b) * dim;
    constexpr int WARPS = (SMALL ? FUSED_SMALL_THREADS : SEL_THREADS) / 32;
    for (int c = int(blockIdx.x) * WARPS + int(threadIdx.x >> 5); c < n_cand; c += int(gridDim.x) * WARPS) {
        const size_t w = size_t(b) * n_cand + c;
        const uint64_t e = rescore_one<CT>(corpus, n_rows, dim, stride, id_base, qb, cand_ids[w], a.mode, lane);
        if (lane == 0) packed_out[w] = e;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(tickets + b, 1u);
        sh_last = (tk == gridDim.x - 1u) ? 1u : 0u;
        if (sh_last) tickets[b] = 0u;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    if (SMALL) select_small_body<FUSED_SMALL_THREADS, FUSED_SMALL_EPT>(a, b, gridDim.y);
    else select_body(a, b, reinterpret_cast<uint64_t*>(smem_raw), gridDim.y);
}

// ------------------------------------------------------------------ stage 2 with a pre-filter: ONE block per query
// The shortlist holds thousands of candidates (148 lists x K'), the answer k of them.  With |approx - exact| <= eps --
// the very bound the certificate rests on -- a candidate whose approximate score lies more than 2 eps below the k-th
// best approximate score a_k cannot reach the exact top-k (its exact score is < a_k - eps <= the k-th exact score).
// So: find a_k (k rounds of a block-wide maximum over register-resident 32-bit keys; ties fall together, which only
// lowers the bound), keep the candidates with approx >= a_k - window, re-score only THOSE in fp64 (a few dozen rows
// instead of 4736: one warp each), sort them, emit.  One block, no second wave, no cross-block hand-over -- the
// latency chain of a batch-1 query shrinks by about ten microseconds and stage 2's HBM traffic by two orders of
// magnitude.  More than S2_CAP survivors (a dense neighbourhood under a wide hi-only window, or near-duplicates) is
// reported as an unproven query (margin = -inf): the caller's repair ladder re-runs it.
constexpr int S2_THREADS = 512;
constexpr int S2_EPT = 10;    // candidates per thread: 5120 per query
constexpr int S2_CAP = 1024;  // survivors per query

struct S2Src {
    const void* corpus;
    int64_t n_rows;
    int dim;
    int64_t stride, id_base;
    const float* q;            // [n_q, dim]
    const int64_t* cand_ids;   // [n_q, n_cand]
    const float* cand_approx;  // [n_q, n_cand]
    float window;              // 2 eps
};

union S2Smem {
    AmSmem am;
    struct {
        uint64_t ex[S2_CAP];
        uint32_t rows[S2_CAP];
    } s;
};

template <typename CT>
__global__ void __launch_bounds__(S2_THREADS) stage2_prefilter_kernel(const S2Src r, const SelArgs a) {
    __shared__ S2Smem sm;
    __shared__ uint32_t part[2][S2_THREADS / 32];
    __shared__ float red[32];
    __shared__ double red64[32];
    __shared__ int sh_n;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n_cand = a.n_in, k = a.k;
    pdl_launch_dependents();  // (row-sharded: the merging kernel may become resident)
    pdl_wait();               // everything below reads what the scan wrote
    const XNow now = xchg_now(a.x, a.x.push_world > 0);
    const int64_t* ids = r.cand_ids + size_t(b) * n_cand;
    const float* apx = r.cand_approx + size_t(b) * n_cand;
    if (t == 0) sh_n = 0;

    // ---- 1. order-preserving 32-bit images of the approximate scores, register-resident (0 = no candidate)
    uint32_t ok[S2_EPT], wk[S2_EPT];
#pragma unroll
    for (int i = 0; i < S2_EPT; ++i) {
        const int j = t + i * S2_THREADS;
        uint32_t v = 0u;
        if (j < n_cand && __ldcg(ids + j) >= 0) v = okey_of(__ldcg(apx + j));
        ok[i] = wk[i] = v;
    }
    // ---- 2. the k-th largest distinct key: a lower bound on the k-th best approximate score
    uint32_t kth = 0u;
    int found = 0;
    for (int rd = 0; rd < k; ++rd) {
        uint32_t m = wk[0];
#pragma unroll
        for (int i = 1; i < S2_EPT; ++i) m = max(m, wk[i]);
        m = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0) part[rd & 1][warp] = m;
        __syncthreads();
        const uint32_t w = __reduce_max_sync(0xffffffffu, lane < S2_THREADS / 32 ? part[rd & 1][lane] : 0u);
        if (w == 0u) break;  // fewer than k distinct scores: everything survives (uniform: every thread sees the same w)
        kth = w;
        ++found;
#pragma unroll
        for (int i = 0; i < S2_EPT; ++i)
            if (wk[i] == w) wk[i] = 0u;
    }
    const float cut = (found == k) ? key_of_okey(kth) - r.window : -INFINITY;
    __syncthreads();
    // ---- 3. survivors
#pragma unroll
    for (int i = 0; i < S2_EPT; ++i) {
        if (ok[i] != 0u && key_of_okey(ok[i]) >= cut) {
            const int slot = atomicAdd(&sh_n, 1);
            if (slot < S2_CAP) sm.s.rows[slot] = uint32_t(__ldcg(ids + t + i * S2_THREADS) - r.id_base);
        }
    }
    __syncthreads();
    const bool overflow = sh_n > S2_CAP;
    const int n_s = min(sh_n, S2_CAP);
    // ---- 4. exact fp64 re-score of the survivors, one warp each
    const float* qb = r.q + size_t(b) * r.dim;
    for (int i = warp; i < n_s; i += S2_THREADS / 32) {
        const uint64_t e = rescore_one<CT>(reinterpret_cast<const CT*>(r.corpus), r.n_rows, r.dim, r.stride, r.id_base, qb,
                                           r.id_base + int64_t(sm.s.rows[i]), a.mode, lane);
        if (lane == 0) sm.s.ex[i] = e;
    }
    int lim = 32;
    while (lim < n_s) lim <<= 1;
    __syncthreads();
    for (int i = n_s + t; i < lim; i += S2_THREADS) sm.s.ex[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(sm.s.ex, lim);
    // ---- 5. emit (+ push to the peers / + auto-merge)
    for (int i = t; i < k; i += S2_THREADS) sel_emit(a, now, b, i, i < n_s ? sm.s.ex[i] : 0ull);
    const uint64_t kth_exact = (k - 1 < n_s) ? sm.s.ex[k - 1] : 0ull;
    sel_margin<S2_THREADS>(a, now, b, kth_exact, red, red64, overflow);
    if (a.x.push_world) xchg_publish(a.x, now, gridDim.x);
    sel_automerge(a, b, sm.am);  // (starts with a barrier: the survivors' storage is dead by then)
}

// ------------------------------------------------------------------ exchange without compute
// one block per peer copies the finished local record to it, then the flags are raised (used when the record was
// repaired on the host side of the certificate check)
__global__ void __launch_bounds__(256) exchange_push_kernel(const unsigned char* __restrict__ rec, unsigned long long nbytes,
                                                            const Xchg x) {
    const XNow now = xchg_now(x, true);
    const int p = blockIdx.x;
    unsigned char* dst = reinterpret_cast<unsigned char*>(x.recv[p] + now.roff + (unsigned long long)x.rank * x.rec_stride);
    for (unsigned long long i = threadIdx.x * 4ull; i < nbytes; i += 256 * 4ull)
        *reinterpret_cast<unsigned*>(dst + i) = *reinterpret_cast<const unsigned*>(rec + i);
    xchg_publish(x, now, gridDim.x);
}

// device-side rendezvous of all ranks: raise our flag on every peer, wait for everybody's (one block, one warp)
__global__ void __launch_bounds__(32) peer_barrier_kernel(const Xchg x) {
    const XNow now = xchg_now(x, true);
    if (int(threadIdx.x) < x.push_world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned*>(x.flags[threadIdx.x]) + now.foff + x.rank, now.epoch);
    }
    xchg_spin(reinterpret_cast<const unsigned*>(x.flags[x.rank]) + now.foff, x.push_world, now.epoch);
    __syncwarp();
    if (threadIdx.x == 0 && x.epoch_dev) *x.epoch_dev = now.epoch;
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
    x->epoch_dev = h->epoch_dev;
    x->n_slots = h->n_slots ? h->n_slots : 1u;
    x->slot_stride = h->slot_stride_bytes;
    x->flag_slot_stride = h->flag_slot_stride;
    TT_CHECK_ARG(x->n_slots == 1u || (h->slot_stride_bytes % 8 == 0 && h->flag_slot_stride >= uint64_t(h->world)),
                 "tt_exchange: slot strides (%llu bytes, %llu flags) for %u slots", (unsigned long long)h->slot_stride_bytes,
                 (unsigned long long)h->flag_slot_stride, x->n_slots);
    x->rec_stride = h->rec_stride_bytes;
    x->ids_off = h->ids_off_bytes;
    x->margins_off = h->margins_off_bytes;
    TT_CHECK_ARG(h->peer_flags[h->rank] != nullptr, "tt_exchange: null local flags");
    x->flags[h->rank] = reinterpret_cast<unsigned long long>(h->peer_flags[h->rank]);
    x->recv[h->rank] = reinterpret_cast<unsigned long long>(h->peer_recv[h->rank]);
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
    if (wait) x->wait_world = h->world;
    return TT_OK;
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

int launch_peer_barrier(const tt_exchange_t* h, cudaStream_t st) {
    Xchg x;
    int rc = fill_xchg(&x, h, true, false);
    if (rc) return rc;
    TT_CHECK_ARG(h->world <= 32, "tt_peer_barrier: world=%d", h->world);
    peer_barrier_kernel<<<1, 32, 0, st>>>(x);
    TT_LAUNCH_OK("peer_barrier_kernel");
    return TT_OK;
}

static void fill_am(AmArgs* am, const tt_automerge_args_t* h) {
    memset(am, 0, sizeof(*am));
    if (!h) return;
    am->parent_of = h->parent_of;
    am->child_count = h->child_count;
    am->prev_id = h->prev_id;
    am->next_id = h->next_id;
    am->n_nodes = h->n_nodes;
    am->ratio_thresh = h->ratio_thresh;
    am->max_rounds = h->max_rounds;
    am->out_ids = h->out_ids;
    am->out_scores = h->out_scores;
    am->out_len = h->out_len;
    am->max_out = h->max_out;
}

constexpr int MAX_CHUNK = 8192;  // 64 KB of shared memory

// Selection launch.  Input: `packed` [n_q, n_in] (after a re-score) or (in_keys, in_ids) lists (shard merge).
// `rs` != nullptr fuses the re-score in front of it (rescore_select_kernel).
struct RescoreSrc {
    const void* corpus;
    int dtype;
    int64_t n_rows;
    int dim;
    int64_t stride, id_base;
    const float* q;
    const int64_t* cand_ids;
    uint64_t* packed;
    unsigned* tickets;
};

int launch_select(const uint64_t* packed, int n_in, const float* in_keys, const int64_t* in_ids, int n_lists,
                  int64_t keys_stride, int64_t ids_stride, int n_q, int k_in, int k, int mode, const float* thresh, int n_thresh, float* out_keys,
                  float* out_scores, int64_t* out_ids, float* out_margin, cudaStream_t st,
                  const tt_exchange_t* xh = nullptr, bool push = false, bool wait = false,
                  const tt_l2_cert_t* l2 = nullptr, const float* q_f32 = nullptr, int dim = 0,
                  const tt_automerge_args_t* amh = nullptr, float* out_all_margins = nullptr, const RescoreSrc* rs = nullptr) {
    SelArgs a;
    memset(&a, 0, sizeof(a));
    a.cert.q = (l2 && q_f32) ? q_f32 : nullptr;
    a.cert.dim = dim;
    a.cert.nlo = l2 ? l2->row_norm_min : 0.f;
    a.cert.nhi = l2 ? l2->row_norm_max : 0.f;
    a.cert.eps = l2 ? l2->eps : 0.f;
    {
        int rc = fill_xchg(&a.x, xh, push, wait);
        if (rc) return rc;
    }
    fill_am(&a.am, amh);
    if (a.am.out_len) {
        TT_CHECK_ARG(out_ids && out_scores && a.am.out_ids && a.am.out_scores && a.am.max_out >= 1 && a.am.max_rounds >= 1,
                     "fused auto-merge: null output or max_out / max_rounds < 1");
        TT_CHECK_ARG(k <= AM_CAP / 2, "fused auto-merge: k=%d exceeds %d", k, AM_CAP / 2);
        TT_CHECK_ARG(a.am.n_nodes == 0 || (a.am.parent_of && a.am.child_count && a.am.prev_id && a.am.next_id),
                     "fused auto-merge: null tree array");
    }
    const int total = packed ? n_in : n_lists * k_in;
    if (keys_stride == 0) keys_stride = int64_t(n_q) * k_in;
    if (ids_stride == 0) ids_stride = int64_t(n_q) * k_in;
    a.packed = packed;
    a.n_in = n_in;
    a.in_keys = in_keys;
    a.in_ids = in_ids;
    a.n_lists = n_lists;
    a.keys_stride = keys_stride;
    a.ids_stride = ids_stride;
    a.n_q = n_q;
    a.k_in = k_in;
    a.k = k;
    a.mode = mode;
    a.thresh = thresh;
    a.n_thresh = n_thresh;
    a.out_keys = out_keys;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.out_margin = out_margin;
    a.out_all_margins = out_all_margins;
    if (n_q == 0) return TT_OK;
    bool small = k <= SMALL_K && total <= 8 * SEL_THREADS;
    if (rs && total > FUSED_SMALL_THREADS * FUSED_SMALL_EPT) small = false;  // the fused small shape holds 5120 entries
    size_t smem = 0;
    if (!small) {
        TT_CHECK_ARG(k >= 1 && k <= MAX_CHUNK / 2, "k=%d out of range [1, %d]", k, MAX_CHUNK / 2);
        int chunk = pow2_at_least(total > k ? total : k);
        if (chunk > MAX_CHUNK) chunk = MAX_CHUNK;
        if (chunk < 2 * k) chunk = pow2_at_least(2 * k);
        a.chunk = chunk;
        smem = size_t(chunk) * sizeof(uint64_t);
    }
    if (rs) {
        TT_CHECK_ARG(n_q <= 65535, "fused re-score: n_q=%d exceeds the grid's y extent", n_q);
        const int threads = small ? FUSED_SMALL_THREADS : SEL_THREADS;
        int gx = (n_in + threads / 32 - 1) / (threads / 32);
        if (gx < 1) gx = 1;
        const int cap = 2 * (sm_count(current_device()) > 0 ? sm_count(current_device()) : 148);
        if (gx > cap) gx = cap;
        const dim3 grid(gx, n_q);
        const bool bf16 = rs->dtype == TT_DTYPE_BF16;
#define TT_RS(CT, SM)                                                                                                        \
    do {                                                                                                                     \
        auto kern = rescore_select_kernel<CT, SM>;                                                                           \
        if (!(SM)) TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_CHUNK * int(sizeof(uint64_t)))); \
        TT_CUDA_OK(launch_kernel(kern, grid, dim3(threads), smem, st, true, reinterpret_cast<const CT*>(rs->corpus), rs->n_rows, \
                                 rs->dim, rs->stride, rs->id_base, rs->q, rs->cand_ids, rs->packed, rs->tickets, a));         \
    } while (0)
        if (bf16 && small) TT_RS(__nv_bfloat16, true);
        else if (bf16) TT_RS(__nv_bfloat16, false);
        else if (small) TT_RS(float, true);
        else TT_RS(float, false);
#undef TT_RS
        TT_LAUNCH_OK("rescore_select_kernel");
        return TT_OK;
    }
    if (small) {
        const bool dep = a.x.wait_world > 0;  // the shard merge follows the pushing kernel in its stream
        if (!packed && total <= 512) TT_CUDA_OK(launch_kernel(select_small_kernel<512, 1>, dim3(n_q), dim3(512), 0, st, dep, a));
        else if (!packed && total <= 1024) TT_CUDA_OK(launch_kernel(select_small_kernel<512, 2>, dim3(n_q), dim3(512), 0, st, dep, a));
        else select_small_kernel_1024<<<n_q, SEL_THREADS, 0, st>>>(a);
        TT_LAUNCH_OK("select_small_kernel");
        return TT_OK;
    }
    // per device and cheap: set it on every call (a process may hold indexes on several GPUs)
    TT_CUDA_OK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_CHUNK * int(sizeof(uint64_t))));
    select_kernel<<<n_q, SEL_THREADS, smem, st>>>(a);
    TT_LAUNCH_OK("select_kernel");
    return TT_OK;
}

// re-score + select (+ push / + auto-merge) in one launch; ws = [packed n_q*n_cand u64 | tickets n_q u32]
int launch_stage2_prefilter(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                            const float* q, int n_q, const int64_t* cand_ids, const float* cand_approx, int n_cand, float window,
                            const float* thresh, int n_thresh, int k, int mode, float* out_keys, float* out_scores,
                            int64_t* out_ids, float* out_margin, const tt_exchange_t* xh, const tt_automerge_args_t* amh,
                            cudaStream_t st) {
    SelArgs a;
    memset(&a, 0, sizeof(a));
    {
        int rc = fill_xchg(&a.x, xh, xh != nullptr, false);
        if (rc) return rc;
    }
    fill_am(&a.am, amh);
    if (a.am.out_len) {
        TT_CHECK_ARG(out_ids && out_scores && a.am.out_ids && a.am.out_scores && a.am.max_out >= 1 && a.am.max_rounds >= 1,
                     "fused auto-merge: null output or max_out / max_rounds < 1");
        TT_CHECK_ARG(a.am.n_nodes == 0 || (a.am.parent_of && a.am.child_count && a.am.prev_id && a.am.next_id),
                     "fused auto-merge: null tree array");
    }
    a.n_in = n_cand;
    a.n_q = n_q;
    a.k = k;
    a.mode = mode;
    a.thresh = thresh;
    a.n_thresh = n_thresh;
    a.out_keys = out_keys;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.out_margin = out_margin;
    S2Src r{corpus, n_rows, dim, stride, id_base, q, cand_ids, cand_approx, window};
    if (dtype == TT_DTYPE_BF16) TT_CUDA_OK(launch_kernel(stage2_prefilter_kernel<__nv_bfloat16>, dim3(n_q), dim3(S2_THREADS), 0, st, true, r, a));
    else TT_CUDA_OK(launch_kernel(stage2_prefilter_kernel<float>, dim3(n_q), dim3(S2_THREADS), 0, st, true, r, a));
    TT_LAUNCH_OK("stage2_prefilter_kernel");
    return TT_OK;
}

bool stage2_prefilter_supported(int n_cand, int k, int mode) {
    return mode == TT_SCORE_COSINE && k <= SMALL_K && n_cand <= S2_THREADS * S2_EPT;
}

int launch_rescore_select(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                          const float* q, int n_q, const int64_t* cand_ids, int n_cand, const float* thresh, int n_thresh,
                          int k, int mode, float* out_keys, float* out_scores, int64_t* out_ids, float* out_margin,
                          uint64_t* packed, unsigned* tickets, const tt_exchange_t* xh, const tt_l2_cert_t* l2,
                          const tt_automerge_args_t* amh, cudaStream_t st) {
    RescoreSrc rs{corpus, dtype, n_rows, dim, stride, id_base, q, cand_ids, packed, tickets};
    return launch_select(packed, n_cand, nullptr, nullptr, 0, 0, 0, n_q, 0, k, mode, thresh, n_thresh, out_keys, out_scores,
                         out_ids, out_margin, st, xh, xh != nullptr, false, l2, q, dim, amh, nullptr, &rs);
}

int launch_prepare_queries(const float* q, int n_q, int dim, void* q_hi, void* q_lo, float* rho, cudaStream_t st) {
    if (n_q == 0) return TT_OK;
    prepare_queries_kernel<<<n_q, 256, 0, st>>>(q, dim, reinterpret_cast<__nv_bfloat16*>(q_hi),
                                                reinterpret_cast<__nv_bfloat16*>(q_lo), rho);
    TT_LAUNCH_OK("prepare_queries_kernel");
    return TT_OK;
}

int launch_certificate_credit(float* thresh, int n_q, int n_lists, const float* rho, float eps_hi_only, cudaStream_t st) {
    const int64_t n = int64_t(n_q) * n_lists;
    if (n == 0) return TT_OK;
    certificate_credit_kernel<<<int((n + 255) / 256), 256, 0, st>>>(thresh, n_q, n_lists, rho, eps_hi_only);
    TT_LAUNCH_OK("certificate_credit_kernel");
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
