// Packed variable-length self-attention + the classification head of the cross-encoder reranker: the two pieces of the
// stage AFTER the retrieval path that round 1 still borrowed from a library (reference:
// SentenceTransformerRerank.postprocess_nodes, wired at /root/reference/src/tensortruth/services/model_manager.py:333-337
// and run at services/rag_service.py:343-346 on the node list the retriever returned).
//
// attn_varlen_kernel: bidirectional (encoder) attention, head_dim = 64, sequences of up to 512 tokens (XLM-RoBERTa's
// limit), bf16 in / bf16 out, fp32 accumulation.  A work item is (128-row query tile of one packed sequence, head):
//
//   S_j = Q K_j^T   tcgen05.mma  M = 128 (query rows) x N = 128 (keys of kv tile j) x K = 64, operands K-major SW128
//                   straight from TMA (the packed [T, 3H] QKV matrix is its own tensor map; a tile is 128 rows x 128 B).
//                   All (at most four) S tiles of the item stay resident in TMEM: 4 x 128 = 512 columns.
//   P_j = exp2((S_j - m) c)   softmax numerators in registers: thread = query row = TMEM lane; pass A reads the S
//                   tiles for the row maximum m, pass B reads them again for the numerators with the FINAL maximum --
//                   so the accumulator O never needs a correction step and nothing is recomputed.
//   O += P_j V_j    tcgen05.mma  M = 128 x N = 64 (head_dim) x K = 128 (keys): P_j goes back through shared memory as
//                   a K-major SW128 A operand (bf16, double-buffered); V_j is used AS IT LIES in memory -- [keys, 64]
//                   rows of 128 B are an MN-major B operand (instruction descriptor bit 16), so nothing is transposed.
//                   O lives in the TMEM columns of S_0, which pass B has finished with before the first P V is issued.
//
// Persistent CTAs (one per SM) walk the items; K / V tiles stream through a 7-stage TMA ring and Q through a 2-stage
// one, so the loads of the next item are in flight while the softmax warps -- the bottleneck -- work on the current one.
// Ten warps: 0-7 softmax / epilogue (warps w and w + 4 share the TMEM lane quarter w % 4 and split every S tile's 128
// key columns -- and O's 64 -- in halves; row maxima and row sums are combined through shared memory), 8 TMA producer,
// 9 TMEM alloc + MMA issue.
#include "tc_ptx.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(attention)
namespace attn {

using namespace tc;

constexpr int HD = 64;               // head_dim
constexpr int QT = 128;              // query rows per item (MMA M)
constexpr int KT = 128;              // keys per kv tile (MMA N of S, MMA K of P V)
constexpr int MAX_KV_TILES = 4;      // sequences up to 512 tokens
constexpr int TILE_BYTES = QT * 128; // a [128 x 64] bf16 tile: 16 KB
constexpr int P_BYTES = 2 * TILE_BYTES;  // a [128 x 128] bf16 P tile: two 64-column chunks
constexpr int KV_STAGES = 7;
constexpr int Q_STAGES = 2;
constexpr int TMEM_COLS = 512;       // S_j at columns 128 j; O in columns 0..63 (over S_0)
constexpr int SM_THREADS = 256;      // softmax / epilogue threads: two per query row
constexpr int A_THREADS = SM_THREADS + 64;
constexpr int MAX_TILES = 1024;      // query tiles per launch (the tile table in shared memory)

struct Params {
    const int* cu_seqlens;  // [n_seq + 1] token offsets of the packed sequences
    int n_seq;
    int n_heads;
    int hidden;             // n_heads * 64
    float scale_log2e;      // softmax scale * log2(e)
    __nv_bfloat16* out;     // [T, hidden]
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
        : "memory");
}

// MN-major, 128-byte-swizzled B operand: [K rows][64 MN elements = 128 B], 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3ffffu) >> 4);  // start address
    d |= uint64_t(1) << 16;                      // leading byte offset: one 64-element MN block only (unused)
    d |= uint64_t(1024 >> 4) << 32;              // stride byte offset: 8 K-rows x 128 B
    d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                      // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D = fp32, A = B = bf16, A K-major, B K-major or MN-major, M = 128
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (uint32_t(n >> 3) << 17) | (uint32_t(QT >> 4) << 24);
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(A_THREADS, 1)
attn_varlen_kernel(const __grid_constant__ CUtensorMap map_qkv, const Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* q_s = smem;                                   // Q_STAGES x 16 KB
    unsigned char* kv_s = q_s + Q_STAGES * TILE_BYTES;           // KV_STAGES x 16 KB
    unsigned char* p_s = kv_s + KV_STAGES * TILE_BYTES;          // 2 x 32 KB
    int2* tiles = reinterpret_cast<int2*>(p_s + 2 * P_BYTES);    // MAX_TILES x {sequence start, length | tile index << 16}
    float* xch = reinterpret_cast<float*>(tiles + MAX_TILES);    // [2][QT] row maxima, then [2][QT] row sums, of the two column halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 4 * QT);
    uint64_t* kv_full = bars;                      // KV_STAGES
    uint64_t* kv_empty = kv_full + KV_STAGES;      // KV_STAGES
    uint64_t* q_full = kv_empty + KV_STAGES;       // Q_STAGES
    uint64_t* q_empty = q_full + Q_STAGES;         // Q_STAGES
    uint64_t* s_full = q_empty + Q_STAGES;         // MAX_KV_TILES
    uint64_t* p_full = s_full + MAX_KV_TILES;      // 2
    uint64_t* p_empty = p_full + 2;                // 2
    uint64_t* o_full = p_empty + 2;                // 1
    uint64_t* tmem_free = o_full + 1;              // 1
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_free + 1);
    int* n_tiles_s = reinterpret_cast<int*>(tmem_base_s + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- the tile table: query tile -> (sequence start, sequence length, tile index within the sequence); warp 0
    if (warp == 0) {
        int total = 0;
        for (int base = 0; base < p.n_seq; base += 32) {
            const int i = base + lane;
            int beg = 0, len = 0;
            if (i < p.n_seq) {
                beg = __ldg(p.cu_seqlens + i);
                len = __ldg(p.cu_seqlens + i + 1) - beg;
                if (len > MAX_KV_TILES * KT) len = 0;  // refused by the host; never reached
            }
            const int nt = (len + QT - 1) / QT;
            int inc = nt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            const int first = total + inc - nt;
            for (int t = 0; t < nt; ++t)
                if (first + t < MAX_TILES) tiles[first + t] = make_int2(beg, len | (t << 16));
            total += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) *n_tiles_s = min(total, MAX_TILES);
    }
    if (threadIdx.x == 64) {
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(smem_u32(kv_full + s), 1);
            mbar_init(smem_u32(kv_empty + s), 1);
        }
        for (int s = 0; s < Q_STAGES; ++s) {
            mbar_init(smem_u32(q_full + s), 1);
            mbar_init(smem_u32(q_empty + s), 1);
        }
        for (int j = 0; j < MAX_KV_TILES; ++j) mbar_init(smem_u32(s_full + j), 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(p_full + a), SM_THREADS);
            mbar_init(smem_u32(p_empty + a), 1);
        }
        mbar_init(smem_u32(o_full), 1);
        mbar_init(smem_u32(tmem_free), SM_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)),
                     "r"(uint32_t(TMEM_COLS))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const int n_items = *n_tiles_s * p.n_heads;  // item = tile * n_heads + head

    if (warp == 8) {
        // ===================================================== TMA producer: Q, then the K tiles, then the V tiles of every item
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            int li = 0;  // items this CTA has started
            for (int it = int(blockIdx.x); it < n_items; it += int(gridDim.x), ++li) {
                const int2 tl = tiles[it / p.n_heads];
                const int head = it % p.n_heads;
                const int s_beg = tl.x, s_len = tl.y & 0xffff, row0 = (tl.y >> 16) * QT;
                const int n_kv = (s_len + KT - 1) / KT;
                const int qc = head * HD, kc = p.hidden + head * HD, vc = 2 * p.hidden + head * HD;
                const int qs = li & 1;
                mbar_wait(smem_u32(q_empty + qs), (uint32_t(li >> 1) & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(q_full + qs), TILE_BYTES);
                tma_load_2d(smem_u32(q_s + size_t(qs) * TILE_BYTES), &map_qkv, qc, s_beg + row0, smem_u32(q_full + qs), POLICY_EVICT_FIRST);
                for (int j = 0; j < 2 * n_kv; ++j) {  // K_0 .. K_{n-1}, V_0 .. V_{n-1}
                    const bool is_v = j >= n_kv;
                    mbar_wait(smem_u32(kv_empty + st), ph ^ 1u);
                    mbar_expect_tx(smem_u32(kv_full + st), TILE_BYTES);
                    tma_load_2d(smem_u32(kv_s + size_t(st) * TILE_BYTES), &map_qkv, is_v ? vc : kc, s_beg + (is_v ? j - n_kv : j) * KT,
                                smem_u32(kv_full + st), POLICY_EVICT_LAST);
                    if (++st == KV_STAGES) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc_s = idesc_bf16(KT, false);
            constexpr uint32_t idesc_o = idesc_bf16(HD, true);
            int st = 0;
            uint32_t ph = 0;
            int li = 0;
            uint32_t pt = 0;  // P tiles consumed so far (over all items)
            for (int it = int(blockIdx.x); it < n_items; it += int(gridDim.x), ++li) {
                const int2 tl = tiles[it / p.n_heads];
                const int n_kv = ((tl.y & 0xffff) + KT - 1) / KT;
                const int qs = li & 1;
                mbar_wait(smem_u32(tmem_free), (uint32_t(li) & 1u) ^ 1u);  // the previous item's O and S tiles have been read
                mbar_wait(smem_u32(q_full + qs), uint32_t(li >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t q_base = smem_u32(q_s + size_t(qs) * TILE_BYTES);
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(smem_u32(kv_full + st), ph);
                    tcgen05_fence_after();
                    const uint32_t k_base = smem_u32(kv_s + size_t(st) * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tmem_base + uint32_t(j * KT), umma_desc_sw128(q_base + k * 32), umma_desc_sw128(k_base + k * 32), idesc_s,
                                  uint32_t(k != 0));
                    umma_commit(smem_u32(kv_empty + st));  // the K tile's ring slot is free once these MMAs retire
                    umma_commit(smem_u32(s_full + j));
                    if (++st == KV_STAGES) { st = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(q_empty + qs));       // Q is only read by the S MMAs
                for (int j = 0; j < n_kv; ++j, ++pt) {
                    const uint32_t pb = pt & 1u;
                    mbar_wait(smem_u32(kv_full + st), ph);
                    mbar_wait(smem_u32(p_full + pb), (pt >> 1) & 1u);
                    tcgen05_fence_after();
                    const uint32_t p_base = smem_u32(p_s + size_t(pb) * P_BYTES);
                    const uint32_t v_base = smem_u32(kv_s + size_t(st) * TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < KT / 16; ++kk)  // 16 keys per instruction: P chunk kk / 4, V rows 16 kk ..
                        umma_bf16(tmem_base, umma_desc_sw128(p_base + (kk >> 2) * TILE_BYTES + (kk & 3) * 32),
                                  umma_desc_mn_sw128(v_base + kk * 2048), idesc_o, uint32_t((j | kk) != 0));
                    umma_commit(smem_u32(kv_empty + st));
                    umma_commit(smem_u32(p_empty + pb));
                    if (++st == KV_STAGES) { st = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(o_full));
            }
        }
        __syncwarp();
    } else {
        // ===================================================== softmax + epilogue: thread = (query row = TMEM lane, column half)
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;  // query row within the tile
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        float* mx = xch;            // [2][QT]
        float* ls = xch + 2 * QT;   // [2][QT]
        uint32_t s_uses[MAX_KV_TILES] = {0u, 0u, 0u, 0u};  // completed phases of every s_full barrier
        uint32_t pt = 0;
        int li = 0;
        for (int it = int(blockIdx.x); it < n_items; it += int(gridDim.x), ++li) {
            const int2 tl = tiles[it / p.n_heads];
            const int head = it % p.n_heads;
            const int s_beg = tl.x, s_len = tl.y & 0xffff, row0 = (tl.y >> 16) * QT;
            const int n_kv = (s_len + KT - 1) / KT;
            // ---- pass A: row maximum over this thread's half of every S tile
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < MAX_KV_TILES; ++j) {
                if (j < n_kv) {
                    const int valid = min(64, max(0, s_len - j * KT - half * 64));  // of this thread's 64 columns
                    mbar_wait(smem_u32(s_full + j), s_uses[j] & 1u);
                    ++s_uses[j];
                    tcgen05_fence_after();
                    float v[64];
                    tmem_ld_x32(lane_addr + uint32_t(j * KT + half * 64), v);
                    tmem_ld_x32(lane_addr + uint32_t(j * KT + half * 64 + 32), v + 32);
                    tmem_ld_wait();
                    if (valid == 64) {
#pragma unroll
                        for (int i = 0; i < 64; i += 4) {
                            m0 = fmaxf(m0, v[i]);
                            m1 = fmaxf(m1, v[i + 1]);
                            m2 = fmaxf(m2, v[i + 2]);
                            m3 = fmaxf(m3, v[i + 3]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 64; ++i)
                            if (i < valid) m0 = fmaxf(m0, v[i]);
                    }
                }
            }
            mx[half * QT + r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            epi_bar_sync<SM_THREADS>();
            const float m = fmaxf(mx[r], mx[QT + r]);  // finite: column 0 of tile 0 is always a valid key
            const float mc = m * p.scale_log2e;
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            // ---- pass B: numerators with the final maximum -> this thread's 128-byte line of the P tile (chunk = half)
            for (int j = 0; j < n_kv; ++j, ++pt) {
                const uint32_t pb = pt & 1u;
                const int valid = min(64, max(0, s_len - j * KT - half * 64));
                float v[64];
                tmem_ld_x32(lane_addr + uint32_t(j * KT + half * 64), v);
                tmem_ld_x32(lane_addr + uint32_t(j * KT + half * 64 + 32), v + 32);
                mbar_wait(smem_u32(p_empty + pb), ((pt >> 1) & 1u) ^ 1u);  // the P V MMA that read this buffer has retired
                tmem_ld_wait();
                uint32_t w[32];
                if (valid == 64) {
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        const float e0 = ex2_approx(fmaf(v[i], p.scale_log2e, -mc)), e1 = ex2_approx(fmaf(v[i + 1], p.scale_log2e, -mc));
                        const float e2 = ex2_approx(fmaf(v[i + 2], p.scale_log2e, -mc)), e3 = ex2_approx(fmaf(v[i + 3], p.scale_log2e, -mc));
                        l0 += e0;
                        l1 += e1;
                        l2 += e2;
                        l3 += e3;
                        w[i >> 1] = pack_bf16x2(e0, e1);
                        w[(i >> 1) + 1] = pack_bf16x2(e2, e3);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 64; i += 2) {
                        const float e0 = (i < valid) ? ex2_approx(fmaf(v[i], p.scale_log2e, -mc)) : 0.f;
                        const float e1 = (i + 1 < valid) ? ex2_approx(fmaf(v[i + 1], p.scale_log2e, -mc)) : 0.f;
                        l0 += e0;
                        l1 += e1;
                        w[i >> 1] = pack_bf16x2(e0, e1);
                    }
                }
                // eight 16-byte units of row r's line in chunk `half`; SW128: unit ^= row % 8
                unsigned char* line = p_s + size_t(pb) * P_BYTES + size_t(half) * TILE_BYTES + size_t(r) * 128;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<uint4*>(line + ((u ^ (r & 7)) << 4)) = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                tcgen05_fence_before();     // this thread's reads of S_j (for j = 0: the columns O is about to overwrite) are done
                fence_proxy_async_smem();   // the P tile was written by the generic proxy, the MMA reads it through the async proxy
                mbar_arrive(smem_u32(p_full + pb));
            }
            ls[half * QT + r] = (l0 + l1) + (l2 + l3);
            epi_bar_sync<SM_THREADS>();
            const float inv_l = 1.f / (ls[r] + ls[QT + r]);
            // ---- epilogue: this thread's 32 columns of O / l -> bf16 -> out[row, head * 64 + 32 half ..]
            mbar_wait(smem_u32(o_full), uint32_t(li) & 1u);
            tcgen05_fence_after();
            const int qrow = row0 + r;
            float o[32];
            tmem_ld_x32(lane_addr + uint32_t(half * 32), o);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(smem_u32(tmem_free));  // O (and every S tile) of this item is in registers / consumed
            if (qrow < s_len) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t(s_beg) + qrow) * p.hidden + head * HD + half * 32);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[u] = make_uint4(pack_bf16x2(o[8 * u] * inv_l, o[8 * u + 1] * inv_l), pack_bf16x2(o[8 * u + 2] * inv_l, o[8 * u + 3] * inv_l),
                                        pack_bf16x2(o[8 * u + 4] * inv_l, o[8 * u + 5] * inv_l), pack_bf16x2(o[8 * u + 6] * inv_l, o[8 * u + 7] * inv_l));
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 9) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS)) : "memory");
    }
}

// ------------------------------------------------------------------ classification head on the <s> token of every pair
// logit[i] = w2 . tanh(W1 x[first[i]] + b1) + b2   (RobertaClassificationHead, one output unit), fp32 weights.
// One block per pair; thread t owns hidden units t, t + 256, ...
__global__ void __launch_bounds__(256) cls_head_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ cu_seqlens,
                                                       int hidden, const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       float* __restrict__ logits) {
    extern __shared__ float xs[];  // [hidden]
    __shared__ float red[8];
    const int i = blockIdx.x;
    const __nv_bfloat16* row = x + size_t(cu_seqlens[i]) * hidden;
    for (int d = threadIdx.x; d < hidden; d += blockDim.x) xs[d] = __bfloat162float(row[d]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int u = warp; u < hidden; u += 8) {  // one warp per hidden unit: coalesced reads of W1's row u
        const float* wr = w1 + size_t(u) * hidden;
        float s = 0.f;
        for (int d = lane * 4; d < hidden; d += 128) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + d);
            s = fmaf(wv.x, xs[d], s);
            s = fmaf(wv.y, xs[d + 1], s);
            s = fmaf(wv.z, xs[d + 2], s);
            s = fmaf(wv.w, xs[d + 3], s);
        }
        s = warp_sum_f32(s);
        if (lane == 0) acc = fmaf(w2[u], tanhf(s + b1[u]), acc);
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = b2[0];
        for (int w = 0; w < 8; ++w) t += red[w];
        logits[i] = t;
    }
}

}  // namespace attn

int launch_attention_varlen(const void* qkv, int64_t n_tokens, int n_heads, const int* cu_seqlens, int n_seq, int max_len,
                            int max_tiles, float scale, void* out, int n_sms, cudaStream_t st) {
    using namespace attn;
    const int hidden = n_heads * HD;
    if (max_tiles > MAX_TILES) {
        set_error("tt_attention_varlen_bf16: %d query tiles in one call (limit %d: split the batch)", max_tiles, MAX_TILES);
        return TT_ERR_UNSUPPORTED;
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return TT_ERR_CUDA;
    }
    CUtensorMap map;
    cuuint64_t dims[2] = {cuuint64_t(3 * hidden), cuuint64_t(n_tokens)};
    cuuint64_t strides[1] = {cuuint64_t(3 * hidden) * 2};
    cuuint32_t box[2] = {cuuint32_t(HD), cuuint32_t(QT)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for the packed QKV matrix (%lld tokens)", int(r), (long long)n_tokens);
        return TT_ERR_CUDA;
    }
    Params p;
    p.cu_seqlens = cu_seqlens;
    p.n_seq = n_seq;
    p.n_heads = n_heads;
    p.hidden = hidden;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    const size_t smem = 1024 + size_t(Q_STAGES + KV_STAGES) * TILE_BYTES + 2 * P_BYTES + MAX_TILES * sizeof(int2) + 4 * QT * sizeof(float) + 512;
    TT_CUDA_OK(cudaFuncSetAttribute(attn_varlen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_LIMIT)));
    int64_t items = int64_t(max_tiles) * n_heads;
    int grid = int(items < n_sms ? items : n_sms);
    if (grid < 1) grid = 1;
    attn_varlen_kernel<<<grid, A_THREADS, smem, st>>>(map, p);
    TT_LAUNCH_OK("attn_varlen_kernel");
    return TT_OK;
}

int launch_cls_head(const void* x, const int* cu_seqlens, int n_seq, int hidden, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* logits, cudaStream_t st) {
    attn::cls_head_kernel<<<n_seq, 256, size_t(hidden) * sizeof(float), st>>>(reinterpret_cast<const __nv_bfloat16*>(x), cu_seqlens, hidden,
                                                                               w1, b1, w2, b2, logits);
    TT_LAUNCH_OK("cls_head_kernel");
    return TT_OK;
}

}  // namespace tt
