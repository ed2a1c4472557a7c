// y = act(x W^T + bias) (+ residual): the dense layers of the stage AFTER the retrieval path, the cross-encoder
// reranker (reference: SentenceTransformerRerank.postprocess_nodes, wired at
// /root/reference/src/tensortruth/services/model_manager.py:333-337, run at services/rag_service.py:343-346).
//
// Same tcgen05 pipeline as scan_gemm.cu with the operands renamed: activations x [T, K] (bf16, K-major rows) play the
// corpus, an nn.Linear weight W [N_out, K] (bf16, K-major rows) plays the query block; 256 x 256 output tiles per CTA
// pair (cta_group::2, M = 256, N = 256, K = 16), both operands TMA-streamed two 64-column chunks per ring stage,
// accumulators double-buffered in TMEM.  The epilogue is the layer's tail, fused: + bias, optional exact (erf) GELU,
// optional residual add, rounding to bf16, 64-byte stores of each thread's 32 output columns.
#include "tc_ptx.cuh"

namespace tt {

TT_DEFINE_STATUS_HOOKS(linear)
namespace lin {

using namespace tc;

constexpr int NB = 256;                          // output features per tile (MMA N)
constexpr int NH = NB / 2;                       // weight rows held by each CTA of the pair
constexpr int CH = 2;                            // 64-column chunks per ring stage
constexpr int STAGE_A = CH * CHUNK_BYTES;        // 128 activation rows x 128 columns
constexpr int STAGE_B = CH * NH * 128;           // 128 weight rows x 128 columns
constexpr int STAGE_BYTES = STAGE_A + STAGE_B;   // 64 KB
constexpr int STAGES = 3;
constexpr int EPI = 256;
constexpr int THREADS_L = EPI + 64;
constexpr int TMEM_COLS = 2 * NB;

struct Params {
    int64_t n_rows;       // T
    int n_super;          // 256-row super-tiles
    int n_nb;             // n_out / 256
    int n_chunks;         // k_in / 64
    int n_out;
    const float* bias;                 // [n_out] or NULL
    const __nv_bfloat16* residual;     // [T, n_out] or NULL
    __nv_bfloat16* y;                  // [T, n_out]
    int activation;                    // 0 none, 1 gelu (erf)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS_L, 1)
linear_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + size_t(STAGES) * STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = int(blockIdx.x) >> 1, n_clusters = int(gridDim.x) >> 1;
    const int64_t n_work = int64_t(p.n_super) * p.n_nb;  // (row super-tile, feature block), feature block fastest

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(full_bar + s), 2);
            mbar_init(smem_u32(empty_bar + s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full + a), 1);
            mbar_init(smem_u32(tmem_empty + a), 2 * (EPI / 32));
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
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 8) {
        // ===================================================== TMA producer (both CTAs: own activation rows + own weight half)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t w = cluster_id; w < n_work; w += n_clusters) {
                const int nb = int(w % p.n_nb);
                const int super = int(w / p.n_nb);
                const int row0 = super * 256 + int(rank) * TILE_ROWS;
                const int wrow0 = nb * NB + int(rank) * NH;
                for (int c = 0; c < p.n_chunks; c += CH) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1u);
                    const uint32_t fb = mapa(smem_u32(full_bar + stage), 0);
                    if (rank == 0) mbar_expect_tx(smem_u32(full_bar + stage), 2 * STAGE_BYTES);
                    else mbar_arrive_cluster(fb);
                    const uint32_t dst = smem_u32(ring + size_t(stage) * STAGE_BYTES);
                    tma_load_3d_pair(dst, &map_x, 0, row0, c, fb, POLICY_EVICT_NORMAL);
                    tma_load_3d_pair(dst + STAGE_A, &map_w, 0, wrow0, c, fb, POLICY_EVICT_LAST);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== MMA issuer: one thread of the leader CTA
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_m256(NB);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int64_t w = cluster_id; w < n_work; w += n_clusters, ++it) {
                const int a = it & 1;
                mbar_wait(smem_u32(tmem_empty + a), (uint32_t(it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(a * NB);
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
                    umma_commit_pair(smem_u32(empty_bar + stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair(smem_u32(tmem_full + a));
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: thread = (output row, 128 of the 256 features)
        const int quarter = warp & 3, half = warp >> 2;
        const uint32_t te0 = mapa(smem_u32(tmem_empty), 0), te1 = mapa(smem_u32(tmem_empty + 1), 0);
        int it = 0;
        for (int64_t w = cluster_id; w < n_work; w += n_clusters, ++it) {
            const int a = it & 1;
            const int nb = int(w % p.n_nb);
            const int super = int(w / p.n_nb);
            const int64_t row = int64_t(super) * 256 + int64_t(rank) * TILE_ROWS + quarter * 32 + lane;
            const bool row_ok = row < p.n_rows;
            mbar_wait(smem_u32(tmem_full + a), uint32_t(it >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(a * NB + half * (NB / 2));
#pragma unroll 1
            for (int c32 = 0; c32 < NB / 2; c32 += 32) {
                float v[32];
                tmem_ld_x32(taddr + c32, v);
                tmem_ld_wait();
                const int col0 = nb * NB + half * (NB / 2) + c32;
                if (p.bias) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j4);
                        v[j4 * 4 + 0] += b4.x;
                        v[j4 * 4 + 1] += b4.y;
                        v[j4 * 4 + 2] += b4.z;
                        v[j4 * 4 + 3] += b4.w;
                    }
                }
                if (p.activation == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                }
                if (row_ok) {
                    const size_t o = size_t(row) * p.n_out + col0;
                    if (p.residual) {
                        const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + o);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float f[8];
                            unpack_bf16x8(__ldg(r4 + u), f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[u * 8 + j] += f[j];
                        }
                    }
                    uint4* y4 = reinterpret_cast<uint4*>(p.y + o);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(v[u * 8 + 2 * j], v[u * 8 + 2 * j + 1]);
                            pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        y4[u] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a ? te1 : te0);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 9) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS))
                     : "memory");
    }
}

// ------------------------------------------------------------------ LayerNorm over the feature dimension (one warp per row)
// y = (x - mean) / sqrt(var + eps) * gamma + beta, statistics in fp32 over the bf16 inputs; optionally the sum of up
// to three embedding rows first (word + position + token-type: the encoder's input layer).
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t n_rows, int dim,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y,
                                                        const int* __restrict__ word_ids, const int* __restrict__ pos_ids,
                                                        const __nv_bfloat16* __restrict__ word_emb,
                                                        const __nv_bfloat16* __restrict__ pos_emb,
                                                        const __nv_bfloat16* __restrict__ type_emb) {
    const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    constexpr int MAXV = 8;  // dim <= 32 lanes * 8 values * MAXV = 2048
    float v[MAXV * 8];
    const int n_vec = dim / 8;  // uint4 vectors per row
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vec = lane + 32 * i;
        if (vec < n_vec) {
            float f[8];
            if (word_ids) {
                unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(word_emb + size_t(word_ids[row]) * dim) + vec), f);
                float g[8];
                unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(pos_emb + size_t(pos_ids[row]) * dim) + vec), g);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += g[j];
                if (type_emb) {
                    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(type_emb) + vec), g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += g[j];
                }
            } else {
                unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(x + size_t(row) * dim) + vec), f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[i * 8 + j] = f[j];
                sum += f[j];
            }
        }
    }
    const float mean = warp_sum_f32(sum) / float(dim);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (lane + 32 * i < n_vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i * 8 + j] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum_f32(sq) / float(dim) + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vec = lane + 32 * i;
        if (vec < n_vec) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = vec * 8 + 2 * j;
                const float a = (v[i * 8 + 2 * j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
                const float b = (v[i * 8 + 2 * j + 1] - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
                const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                pk[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            reinterpret_cast<uint4*>(y + size_t(row) * dim)[vec] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

}  // namespace lin

int launch_linear(const void* x, int64_t n_rows, int k_in, const void* w, int n_out, const float* bias, const void* residual,
                  int activation, void* y, int n_sms, cudaStream_t st) {
    using namespace lin;
    if (n_rows == 0) return TT_OK;
    Params p;
    p.n_rows = n_rows;
    p.n_super = int((n_rows + 255) / 256);
    p.n_nb = n_out / NB;
    p.n_chunks = k_in / tc::CHUNK_COLS;
    p.n_out = n_out;
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.activation = activation;
    CUtensorMap map_x, map_w;
    int rc = tc::make_map(&map_x, x, n_rows, k_in, k_in, tc::TILE_ROWS, CH);
    if (rc) return rc;
    rc = tc::make_map(&map_w, w, n_out, k_in, k_in, NH, CH);
    if (rc) return rc;
    const size_t smem = 1024 + size_t(STAGES) * STAGE_BYTES + (2 * STAGES + 4) * 8 + 16;
    TT_CUDA_OK(cudaFuncSetAttribute(linear_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
    linear_gemm_kernel<<<n_sms & ~1, THREADS_L, smem, st>>>(map_x, map_w, p);
    TT_LAUNCH_OK("linear_gemm_kernel");
    return TT_OK;
}

int launch_layernorm(const void* x, int64_t n_rows, int dim, const float* gamma, const float* beta, float eps, void* y,
                     const int* word_ids, const int* pos_ids, const void* word_emb, const void* pos_emb, const void* type_emb,
                     cudaStream_t st) {
    if (n_rows == 0) return TT_OK;
    const int warps = 8;
    lin::layernorm_kernel<<<unsigned((n_rows + warps - 1) / warps), warps * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), n_rows, dim, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y), word_ids,
        pos_ids, reinterpret_cast<const __nv_bfloat16*>(word_emb), reinterpret_cast<const __nv_bfloat16*>(pos_emb),
        reinterpret_cast<const __nv_bfloat16*>(type_emb));
    TT_LAUNCH_OK("layernorm_kernel");
    return TT_OK;
}

}  // namespace tt
