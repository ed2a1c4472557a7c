// C ABI of libtt_b200.so (include/tt_b200.h): argument checking, variant dispatch, error plumbing.
// No kernels here; no CPU implementation of anything behind these entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "tt_common.cuh"

namespace tt {

// ---- implemented in the kernel translation units
int launch_prepare_queries(const float* q, int n_q, int dim, void* q_hi, void* q_lo, float* rho, cudaStream_t st);
int launch_certificate_credit(float* thresh, int n_q, int n_lists, const float* rho, float eps_hi_only, cudaStream_t st);
int scan_simt_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                     const void* q_hi, const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids,
                     float* out_approx, float* out_thresh, int n_lists, cudaStream_t st);
int scan_simt_exact(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t stride, const float* q_f32,
                    int n_q, int kprime, int64_t id_base, int mode, uint64_t* out_packed, int n_lists, cudaStream_t st,
                    const float* row_gate);
bool scan_tc_supported(int64_t n_rows, int dim, int64_t stride, int kprime, const void* corpus);
int scan_tc_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                   const void* q_hi, const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids,
                   float* out_approx, float* out_thresh, int n_lists, int* sched, cudaStream_t st);
int scan_tc_segmented(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                      const void* q_lo, int n_q, int kprime, int64_t id_base, const int64_t* seg_end, int n_seg,
                      int64_t* out_ids, float* out_approx, float* out_thresh, int n_lists, int* sched, cudaStream_t st);
bool scan_tc2_supported(int dim, int kprime, int n_lists);
int scan_tc2_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm,
                    const void* q_hi, int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx,
                    float* out_thresh, int n_lists, cudaStream_t st);
size_t scan_gemm_workspace_bytes(int n_q, int kprime);
bool scan_gemm_supported(int dim, int kprime, int n_lists);
int scan_gemm_approx(const void* corpus, int64_t n_rows, int dim, int64_t stride, const float* inv_norm, const void* q_hi,
                     const void* q_lo, int n_q, int kprime, int64_t id_base, int64_t* out_ids, float* out_approx,
                     float* out_thresh, void* ws, int n_sms, cudaStream_t st);
int launch_linear(const void* x, int64_t n_rows, int k_in, const void* w, int n_out, const float* bias, const void* residual,
                  int activation, void* y, int n_sms, cudaStream_t st);
int launch_layernorm(const void* x, int64_t n_rows, int dim, const float* gamma, const float* beta, float eps, void* y,
                     const int* word_ids, const int* pos_ids, const void* word_emb, const void* pos_emb, const void* type_emb,
                     cudaStream_t st);
int launch_rescore(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                   const float* q, int n_q, const int64_t* cand_ids, int n_cand, int mode, uint64_t* packed,
                   cudaStream_t st);
struct RescoreSrc;
int launch_select(const uint64_t* packed, int n_in, const float* in_keys, const int64_t* in_ids, int n_lists,
                  int64_t keys_stride, int64_t ids_stride, int n_q, int k_in, int k, int mode, const float* thresh, int n_thresh, float* out_keys, float* out_scores,
                  int64_t* out_ids, float* out_margin, cudaStream_t st, const tt_exchange_t* xh = nullptr, bool push = false,
                  bool wait = false, const tt_l2_cert_t* l2 = nullptr, const float* q_f32 = nullptr, int dim = 0,
                  const tt_automerge_args_t* amh = nullptr, float* out_all_margins = nullptr, const RescoreSrc* rs = nullptr);
int launch_rescore_select(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                          const float* q, int n_q, const int64_t* cand_ids, int n_cand, const float* thresh, int n_thresh,
                          int k, int mode, float* out_keys, float* out_scores, int64_t* out_ids, float* out_margin,
                          uint64_t* packed, unsigned* tickets, const tt_exchange_t* xh, const tt_l2_cert_t* l2,
                          const tt_automerge_args_t* amh, cudaStream_t st);
int launch_stage2_prefilter(const void* corpus, int dtype, int64_t n_rows, int dim, int64_t stride, int64_t id_base,
                            const float* q, int n_q, const int64_t* cand_ids, const float* cand_approx, int n_cand, float window,
                            const float* thresh, int n_thresh, int k, int mode, float* out_keys, float* out_scores,
                            int64_t* out_ids, float* out_margin, const tt_exchange_t* xh, const tt_automerge_args_t* amh,
                            cudaStream_t st);
bool stage2_prefilter_supported(int n_cand, int k, int mode);
int launch_exchange_push(const void* rec, size_t nbytes, const tt_exchange_t* h, cudaStream_t st);
int launch_peer_barrier(const tt_exchange_t* h, cudaStream_t st);
// per-translation-unit status words (tt_common.cuh)
#define TT_STATUS_TU(tu)                                       \
    int tu##_status_configure(unsigned* mapped, long long cycles); \
    int tu##_status_read(unsigned* out, bool clear);
TT_STATUS_TU(rescore)
TT_STATUS_TU(scan_tc)
TT_STATUS_TU(scan_tc2)
TT_STATUS_TU(scan_gemm)
TT_STATUS_TU(linear)
TT_STATUS_TU(attention)
#undef TT_STATUS_TU
int launch_attention_varlen(const void* qkv, int64_t n_tokens, int n_heads, const int* cu_seqlens, int n_seq, int max_len,
                            int max_tiles, float scale, void* out, int n_sms, cudaStream_t st);
int launch_cls_head(const void* x, const int* cu_seqlens, int n_seq, int hidden, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* logits, cudaStream_t st);
int launch_automerge(const int64_t* ids, const float* scores, int n_q, int k, const int32_t* parent_of,
                     const int32_t* child_count, const int32_t* prev_id, const int32_t* next_id, int64_t n_nodes,
                     double ratio_thresh, int max_rounds, int64_t* out_ids, double* out_scores, int32_t* out_len,
                     int max_out, cudaStream_t st);
int automerge_max_k();
int kprime_to_E(int kprime);

// ---- error plumbing
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) on = getenv("TT_PDL") ? 1 : 0;  // opt-in until measured
    return on == 1;
}

int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

int sm_count(int device) {
    static int cache[64];
    if (device < 0 || device >= 64) return 0;
    if (cache[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        cache[device] = n;
    }
    return cache[device];
}

static int exact_kprime(int k) {
    if (k <= 32) return 32;
    if (k <= 64) return 64;
    if (k <= 128) return 128;
    if (k <= 256) return 256;
    return -1;
}

}  // namespace tt

using namespace tt;

#define TT_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int tt_version(void) { return 201; }

const char* tt_last_error(void) { return g_err; }

int tt_status_configure(uint32_t* mapped_word_host, int timeout_ms) {
    TT_CHECK_ARG(timeout_ms >= 0, "tt_status_configure: timeout_ms=%d", timeout_ms);
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, current_device()) != cudaSuccess || khz <= 0) {
        cudaGetLastError();
        khz = 1965000;
    }
    const long long cycles = timeout_ms > 0 ? (long long)timeout_ms * khz : 8000000000ll;  // clock64 ticks at the SM clock
    unsigned* dev_view = nullptr;
    if (mapped_word_host) {
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, mapped_word_host, 0) != cudaSuccess) {
            cudaGetLastError();
            set_error("tt_status_configure: the status word must live in pinned (device-mapped) host memory");
            return TT_ERR_INVALID;
        }
        dev_view = reinterpret_cast<unsigned*>(d);
    }
    if (rescore_status_configure(dev_view, cycles) || scan_tc_status_configure(dev_view, cycles) ||
        scan_tc2_status_configure(dev_view, cycles) || scan_gemm_status_configure(dev_view, cycles) ||
        linear_status_configure(dev_view, cycles) || attention_status_configure(dev_view, cycles)) {
        set_error("tt_status_configure: %s", cudaGetErrorString(cudaGetLastError()));
        return TT_ERR_CUDA;
    }
    return TT_OK;
}

int tt_status_read(uint32_t* out_host, int clear) {
    TT_CHECK_ARG(out_host != nullptr, "tt_status_read: null pointer");
    unsigned v = 0u;
    const bool c = clear != 0;
    if (rescore_status_read(&v, c) || scan_tc_status_read(&v, c) || scan_tc2_status_read(&v, c) ||
        scan_gemm_status_read(&v, c) || linear_status_read(&v, c) || attention_status_read(&v, c)) {
        set_error("tt_status_read: %s", cudaGetErrorString(cudaGetLastError()));
        return TT_ERR_CUDA;
    }
    *out_host = v;
    return TT_OK;
}

int tt_stream_synchronize(void* stream) {
    TT_CUDA_OK(cudaStreamSynchronize(TT_STREAM(stream)));
    return TT_OK;
}

int tt_scan_num_lists(int device) { return sm_count(device); }

int tt_scan_max_kprime(void) { return 128; }

size_t tt_scan_workspace_bytes(void) { return 256; }

int tt_prepare_queries(const float* q_f32, int n_q, int dim, void* q_hi_bf16, void* q_lo_bf16, void* stream) {
    TT_CHECK_ARG(n_q >= 0 && dim > 0, "tt_prepare_queries: n_q=%d dim=%d", n_q, dim);
    TT_CHECK_ARG(n_q == 0 || (q_f32 && q_hi_bf16), "tt_prepare_queries: null pointer");
    return launch_prepare_queries(q_f32, n_q, dim, q_hi_bf16, q_lo_bf16, nullptr, TT_STREAM(stream));
}

int tt_prepare_queries_rho(const float* q_f32, int n_q, int dim, void* q_hi_bf16, void* q_lo_bf16, float* out_rho, void* stream) {
    TT_CHECK_ARG(n_q >= 0 && dim > 0, "tt_prepare_queries_rho: n_q=%d dim=%d", n_q, dim);
    TT_CHECK_ARG(n_q == 0 || (q_f32 && q_hi_bf16 && out_rho), "tt_prepare_queries_rho: null pointer");
    return launch_prepare_queries(q_f32, n_q, dim, q_hi_bf16, q_lo_bf16, out_rho, TT_STREAM(stream));
}

int tt_certificate_credit(float* cand_thresh, int n_q, int n_lists, const float* rho, float eps_hi_only, void* stream) {
    TT_CHECK_ARG(n_q >= 0 && n_lists >= 1 && eps_hi_only >= 0.f, "tt_certificate_credit: n_q=%d n_lists=%d eps=%f", n_q, n_lists,
                 double(eps_hi_only));
    TT_CHECK_ARG(n_q == 0 || (cand_thresh && rho), "tt_certificate_credit: null pointer");
    return launch_certificate_credit(cand_thresh, n_q, n_lists, rho, eps_hi_only, TT_STREAM(stream));
}

int tt_scan_topk_bf16(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                      const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16, int n_q, int kprime,
                      int64_t id_base, int variant, int64_t* out_ids, float* out_approx, float* out_thresh,
                      void* ws, size_t ws_bytes, void* stream) {
    TT_CHECK_ARG(ws == nullptr || ws_bytes >= tt_scan_workspace_bytes(), "tt_scan_topk_bf16: workspace %zu < %zu bytes",
                 ws_bytes, tt_scan_workspace_bytes());
    TT_CHECK_ARG(n_rows >= 0 && n_q >= 0, "tt_scan_topk_bf16: n_rows=%lld n_q=%d", (long long)n_rows, n_q);
    TT_CHECK_ARG(dim > 0 && dim % 8 == 0, "tt_scan_topk_bf16: dim=%d must be a positive multiple of 8", dim);
    TT_CHECK_ARG(row_stride_elems >= dim && row_stride_elems % 8 == 0, "tt_scan_topk_bf16: row stride %lld",
                 (long long)row_stride_elems);
    TT_CHECK_ARG(kprime == 32 || kprime == 64 || kprime == 128, "tt_scan_topk_bf16: kprime=%d not in {32,64,128}", kprime);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32), "tt_scan_topk_bf16: ids must stay below 2^32");
    TT_CHECK_ARG(n_rows < (int64_t(1) << 32), "tt_scan_topk_bf16: shard too large");
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(q_hi_bf16 && out_ids && out_approx && out_thresh && (n_rows == 0 || corpus_bf16),
                 "tt_scan_topk_bf16: null pointer");
    const int n_lists = sm_count(current_device());
    if (n_lists <= 0) {
        set_error("tt_scan_topk_bf16: no CUDA device");
        return TT_ERR_CUDA;
    }
    const bool tc_ok = n_rows > 0 && scan_tc_supported(n_rows, dim, row_stride_elems, kprime, corpus_bf16);
    if (variant == TT_SCAN_AUTO) variant = tc_ok ? TT_SCAN_TCGEN05 : TT_SCAN_SIMT;
    if (variant == TT_SCAN_TCGEN05) {
        if (!tc_ok) {
            set_error("tt_scan_topk_bf16: the tcgen05 variant needs dim %% 128 == 0, 128 <= dim <= 2048, n_rows > 0");
            return TT_ERR_UNSUPPORTED;
        }
        // wide hi-only batches: CTA pairs (cta_group::2) hold half of the query block each -> 64 queries per pass
        if (!q_lo_bf16 && n_q > 32 && scan_tc2_supported(dim, kprime, n_lists) && !getenv("TT_SCAN_NO_PAIR"))
            return scan_tc2_approx(corpus_bf16, n_rows, dim, row_stride_elems, inv_norm, q_hi_bf16, n_q, kprime, id_base,
                                   out_ids, out_approx, out_thresh, n_lists, TT_STREAM(stream));
        return scan_tc_approx(corpus_bf16, n_rows, dim, row_stride_elems, inv_norm, q_hi_bf16, q_lo_bf16, n_q, kprime,
                              id_base, out_ids, out_approx, out_thresh, n_lists, reinterpret_cast<int*>(ws),
                              TT_STREAM(stream));
    }
    if (variant == TT_SCAN_SIMT)
        return scan_simt_approx(corpus_bf16, n_rows, dim, row_stride_elems, inv_norm, q_hi_bf16, q_lo_bf16, n_q, kprime,
                                id_base, out_ids, out_approx, out_thresh, n_lists, TT_STREAM(stream));
    set_error("tt_scan_topk_bf16: unknown variant %d", variant);
    return TT_ERR_INVALID;
}

int tt_scan_topk_bf16_segmented(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                                const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16, int n_q, int kprime,
                                int64_t id_base, const int64_t* seg_end_host, int n_seg, int64_t* out_ids,
                                float* out_approx, float* out_thresh, void* ws, size_t ws_bytes, void* stream) {
    TT_CHECK_ARG(ws == nullptr || ws_bytes >= tt_scan_workspace_bytes(), "tt_scan_topk_bf16_segmented: workspace %zu < %zu bytes",
                 ws_bytes, tt_scan_workspace_bytes());
    TT_CHECK_ARG(n_rows > 0 && n_q >= 0, "tt_scan_topk_bf16_segmented: n_rows=%lld n_q=%d", (long long)n_rows, n_q);
    TT_CHECK_ARG(kprime == 32 || kprime == 64 || kprime == 128, "tt_scan_topk_bf16_segmented: kprime=%d not in {32,64,128}", kprime);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32), "tt_scan_topk_bf16_segmented: ids must stay below 2^32");
    TT_CHECK_ARG(seg_end_host && n_seg >= 1 && n_seg <= TT_MAX_SEGMENTS, "tt_scan_topk_bf16_segmented: n_seg=%d not in [1, %d]",
                 n_seg, TT_MAX_SEGMENTS);
    for (int s = 0; s < n_seg; ++s)
        TT_CHECK_ARG(seg_end_host[s] >= (s ? seg_end_host[s - 1] : 0) && seg_end_host[s] <= n_rows,
                     "tt_scan_topk_bf16_segmented: segment ends must be non-decreasing and <= n_rows");
    TT_CHECK_ARG(seg_end_host[n_seg - 1] == n_rows, "tt_scan_topk_bf16_segmented: the last segment must end at n_rows");
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(corpus_bf16 && q_hi_bf16 && q_lo_bf16 && out_ids && out_approx && out_thresh,
                 "tt_scan_topk_bf16_segmented: null pointer");
    const int n_lists = sm_count(current_device());
    if (n_lists <= 0) {
        set_error("tt_scan_topk_bf16_segmented: no CUDA device");
        return TT_ERR_CUDA;
    }
    return scan_tc_segmented(corpus_bf16, n_rows, dim, row_stride_elems, inv_norm, q_hi_bf16, q_lo_bf16, n_q, kprime, id_base,
                             seg_end_host, n_seg, out_ids, out_approx, out_thresh, n_lists, reinterpret_cast<int*>(ws),
                             TT_STREAM(stream));
}

size_t tt_scan_gemm_workspace_bytes(int n_q, int kprime) {
    if (n_q <= 0 || kprime <= 0) return 0;
    return scan_gemm_workspace_bytes(n_q, kprime);
}

int tt_scan_gemm_topk_bf16(const void* corpus_bf16, int64_t n_rows, int dim, int64_t row_stride_elems,
                           const float* inv_norm, const void* q_hi_bf16, const void* q_lo_bf16, int n_q, int kprime,
                           int64_t id_base, int64_t* out_ids, float* out_approx, float* out_thresh, void* ws, size_t ws_bytes,
                           void* stream) {
    TT_CHECK_ARG(!q_lo_bf16 || (n_q <= 32 && (reinterpret_cast<uintptr_t>(q_lo_bf16) & 15) == 0),
                 "tt_scan_gemm_topk_bf16: the hi+lo pass takes at most 32 queries (n_q=%d), 16-byte aligned", n_q);
    TT_CHECK_ARG(n_rows >= 0 && n_q >= 0, "tt_scan_gemm_topk_bf16: n_rows=%lld n_q=%d", (long long)n_rows, n_q);
    TT_CHECK_ARG(dim > 0 && dim % 64 == 0, "tt_scan_gemm_topk_bf16: dim=%d must be a positive multiple of 64", dim);
    TT_CHECK_ARG(row_stride_elems >= dim && row_stride_elems % 8 == 0, "tt_scan_gemm_topk_bf16: row stride %lld",
                 (long long)row_stride_elems);
    TT_CHECK_ARG(kprime == 128 || kprime == 256 || kprime == 512, "tt_scan_gemm_topk_bf16: kprime=%d not in {128,256,512}",
                 kprime);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32) && n_rows < (int64_t(1) << 31) * 256,
                 "tt_scan_gemm_topk_bf16: ids must stay below 2^32");
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(q_hi_bf16 && out_ids && out_approx && out_thresh && (n_rows == 0 || corpus_bf16),
                 "tt_scan_gemm_topk_bf16: null pointer");
    TT_CHECK_ARG((reinterpret_cast<uintptr_t>(corpus_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_hi_bf16) & 15) == 0,
                 "tt_scan_gemm_topk_bf16: corpus and queries must be 16-byte aligned");
    const int n_sms = sm_count(current_device());
    if (n_sms <= 0) {
        set_error("tt_scan_gemm_topk_bf16: no CUDA device");
        return TT_ERR_CUDA;
    }
    if (!scan_gemm_supported(dim, kprime, n_sms)) {
        set_error("tt_scan_gemm_topk_bf16: dim=%d kprime=%d unsupported on this device", dim, kprime);
        return TT_ERR_UNSUPPORTED;
    }
    const size_t need = tt_scan_gemm_workspace_bytes(n_q, kprime);
    if (!ws || ws_bytes < need) {
        set_error("tt_scan_gemm_topk_bf16: workspace %zu < %zu bytes", ws_bytes, need);
        return TT_ERR_WORKSPACE;
    }
    return scan_gemm_approx(corpus_bf16, n_rows, dim, row_stride_elems, inv_norm, q_hi_bf16, q_lo_bf16, n_q, kprime, id_base,
                            out_ids, out_approx, out_thresh, ws, n_sms, TT_STREAM(stream));
}

size_t tt_rescore_workspace_bytes(int n_q, int n_cand) {
    if (n_q <= 0 || n_cand <= 0) return 0;
    return size_t(n_q) * size_t(n_cand) * sizeof(uint64_t);
}

int tt_rescore_topk(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                    int64_t id_base, const float* q_f32, int n_q, const int64_t* cand_ids, int n_cand,
                    const float* cand_thresh, int n_lists, int k, int score_mode, float* out_keys, float* out_scores,
                    int64_t* out_ids, float* out_margin, void* ws, size_t ws_bytes, void* stream) {
    return tt_rescore_topk_push(corpus, corpus_dtype, n_rows, dim, row_stride_elems, id_base, q_f32, n_q, cand_ids, n_cand,
                                cand_thresh, n_lists, k, score_mode, out_keys, out_scores, out_ids, out_margin, ws, ws_bytes,
                                nullptr, nullptr, stream);
}

int tt_rescore_topk_push(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                         int64_t id_base, const float* q_f32, int n_q, const int64_t* cand_ids, int n_cand,
                         const float* cand_thresh, int n_lists, int k, int score_mode, float* out_keys, float* out_scores,
                         int64_t* out_ids, float* out_margin, void* ws, size_t ws_bytes, const tt_exchange_t* xchg,
                         const tt_l2_cert_t* l2_cert, void* stream) {
    TT_CHECK_ARG(!l2_cert || (l2_cert->row_norm_min >= 0.f && l2_cert->row_norm_max >= l2_cert->row_norm_min &&
                              l2_cert->eps >= 0.f),
                 "tt_rescore_topk: bad L2 certificate bounds");
    TT_CHECK_ARG(corpus_dtype == TT_DTYPE_BF16 || corpus_dtype == TT_DTYPE_F32, "tt_rescore_topk: dtype %d", corpus_dtype);
    TT_CHECK_ARG(score_mode == TT_SCORE_COSINE || score_mode == TT_SCORE_CHROMA_L2_EXP, "tt_rescore_topk: score_mode %d",
                 score_mode);
    TT_CHECK_ARG(dim > 0 && dim % 8 == 0 && row_stride_elems >= dim && row_stride_elems % 8 == 0,
                 "tt_rescore_topk: dim=%d stride=%lld", dim, (long long)row_stride_elems);
    TT_CHECK_ARG(n_q >= 0 && n_cand >= 0 && k >= 1, "tt_rescore_topk: n_q=%d n_cand=%d k=%d", n_q, n_cand, k);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32), "tt_rescore_topk: ids must stay below 2^32");
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(q_f32 && (out_ids || xchg) && (n_cand == 0 || cand_ids), "tt_rescore_topk: null pointer");
    if (ws_bytes < tt_rescore_workspace_bytes(n_q, n_cand) || (n_cand > 0 && !ws)) {
        set_error("tt_rescore_topk: workspace %zu < %zu bytes", ws_bytes, tt_rescore_workspace_bytes(n_q, n_cand));
        return TT_ERR_WORKSPACE;
    }
    uint64_t* packed = reinterpret_cast<uint64_t*>(ws);
    int rc = launch_rescore(corpus, corpus_dtype, n_rows, dim, row_stride_elems, id_base, q_f32, n_q, cand_ids, n_cand,
                            score_mode, packed, TT_STREAM(stream));
    if (rc) return rc;
    return launch_select(packed, n_cand, nullptr, nullptr, 0, 0, 0, n_q, 0, k, score_mode, cand_thresh,
                         cand_thresh ? n_lists : 0, out_keys, out_scores, out_ids, out_margin, TT_STREAM(stream), xchg,
                         xchg != nullptr, false, l2_cert, q_f32, dim);
}

int tt_exchange_push(const void* record, size_t nbytes, const tt_exchange_t* xchg, void* stream) {
    TT_CHECK_ARG(record && xchg, "tt_exchange_push: null pointer");
    return launch_exchange_push(record, nbytes, xchg, TT_STREAM(stream));
}

int tt_peer_barrier(const tt_exchange_t* xchg, void* stream) {
    TT_CHECK_ARG(xchg != nullptr, "tt_peer_barrier: null pointer");
    return launch_peer_barrier(xchg, TT_STREAM(stream));
}

size_t tt_rescore_fused_workspace_bytes(int n_q, int n_cand) {
    if (n_q <= 0 || n_cand < 0) return 0;
    return size_t(n_q) * size_t(n_cand) * sizeof(uint64_t) + size_t(n_q) * sizeof(uint32_t);
}

int tt_rescore_topk_fused(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                          int64_t id_base, const float* q_f32, int n_q, const int64_t* cand_ids, int n_cand,
                          const float* cand_thresh, int n_lists, int k, int score_mode, float* out_keys, float* out_scores,
                          int64_t* out_ids, float* out_margin, void* ws, size_t ws_bytes, const tt_exchange_t* xchg,
                          const tt_l2_cert_t* l2_cert, const tt_automerge_args_t* am, const float* cand_approx,
                          float prefilter_window, void* stream) {
    TT_CHECK_ARG(!l2_cert || (l2_cert->row_norm_min >= 0.f && l2_cert->row_norm_max >= l2_cert->row_norm_min &&
                              l2_cert->eps >= 0.f),
                 "tt_rescore_topk_fused: bad L2 certificate bounds");
    TT_CHECK_ARG(corpus_dtype == TT_DTYPE_BF16 || corpus_dtype == TT_DTYPE_F32, "tt_rescore_topk_fused: dtype %d", corpus_dtype);
    TT_CHECK_ARG(score_mode == TT_SCORE_COSINE || score_mode == TT_SCORE_CHROMA_L2_EXP, "tt_rescore_topk_fused: score_mode %d",
                 score_mode);
    TT_CHECK_ARG(dim > 0 && dim % 8 == 0 && row_stride_elems >= dim && row_stride_elems % 8 == 0,
                 "tt_rescore_topk_fused: dim=%d stride=%lld", dim, (long long)row_stride_elems);
    TT_CHECK_ARG(n_q >= 0 && n_q <= 65535 && n_cand >= 0 && k >= 1, "tt_rescore_topk_fused: n_q=%d n_cand=%d k=%d", n_q, n_cand, k);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32), "tt_rescore_topk_fused: ids must stay below 2^32");
    TT_CHECK_ARG(!(xchg && am), "tt_rescore_topk_fused: a pushed record is merged before stage 3 (tt_merge_topk_fused)");
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(q_f32 && (out_ids || xchg) && (n_cand == 0 || cand_ids), "tt_rescore_topk_fused: null pointer");
    TT_CHECK_ARG(prefilter_window >= 0.f, "tt_rescore_topk_fused: prefilter_window=%f", double(prefilter_window));
    if (cand_approx && prefilter_window > 0.f && n_cand > 0 && stage2_prefilter_supported(n_cand, k, score_mode))
        return launch_stage2_prefilter(corpus, corpus_dtype, n_rows, dim, row_stride_elems, id_base, q_f32, n_q, cand_ids,
                                       cand_approx, n_cand, prefilter_window, cand_thresh, cand_thresh ? n_lists : 0, k,
                                       score_mode, out_keys, out_scores, out_ids, out_margin, xchg, am, TT_STREAM(stream));
    const size_t need = tt_rescore_fused_workspace_bytes(n_q, n_cand);
    if (ws_bytes < need || !ws) {
        set_error("tt_rescore_topk_fused: workspace %zu < %zu bytes", ws_bytes, need);
        return TT_ERR_WORKSPACE;
    }
    uint64_t* packed = reinterpret_cast<uint64_t*>(ws);
    unsigned* tickets = reinterpret_cast<unsigned*>(packed + size_t(n_q) * n_cand);
    return launch_rescore_select(corpus, corpus_dtype, n_rows, dim, row_stride_elems, id_base, q_f32, n_q, cand_ids, n_cand,
                                 cand_thresh, cand_thresh ? n_lists : 0, k, score_mode, out_keys, out_scores, out_ids,
                                 out_margin, packed, tickets, xchg, l2_cert, am, TT_STREAM(stream));
}

size_t tt_scan_exact_workspace_bytes(int device, int n_q, int k) {
    const int kp = exact_kprime(k);
    const int n_lists = sm_count(device);
    if (kp < 0 || n_lists <= 0 || n_q <= 0) return 0;
    return size_t(n_q) * n_lists * kp * sizeof(uint64_t);
}

int tt_scan_exact_f64(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                      int64_t id_base, const float* q_f32, int n_q, int k, int score_mode, float* out_keys,
                      float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes, void* stream) {
    return tt_scan_exact_f64_gated(corpus, corpus_dtype, n_rows, dim, row_stride_elems, id_base, q_f32, n_q, k, score_mode,
                                   nullptr, out_keys, out_scores, out_ids, ws, ws_bytes, stream);
}

int tt_scan_exact_f64_gated(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride_elems,
                            int64_t id_base, const float* q_f32, int n_q, int k, int score_mode, const float* row_gate,
                            float* out_keys, float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes, void* stream) {
    TT_CHECK_ARG(corpus_dtype == TT_DTYPE_BF16 || corpus_dtype == TT_DTYPE_F32, "tt_scan_exact_f64: dtype %d", corpus_dtype);
    TT_CHECK_ARG(score_mode == TT_SCORE_COSINE || score_mode == TT_SCORE_CHROMA_L2_EXP, "tt_scan_exact_f64: score_mode %d",
                 score_mode);
    TT_CHECK_ARG(dim > 0 && dim % 8 == 0 && row_stride_elems >= dim && row_stride_elems % 8 == 0,
                 "tt_scan_exact_f64: dim=%d stride=%lld", dim, (long long)row_stride_elems);
    TT_CHECK_ARG(n_q >= 0 && n_rows >= 0, "tt_scan_exact_f64: n_q=%d n_rows=%lld", n_q, (long long)n_rows);
    TT_CHECK_ARG(id_base >= 0 && id_base + n_rows <= (int64_t(1) << 32) && n_rows < (int64_t(1) << 32),
                 "tt_scan_exact_f64: ids must stay below 2^32");
    const int kp = exact_kprime(k);
    TT_CHECK_ARG(k >= 1 && kp > 0, "tt_scan_exact_f64: k=%d out of range [1, 256]", k);
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(q_f32 && out_ids && (n_rows == 0 || corpus), "tt_scan_exact_f64: null pointer");
    const int dev = current_device();
    const int n_lists = sm_count(dev);
    if (n_lists <= 0) {
        set_error("tt_scan_exact_f64: no CUDA device");
        return TT_ERR_CUDA;
    }
    const size_t need = tt_scan_exact_workspace_bytes(dev, n_q, k);
    if (ws_bytes < need || !ws) {
        set_error("tt_scan_exact_f64: workspace %zu < %zu bytes", ws_bytes, need);
        return TT_ERR_WORKSPACE;
    }
    uint64_t* packed = reinterpret_cast<uint64_t*>(ws);
    int rc = scan_simt_exact(corpus, corpus_dtype, n_rows, dim, row_stride_elems, q_f32, n_q, kp, id_base, score_mode,
                             packed, n_lists, TT_STREAM(stream), row_gate);
    if (rc) return rc;
    return launch_select(packed, n_lists * kp, nullptr, nullptr, 0, 0, 0, n_q, 0, k, score_mode, nullptr, 0, out_keys,
                         out_scores, out_ids, nullptr, TT_STREAM(stream));
}

int tt_merge_topk(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride, int64_t ids_list_stride,
                  int n_q, int k_in, int k_out, int score_mode, float* out_scores, int64_t* out_ids, void* stream) {
    return tt_merge_topk_pulled(keys, ids, n_lists, keys_list_stride, ids_list_stride, n_q, k_in, k_out, score_mode,
                                out_scores, out_ids, nullptr, stream);
}

int tt_merge_topk_pulled(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride,
                         int64_t ids_list_stride, int n_q, int k_in, int k_out, int score_mode, float* out_scores,
                         int64_t* out_ids, const tt_exchange_t* xchg, void* stream) {
    return tt_merge_topk_fused(keys, ids, n_lists, keys_list_stride, ids_list_stride, n_q, k_in, k_out, score_mode, out_scores,
                               out_ids, xchg, nullptr, nullptr, stream);
}

int tt_merge_topk_fused(const float* keys, const int64_t* ids, int n_lists, int64_t keys_list_stride,
                        int64_t ids_list_stride, int n_q, int k_in, int k_out, int score_mode, float* out_scores,
                        int64_t* out_ids, const tt_exchange_t* xchg, float* out_all_margins,
                        const tt_automerge_args_t* am, void* stream) {
    TT_CHECK_ARG(keys_list_stride >= 0 && ids_list_stride >= 0, "tt_merge_topk: negative list stride");
    TT_CHECK_ARG(!out_all_margins || (xchg && xchg->margins_off_bytes && n_lists == xchg->world),
                 "tt_merge_topk_fused: margins are gathered from an exchange that carries them, one list per rank");
    TT_CHECK_ARG(n_lists >= 1 && n_q >= 0 && k_in >= 1 && k_out >= 1, "tt_merge_topk: n_lists=%d n_q=%d k_in=%d k_out=%d",
                 n_lists, n_q, k_in, k_out);
    TT_CHECK_ARG(score_mode == TT_SCORE_COSINE || score_mode == TT_SCORE_CHROMA_L2_EXP, "tt_merge_topk: score_mode %d",
                 score_mode);
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(keys && ids && out_ids, "tt_merge_topk: null pointer");
    return launch_select(nullptr, 0, keys, ids, n_lists, keys_list_stride, ids_list_stride, n_q, k_in, k_out, score_mode,
                         nullptr, 0, nullptr, out_scores, out_ids, nullptr, TT_STREAM(stream), xchg, false, xchg != nullptr,
                         nullptr, nullptr, 0, am, out_all_margins);
}


static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int tt_linear_bf16(const void* x_bf16, int64_t n_rows, int k_in, const void* w_bf16, int n_out, const float* bias,
                   const void* residual_bf16, int activation, void* y_bf16, void* stream) {
    TT_CHECK_ARG(n_rows >= 0 && n_rows < (int64_t(1) << 31) - 256, "tt_linear_bf16: n_rows=%lld", (long long)n_rows);
    TT_CHECK_ARG(k_in >= 128 && k_in % 128 == 0, "tt_linear_bf16: k_in=%d must be a positive multiple of 128", k_in);
    TT_CHECK_ARG(n_out >= 256 && n_out % 256 == 0, "tt_linear_bf16: n_out=%d must be a positive multiple of 256", n_out);
    TT_CHECK_ARG(activation == TT_ACT_NONE || activation == TT_ACT_GELU, "tt_linear_bf16: activation %d", activation);
    if (n_rows == 0) return TT_OK;
    TT_CHECK_ARG(x_bf16 && w_bf16 && y_bf16, "tt_linear_bf16: null pointer");
    TT_CHECK_ARG(aligned16(x_bf16) && aligned16(w_bf16) && aligned16(y_bf16) && aligned16(residual_bf16) && aligned16(bias),
                 "tt_linear_bf16: pointers must be 16-byte aligned");
    const int n_sms = sm_count(current_device());
    if (n_sms < 2) {
        set_error("tt_linear_bf16: no CUDA device");
        return TT_ERR_CUDA;
    }
    return launch_linear(x_bf16, n_rows, k_in, w_bf16, n_out, bias, residual_bf16, activation, y_bf16, n_sms, TT_STREAM(stream));
}

int tt_layernorm_bf16(const void* x_bf16, int64_t n_rows, int dim, const float* gamma, const float* beta, float eps,
                      void* y_bf16, void* stream) {
    TT_CHECK_ARG(n_rows >= 0 && dim >= 8 && dim % 8 == 0 && dim <= 2048, "tt_layernorm_bf16: n_rows=%lld dim=%d",
                 (long long)n_rows, dim);
    if (n_rows == 0) return TT_OK;
    TT_CHECK_ARG(x_bf16 && gamma && beta && y_bf16 && aligned16(x_bf16) && aligned16(y_bf16), "tt_layernorm_bf16: bad pointer");
    return launch_layernorm(x_bf16, n_rows, dim, gamma, beta, eps, y_bf16, nullptr, nullptr, nullptr, nullptr, nullptr,
                            TT_STREAM(stream));
}

int tt_embed_layernorm_bf16(const int32_t* word_ids, const int32_t* pos_ids, int64_t n_rows, int dim,
                            const void* word_emb_bf16, const void* pos_emb_bf16, const void* type_emb_bf16,
                            const float* gamma, const float* beta, float eps, void* y_bf16, void* stream) {
    TT_CHECK_ARG(n_rows >= 0 && dim >= 8 && dim % 8 == 0 && dim <= 2048, "tt_embed_layernorm_bf16: n_rows=%lld dim=%d",
                 (long long)n_rows, dim);
    if (n_rows == 0) return TT_OK;
    TT_CHECK_ARG(word_ids && pos_ids && word_emb_bf16 && pos_emb_bf16 && gamma && beta && y_bf16 && aligned16(word_emb_bf16) &&
                     aligned16(pos_emb_bf16) && aligned16(type_emb_bf16) && aligned16(y_bf16),
                 "tt_embed_layernorm_bf16: bad pointer");
    return launch_layernorm(nullptr, n_rows, dim, gamma, beta, eps, y_bf16, word_ids, pos_ids, word_emb_bf16, pos_emb_bf16,
                            type_emb_bf16, TT_STREAM(stream));
}

int tt_attention_varlen_bf16(const void* qkv_bf16, int64_t n_tokens, int n_heads, int head_dim, const int32_t* cu_seqlens,
                             int n_seq, int max_len, int max_tiles, float scale, void* out_bf16, void* stream) {
    TT_CHECK_ARG(n_tokens >= 0 && n_tokens < (int64_t(1) << 31) && n_seq >= 0 && n_heads >= 1, "tt_attention_varlen_bf16: n_tokens=%lld n_seq=%d n_heads=%d",
                 (long long)n_tokens, n_seq, n_heads);
    if (head_dim != 64 || max_len > 512) {
        set_error("tt_attention_varlen_bf16: head_dim=%d max_len=%d (this build: head_dim 64, sequences of up to 512 tokens)", head_dim, max_len);
        return TT_ERR_UNSUPPORTED;
    }
    TT_CHECK_ARG(max_len >= 1 && max_tiles >= 0 && max_tiles <= 65535 * 32 && n_heads <= 65535, "tt_attention_varlen_bf16: max_len=%d max_tiles=%d", max_len, max_tiles);
    if (n_tokens == 0 || n_seq == 0 || max_tiles == 0) return TT_OK;
    TT_CHECK_ARG(qkv_bf16 && cu_seqlens && out_bf16 && aligned16(qkv_bf16) && aligned16(out_bf16), "tt_attention_varlen_bf16: bad pointer");
    const int n_sms = sm_count(current_device());
    if (n_sms <= 0) {
        set_error("tt_attention_varlen_bf16: no CUDA device");
        return TT_ERR_CUDA;
    }
    return launch_attention_varlen(qkv_bf16, n_tokens, n_heads, cu_seqlens, n_seq, max_len, max_tiles, scale, out_bf16, n_sms,
                                   TT_STREAM(stream));
}

int tt_cls_head_f32(const void* x_bf16, const int32_t* cu_seqlens, int n_seq, int hidden, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* logits, void* stream) {
    TT_CHECK_ARG(n_seq >= 0 && hidden >= 4 && hidden % 4 == 0 && hidden <= 8192, "tt_cls_head_f32: n_seq=%d hidden=%d", n_seq, hidden);
    if (n_seq == 0) return TT_OK;
    TT_CHECK_ARG(x_bf16 && cu_seqlens && w1 && b1 && w2 && b2 && logits && aligned16(w1), "tt_cls_head_f32: bad pointer");
    return launch_cls_head(x_bf16, cu_seqlens, n_seq, hidden, w1, b1, w2, b2, logits, TT_STREAM(stream));
}

int tt_automerge_max_k(void) { return automerge_max_k(); }

int tt_automerge(const int64_t* ids, const float* scores, int n_q, int k, const int32_t* parent_of,
                 const int32_t* child_count, const int32_t* prev_id, const int32_t* next_id, int64_t n_nodes,
                 double ratio_thresh, int max_rounds, int64_t* out_ids, double* out_scores, int32_t* out_len, int max_out,
                 void* stream) {
    TT_CHECK_ARG(n_q >= 0 && k >= 0 && max_out >= 1, "tt_automerge: n_q=%d k=%d max_out=%d", n_q, k, max_out);
    TT_CHECK_ARG(k <= automerge_max_k(), "tt_automerge: k=%d exceeds tt_automerge_max_k()=%d", k, automerge_max_k());
    TT_CHECK_ARG(n_nodes >= 0 && n_nodes < (int64_t(1) << 31), "tt_automerge: n_nodes=%lld", (long long)n_nodes);
    TT_CHECK_ARG(max_rounds >= 1, "tt_automerge: max_rounds=%d", max_rounds);
    if (n_q == 0) return TT_OK;
    TT_CHECK_ARG(out_ids && out_scores && out_len && (k == 0 || (ids && scores)), "tt_automerge: null pointer");
    TT_CHECK_ARG(n_nodes == 0 || (parent_of && child_count && prev_id && next_id), "tt_automerge: null tree array");
    return launch_automerge(ids, scores, n_q, k, parent_of, child_count, prev_id, next_id, n_nodes, ratio_thresh,
                            max_rounds, out_ids, out_scores, out_len, max_out, TT_STREAM(stream));
}

}  // extern "C"
