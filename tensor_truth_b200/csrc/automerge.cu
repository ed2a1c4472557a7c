// Stage 3: auto-merge over the hierarchical node tree -- one CTA per query, one thread per list slot.
//
// Device restatement of llama_index's AutoMergingRetriever as the reference instantiates it at
// /root/reference/src/tensortruth/rag_engine.py:641-643 (simple_ratio_thresh left at 0.5):
//   loop { _fill_in_nodes; _get_parents_and_merge } until neither changed; stable sort by score.
// The sequential list semantics (insert-after, first-encounter group order, survivors ++ parents,
// left-to-right float64 sums) are reproduced with flags + block prefix sums, so ids AND float64
// scores come out identical to the Python procedure (oracle/automerge.py).
#include "automerge.cuh"

namespace tt {

__global__ void __launch_bounds__(AM_CAP) automerge_kernel(const int64_t* __restrict__ in_ids,
                                                           const float* __restrict__ in_scores, int k, const AmArgs a) {
    __shared__ AmSmem S;
    const int b = blockIdx.x;
    automerge_block(S, in_ids + size_t(b) * k, in_scores + size_t(b) * k, k, b, a);
}

int automerge_max_k() { return AM_CAP / 2; }

int launch_automerge(const int64_t* ids, const float* scores, int n_q, int k, const int32_t* parent_of,
                     const int32_t* child_count, const int32_t* prev_id, const int32_t* next_id, int64_t n_nodes,
                     double ratio_thresh, int max_rounds, int64_t* out_ids, double* out_scores, int32_t* out_len,
                     int max_out, cudaStream_t st) {
    if (n_q == 0) return TT_OK;
    AmArgs a{parent_of, child_count, prev_id, next_id, n_nodes, ratio_thresh, max_rounds, out_ids, out_scores, out_len, max_out};
    automerge_kernel<<<n_q, AM_CAP, 0, st>>>(ids, scores, k, a);
    TT_LAUNCH_OK("automerge_kernel");
    return TT_OK;
}

}  // namespace tt
