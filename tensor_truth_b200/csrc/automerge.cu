// Stage 3: auto-merge over the hierarchical node tree -- one CTA per query, one thread per list slot.
//
// Device restatement of llama_index's AutoMergingRetriever as the reference instantiates it at
// /root/reference/src/tensortruth/rag_engine.py:641-643 (simple_ratio_thresh left at 0.5):
//   loop { _fill_in_nodes; _get_parents_and_merge } until neither changed; stable sort by score.
// The sequential list semantics (insert-after, first-encounter group order, survivors ++ parents,
// left-to-right float64 sums) are reproduced with flags + block prefix sums, so ids AND float64
// scores come out identical to the Python procedure (oracle/automerge.py).
#include "tt_common.cuh"

namespace tt {

constexpr int AM_CAP = 512;  // list capacity == threads per CTA; k <= AM_CAP/2 so one fill-in pass always fits

__global__ void __launch_bounds__(AM_CAP) automerge_kernel(
    const int64_t* __restrict__ in_ids, const float* __restrict__ in_scores, int k,
    const int32_t* __restrict__ parent_of, const int32_t* __restrict__ child_count,
    const int32_t* __restrict__ prev_id, const int32_t* __restrict__ next_id, int64_t n_nodes,
    double ratio_thresh, int max_rounds,
    int64_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_len, int max_out) {
    __shared__ int32_t idA[AM_CAP], idB[AM_CAP];
    __shared__ double scA[AM_CAP], scB[AM_CAP];
    __shared__ int32_t par[AM_CAP];
    __shared__ int32_t flag[AM_CAP];
    __shared__ int warp_sums[33];
    __shared__ int bad;

    const int t = threadIdx.x;
    const int b = blockIdx.x;
    if (t == 0) bad = 0;

    // ---- load: valid entries are a prefix (padding -1 at the tail), keep their order
    int v = 0;
    int64_t my_in = -1;
    if (t < k) {
        my_in = in_ids[size_t(b) * k + t];
        v = (my_in >= 0 && my_in < n_nodes) ? 1 : 0;
    }
    int n;
    int pos = block_excl_scan(v, warp_sums, &n);
    if (v) {
        idA[pos] = int32_t(my_in);
        scA[pos] = double(in_scores[size_t(b) * k + t]);
    }
    __syncthreads();

    bool changed = true;
    int rounds = 0;
    while (changed && rounds < max_rounds) {
        // ================= _fill_in_nodes: A -> B =================
        int f = 0;
        int32_t nxt = -1;
        int32_t my = -1;
        double ms = 0.0;
        if (t < n) {
            my = idA[t];
            ms = scA[t];
            if (t < n - 1) {
                nxt = next_id[my];
                if (nxt != -1 && nxt == prev_id[idA[t + 1]]) f = 1;
            }
        }
        int n_ins;
        int ex = block_excl_scan(f, warp_sums, &n_ins);
        const int n2 = n + n_ins;
        if (n2 > AM_CAP) {  // cannot happen while k <= AM_CAP/2 and merges only shrink; guard anyway
            if (t == 0) bad = 1;
            __syncthreads();
            break;
        }
        if (t < n) {
            idB[t + ex] = my;
            scB[t + ex] = ms;
            if (f) {
                idB[t + ex + 1] = nxt;
                scB[t + ex + 1] = (ms + scA[t + 1]) / 2;
            }
        }
        __syncthreads();

        // ================= _get_parents_and_merge: B -> A =================
        int32_t p = -1;
        if (t < n2) p = parent_of[idB[t]];
        par[t] = p;
        flag[t] = 0;
        __syncthreads();
        int first = -1, cnt = 0;
        if (p >= 0) {
            for (int j = 0; j < n2; ++j) {
                if (par[j] == p) {
                    if (first < 0) first = j;
                    ++cnt;
                }
            }
        }
        const bool leader = (p >= 0 && first == t);
        int merge = 0;
        double mean = 0.0;
        if (leader) {
            double sum = 0.0;  // left-to-right float64, like sum() on CPython <= 3.11
            for (int j = t; j < n2; ++j)
                if (par[j] == p) sum = sum + scB[j];
            int cc = child_count[p];
            if (cc <= 0) cc = 1;
            const double ratio = double(cnt) / double(cc);
            if (ratio > ratio_thresh) {
                merge = 1;
                mean = sum / double(cnt);
                flag[t] = 1;
            }
        }
        __syncthreads();
        const int drop = (p >= 0) ? flag[first] : 0;
        const int keep = (t < n2 && !drop) ? 1 : 0;
        int n_keep, n_merge;
        const int exk = block_excl_scan(keep, warp_sums, &n_keep);
        const int exm = block_excl_scan(merge, warp_sums, &n_merge);
        if (keep) {
            idA[exk] = idB[t];
            scA[exk] = scB[t];
        }
        if (merge) {
            idA[n_keep + exm] = p;
            scA[n_keep + exm] = mean;
        }
        n = n_keep + n_merge;
        changed = (n_ins > 0) || (n_merge > 0);
        ++rounds;
        __syncthreads();
    }

    // ================= stable sort by score, descending =================
    if (t < n) {
        const double s = scA[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double sj = scA[j];
            rank += (sj > s || (sj == s && j < t)) ? 1 : 0;
        }
        if (rank < max_out) {
            out_ids[size_t(b) * max_out + rank] = idA[t];
            out_scores[size_t(b) * max_out + rank] = s;
        }
    }
    for (int i = n + t; i < max_out; i += AM_CAP) {
        out_ids[size_t(b) * max_out + i] = -1;
        out_scores[size_t(b) * max_out + i] = 0.0;
    }
    if (t == 0) out_len[b] = (bad || n > max_out) ? -1 : n;
}

int automerge_max_k() { return AM_CAP / 2; }

int launch_automerge(const int64_t* ids, const float* scores, int n_q, int k, const int32_t* parent_of,
                     const int32_t* child_count, const int32_t* prev_id, const int32_t* next_id, int64_t n_nodes,
                     double ratio_thresh, int max_rounds, int64_t* out_ids, double* out_scores, int32_t* out_len,
                     int max_out, cudaStream_t st) {
    if (n_q == 0) return TT_OK;
    automerge_kernel<<<n_q, AM_CAP, 0, st>>>(ids, scores, k, parent_of, child_count, prev_id, next_id, n_nodes,
                                             ratio_thresh, max_rounds, out_ids, out_scores, out_len, max_out);
    TT_LAUNCH_OK("automerge_kernel");
    return TT_OK;
}

}  // namespace tt
