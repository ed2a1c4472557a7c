// Stage 3 as a block-level device function, so that it can run as the tail of the kernel that produced the list
// (select / shard merge, rescore.cu) as well as on its own (automerge.cu).  See automerge.cu for what it restates.
#pragma once

#include "tt_common.cuh"

namespace tt {

constexpr int AM_CAP = 512;  // list capacity; k <= AM_CAP/2 so one fill-in pass always fits

// device-side view of tt_automerge_args_t; out_len == nullptr means "no auto-merge"
struct AmArgs {
    const int32_t* parent_of;
    const int32_t* child_count;
    const int32_t* prev_id;
    const int32_t* next_id;
    int64_t n_nodes;
    double ratio_thresh;
    int max_rounds;
    int64_t* out_ids;
    double* out_scores;
    int32_t* out_len;
    int max_out;
};

struct AmSmem {
    int32_t idA[AM_CAP], idB[AM_CAP];
    double scA[AM_CAP], scB[AM_CAP];
    int32_t par[AM_CAP];   // parent of B-list entry j (fill-in phase: prev_id of A-list entry j)
    int32_t cc[AM_CAP];    // child_count of that parent
    uint8_t flag[AM_CAP];
    int warp_sums[33];
    int bad;
};

// Every thread of the block calls this (blockDim.x >= AM_CAP, a multiple of 32); thread t < AM_CAP owns list slot t.
// in_ids / in_scores: the k entries of THIS query (valid entries first, -1 padding); b: the query's output row.
__device__ __forceinline__ void automerge_block(AmSmem& S, const int64_t* in_ids, const float* in_scores, int k, int b,
                                                const AmArgs& a) {
    const int t = threadIdx.x;
    if (t == 0) S.bad = 0;

    // ---- load: valid entries are a prefix (padding -1 at the tail), keep their order
    int v = 0;
    int64_t my_in = -1;
    if (t < k && t < AM_CAP) {
        my_in = in_ids[t];
        v = (my_in >= 0 && my_in < a.n_nodes) ? 1 : 0;
    }
    int n;
    int pos = block_excl_scan(v, S.warp_sums, &n);
    if (v) {
        S.idA[pos] = int32_t(my_in);
        S.scA[pos] = double(in_scores[t]);
    }
    __syncthreads();

    bool changed = true;
    int rounds = 0;
    while (changed && rounds < a.max_rounds) {
        // ================= _fill_in_nodes: A -> B =================
        // Every relation of "my" node is fetched at once (three independent loads: one DRAM latency), its parent's
        // child count right behind them; the merge phase below then runs entirely out of shared memory.
        int32_t my = -1, nxt = -1, prv = -1, pr = -1, ccp = 1;
        double ms = 0.0;
        if (t < n) {
            my = S.idA[t];
            ms = S.scA[t];
            nxt = a.next_id[my];
            prv = a.prev_id[my];
            pr = a.parent_of[my];
            if (pr >= 0) ccp = a.child_count[pr];
        }
        if (t < AM_CAP) S.par[t] = prv;
        __syncthreads();
        int f = 0;
        if (t < n - 1 && nxt != -1 && nxt == S.par[t + 1]) f = 1;
        int32_t pin = -1, ccin = 1;  // the inserted node's own parent (upstream looks it up like any node's)
        if (f) {
            pin = a.parent_of[nxt];
            if (pin >= 0) ccin = (pin == pr) ? ccp : a.child_count[pin];
        }
        int n_ins;
        int ex = block_excl_scan(f, S.warp_sums, &n_ins);  // (its barriers also order the S.par reads above before the writes below)
        const int n2 = n + n_ins;
        if (n2 > AM_CAP) {  // cannot happen while k <= AM_CAP/2 and merges only shrink; guard anyway
            if (t == 0) S.bad = 1;
            __syncthreads();
            break;
        }
        if (t < n) {
            S.idB[t + ex] = my;
            S.scB[t + ex] = ms;
            S.par[t + ex] = pr;
            S.cc[t + ex] = ccp;
            S.flag[t + ex] = 0;
            if (f) {
                S.idB[t + ex + 1] = nxt;
                S.scB[t + ex + 1] = (ms + S.scA[t + 1]) / 2;
                S.par[t + ex + 1] = pin;
                S.cc[t + ex + 1] = ccin;
                S.flag[t + ex + 1] = 0;
            }
        }
        __syncthreads();

        // ================= _get_parents_and_merge: B -> A =================
        int32_t p = -1;
        if (t < n2) p = S.par[t];
        int first = -1, cnt = 0;
        if (p >= 0) {
            for (int j = 0; j < n2; ++j) {
                if (S.par[j] == p) {
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
                if (S.par[j] == p) sum = sum + S.scB[j];
            int cc = S.cc[t];
            if (cc <= 0) cc = 1;
            const double ratio = double(cnt) / double(cc);
            if (ratio > a.ratio_thresh) {
                merge = 1;
                mean = sum / double(cnt);
                S.flag[t] = 1;
            }
        }
        __syncthreads();
        const int drop = (p >= 0) ? int(S.flag[first]) : 0;
        const int keep = (t < n2 && !drop) ? 1 : 0;
        int n_keep, n_merge;
        const int exk = block_excl_scan(keep, S.warp_sums, &n_keep);
        const int exm = block_excl_scan(merge, S.warp_sums, &n_merge);
        if (keep) {
            S.idA[exk] = S.idB[t];
            S.scA[exk] = S.scB[t];
        }
        if (merge) {
            S.idA[n_keep + exm] = p;
            S.scA[n_keep + exm] = mean;
        }
        n = n_keep + n_merge;
        changed = (n_ins > 0) || (n_merge > 0);
        ++rounds;
        __syncthreads();
    }

    // ================= stable sort by score, descending =================
    if (t < n) {
        const double s = S.scA[t];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double sj = S.scA[j];
            rank += (sj > s || (sj == s && j < t)) ? 1 : 0;
        }
        if (rank < a.max_out) {
            a.out_ids[size_t(b) * a.max_out + rank] = S.idA[t];
            a.out_scores[size_t(b) * a.max_out + rank] = s;
        }
    }
    for (int i = n + t; i < a.max_out; i += int(blockDim.x)) {
        a.out_ids[size_t(b) * a.max_out + i] = -1;
        a.out_scores[size_t(b) * a.max_out + i] = 0.0;
    }
    if (t == 0) a.out_len[b] = (S.bad || n > a.max_out) ? -1 : n;
}

}  // namespace tt
