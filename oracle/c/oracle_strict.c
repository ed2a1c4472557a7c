/*
 * Oracle (C restatement, strict mode) -- TEST INFRASTRUCTURE ONLY, parity unpinned
 * (see oracle/__init__.py).  Independent second statement of oracle/oracle.py and
 * oracle/automerge.py, used to cross-check the numpy/Python one and to score larger
 * samples in seconds.  Never linked into, or called from, the product library.
 *
 * Follows:
 *   scan target ... brute-force form of the Chroma query the reference issues at
 *                   /root/reference/src/tensortruth/rag_engine.py:628-639
 *   auto-merge .... llama_index AutoMergingRetriever as instantiated at
 *                   /root/reference/src/tensortruth/rag_engine.py:641-643
 *                   (_fill_in_nodes, _get_parents_and_merge, _try_merging, _retrieve)
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MODE_COSINE 0
#define MODE_CHROMA_L2_EXP 1

static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* key desc, id asc */
static inline int better(float ka, int64_t ia, float kb, int64_t ib) {
    if (ka > kb) return 1;
    if (ka < kb) return 0;
    return ia < ib;
}

typedef struct { float key; int64_t id; } ent_t;

static void list_insert(ent_t* list, int* len, int k, float key, int64_t id) {
    if (key != key) key = -INFINITY; /* NaN sorts last */
    if (*len == k && !better(key, id, list[k - 1].key, list[k - 1].id)) return;
    int pos = (*len < k) ? (*len)++ : k - 1;
    while (pos > 0 && better(key, id, list[pos - 1].key, list[pos - 1].id)) {
        list[pos] = list[pos - 1];
        --pos;
    }
    list[pos].key = key;
    list[pos].id = id;
}

static float key_of(double dot, double qq, double nn, int mode) {
    if (mode == MODE_COSINE) {
        double den = sqrt(qq) * sqrt(nn);
        return den > 0.0 ? (float)(dot / den) : 0.0f;
    }
    return -(float)(qq + nn - 2.0 * dot);
}

static float score_of(float key, int mode) {
    if (mode == MODE_COSINE) return key;
    return (float)exp((double)key); /* key = -d */
}

/* corpus_dtype: 0 = bf16 bits, 1 = fp32.  Outputs are [n_q, k], padded with id -1 / -inf. */
int oracle_scan_topk(const void* corpus, int corpus_dtype, int64_t n_rows, int dim, int64_t row_stride,
                     const float* q, int n_q, int k, int mode, int64_t id_base,
                     float* out_scores, float* out_keys, int64_t* out_ids) {
    if (k <= 0 || dim <= 0 || n_q < 0 || n_rows < 0) return -1;
    for (int b = 0; b < n_q; ++b) {
        const float* qb = q + (size_t)b * dim;
        double qq = 0.0;
        for (int i = 0; i < dim; ++i) qq += (double)qb[i] * (double)qb[i];
        int n_thr = 1;
        ent_t* lists = NULL;
        int* lens = NULL;
#pragma omp parallel
        {
#pragma omp single
            {
#ifdef _OPENMP
                n_thr = omp_get_num_threads();
#endif
                lists = (ent_t*)malloc(sizeof(ent_t) * (size_t)k * n_thr);
                lens = (int*)calloc(n_thr, sizeof(int));
            }
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            ent_t* mine = lists + (size_t)tid * k;
            int len = 0;
#pragma omp for schedule(static)
            for (int64_t r = 0; r < n_rows; ++r) {
                double dot = 0.0, nn = 0.0;
                if (corpus_dtype == 0) {
                    const uint16_t* row = (const uint16_t*)corpus + (size_t)r * row_stride;
                    for (int i = 0; i < dim; ++i) {
                        double c = (double)bf16_to_f32(row[i]);
                        dot += (double)qb[i] * c;
                        nn += c * c;
                    }
                } else {
                    const float* row = (const float*)corpus + (size_t)r * row_stride;
                    for (int i = 0; i < dim; ++i) {
                        double c = (double)row[i];
                        dot += (double)qb[i] * c;
                        nn += c * c;
                    }
                }
                list_insert(mine, &len, k, key_of(dot, qq, nn, mode), r + id_base);
            }
            lens[tid] = len;
        }
        ent_t* fin = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        int flen = 0;
        for (int t = 0; t < n_thr; ++t)
            for (int j = 0; j < lens[t]; ++j)
                list_insert(fin, &flen, k, lists[(size_t)t * k + j].key, lists[(size_t)t * k + j].id);
        for (int j = 0; j < k; ++j) {
            size_t o = (size_t)b * k + j;
            if (j < flen) {
                out_ids[o] = fin[j].id;
                out_keys[o] = fin[j].key;
                out_scores[o] = score_of(fin[j].key, mode);
            } else {
                out_ids[o] = -1;
                out_keys[o] = -INFINITY;
                out_scores[o] = -INFINITY;
            }
        }
        free(fin);
        free(lists);
        free(lens);
    }
    return 0;
}

/* ---- auto-merge ------------------------------------------------------------------- */

static int fill_in(int64_t** ids, double** sc, int* n, int* cap,
                   const int32_t* prev_id, const int32_t* next_id) {
    int m = *n, changed = 0, w = 0;
    int64_t* oi = (int64_t*)malloc(sizeof(int64_t) * (size_t)(2 * m + 1));
    double* os = (double*)malloc(sizeof(double) * (size_t)(2 * m + 1));
    for (int i = 0; i < m; ++i) {
        oi[w] = (*ids)[i];
        os[w] = (*sc)[i];
        ++w;
        if (i >= m - 1) continue;
        int32_t nxt = next_id[(*ids)[i]];
        if (nxt != -1 && nxt == prev_id[(*ids)[i + 1]]) {
            changed = 1;
            oi[w] = nxt;
            os[w] = ((*sc)[i] + (*sc)[i + 1]) / 2;
            ++w;
        }
    }
    free(*ids);
    free(*sc);
    *ids = oi;
    *sc = os;
    *n = w;
    *cap = 2 * m + 1;
    return changed;
}

static int merge_up(int64_t** ids, double** sc, int* n, const int32_t* parent_of,
                    const int32_t* child_count, double thresh) {
    int m = *n;
    int32_t* gp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1)); /* group parents, first-encounter order */
    int* gcnt = (int*)calloc((size_t)m + 1, sizeof(int));
    double* gsum = (double*)calloc((size_t)m + 1, sizeof(double));
    char* gmerge = (char*)calloc((size_t)m + 1, 1);
    int ng = 0;
    for (int i = 0; i < m; ++i) {
        int32_t p = parent_of[(*ids)[i]];
        if (p < 0) continue;
        int g = 0;
        while (g < ng && gp[g] != p) ++g;
        if (g == ng) { gp[ng++] = p; }
        gcnt[g] += 1;
        gsum[g] += (*sc)[i]; /* list order, like Python's sum() */
    }
    int changed = 0;
    for (int g = 0; g < ng; ++g) {
        int cc = child_count[gp[g]] > 0 ? child_count[gp[g]] : 1;
        double ratio = (double)gcnt[g] / (double)cc;
        if (ratio > thresh) { gmerge[g] = 1; changed = 1; }
    }
    int64_t* oi = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + ng + 1));
    double* os = (double*)malloc(sizeof(double) * (size_t)(m + ng + 1));
    int w = 0;
    for (int i = 0; i < m; ++i) {
        int32_t p = parent_of[(*ids)[i]];
        int drop = 0;
        if (p >= 0) {
            int g = 0;
            while (gp[g] != p) ++g;
            drop = gmerge[g];
        }
        if (!drop) { oi[w] = (*ids)[i]; os[w] = (*sc)[i]; ++w; }
    }
    for (int g = 0; g < ng; ++g)
        if (gmerge[g]) { oi[w] = gp[g]; os[w] = gsum[g] / (double)gcnt[g]; ++w; }
    free(*ids); free(*sc); free(gp); free(gcnt); free(gsum); free(gmerge);
    *ids = oi; *sc = os; *n = w;
    return changed;
}

/* Returns the merged list length (<= max_out), or -1 if it would not fit. */
int oracle_automerge(const int64_t* in_ids, const double* in_scores, int n_in,
                     const int32_t* parent_of, const int32_t* child_count,
                     const int32_t* prev_id, const int32_t* next_id,
                     double ratio_thresh, int max_rounds,
                     int64_t* out_ids, double* out_scores, int max_out) {
    int n = n_in, cap = n_in + 1;
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
    double* sc = (double*)malloc(sizeof(double) * (size_t)cap);
    memcpy(ids, in_ids, sizeof(int64_t) * (size_t)n);
    memcpy(sc, in_scores, sizeof(double) * (size_t)n);
    int rounds = 0, changed = 1;
    while (changed && rounds < max_rounds) {
        int c0 = fill_in(&ids, &sc, &n, &cap, prev_id, next_id);
        int c1 = merge_up(&ids, &sc, &n, parent_of, child_count, ratio_thresh);
        changed = c0 || c1;
        ++rounds;
    }
    /* stable sort by score descending (insertion sort keeps ties in list order) */
    for (int i = 1; i < n; ++i) {
        int64_t ti = ids[i];
        double ts = sc[i];
        int j = i;
        while (j > 0 && sc[j - 1] < ts) { ids[j] = ids[j - 1]; sc[j] = sc[j - 1]; --j; }
        ids[j] = ti;
        sc[j] = ts;
    }
    int ret = n;
    if (n > max_out) ret = -1;
    else { memcpy(out_ids, ids, sizeof(int64_t) * (size_t)n); memcpy(out_scores, sc, sizeof(double) * (size_t)n); }
    free(ids);
    free(sc);
    return ret;
}
