/*
 * Oracle (C restatement, FAST mode) -- the timed CPU arm of bench.py (cpu_baseline /
 * --impl reference, kind "port").  TEST/BENCH INFRASTRUCTURE ONLY, parity unpinned (see
 * oracle/__init__.py); never linked into the product library.
 *
 * Same path as oracle_strict.c (brute-force form of the query issued at
 * /root/reference/src/tensortruth/rag_engine.py:628-639) but scored the way a CPU
 * vector store would: fp32 accumulation, SIMD, every host thread (OpenMP over rows),
 * corpus streamed once per query batch.  Ids can differ from the strict oracle on
 * near-ties; it is a throughput baseline, not the parity reference.
 *
 * Build: oracle/Makefile (gcc -O3 -ffast-math -mavx2 -mfma -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float key; int64_t id; } ent_t;

static inline int better(float ka, int64_t ia, float kb, int64_t ib) {
    return ka > kb || (ka == kb && ia < ib);
}

static void list_insert(ent_t* list, int* len, int k, float key, int64_t id) {
    if (*len == k && !better(key, id, list[k - 1].key, list[k - 1].id)) return;
    int pos = (*len < k) ? (*len)++ : k - 1;
    while (pos > 0 && better(key, id, list[pos - 1].key, list[pos - 1].id)) {
        list[pos] = list[pos - 1];
        --pos;
    }
    list[pos].key = key;
    list[pos].id = id;
}

int oracle_fast_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define QB 8 /* queries scored per pass over a row */

/* corpus: bf16 bits [n_rows, dim]; inv_norm: fp32 [n_rows]; q: fp32 [n_q, dim] (any norm).
 * Cosine only.  out_*: [n_q, k], padded with -1 / -inf. */
int oracle_fast_scan_topk(const uint16_t* corpus, const float* inv_norm, int64_t n_rows, int dim,
                          const float* q, int n_q, int k, int64_t id_base,
                          float* out_scores, int64_t* out_ids) {
    if (k <= 0 || dim <= 0 || dim % 8) return -1;
    int n_thr = oracle_fast_threads();
    ent_t* lists = (ent_t*)malloc(sizeof(ent_t) * (size_t)k * n_q * n_thr);
    int* lens = (int*)calloc((size_t)n_q * n_thr, sizeof(int));
    float* qinv = (float*)malloc(sizeof(float) * (size_t)n_q);
    for (int b = 0; b < n_q; ++b) {
        double qq = 0;
        for (int i = 0; i < dim; ++i) qq += (double)q[(size_t)b * dim + i] * q[(size_t)b * dim + i];
        qinv[b] = qq > 0 ? (float)(1.0 / sqrt(qq)) : 0.f;
    }
#pragma omp parallel num_threads(n_thr)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        float* rowf = (float*)malloc(sizeof(float) * (size_t)dim);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n_rows; ++r) {
            const uint16_t* row = corpus + (size_t)r * dim;
            for (int i = 0; i < dim; ++i) {
                uint32_t u = ((uint32_t)row[i]) << 16;
                memcpy(&rowf[i], &u, 4);
            }
            float inv = inv_norm[r];
            for (int b0 = 0; b0 < n_q; b0 += QB) {
                int nb = n_q - b0 < QB ? n_q - b0 : QB;
                float acc[QB] = {0};
                for (int j = 0; j < nb; ++j) {
                    const float* qb = q + (size_t)(b0 + j) * dim;
                    float a = 0.f;
#pragma omp simd reduction(+ : a)
                    for (int i = 0; i < dim; ++i) a += qb[i] * rowf[i];
                    acc[j] = a;
                }
                for (int j = 0; j < nb; ++j) {
                    int b = b0 + j;
                    list_insert(lists + ((size_t)tid * n_q + b) * k, &lens[(size_t)tid * n_q + b], k,
                                acc[j] * inv * qinv[b], r + id_base);
                }
            }
        }
        free(rowf);
    }
    ent_t* fin = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
    for (int b = 0; b < n_q; ++b) {
        int flen = 0;
        for (int t = 0; t < n_thr; ++t) {
            ent_t* l = lists + ((size_t)t * n_q + b) * k;
            for (int j = 0; j < lens[(size_t)t * n_q + b]; ++j) list_insert(fin, &flen, k, l[j].key, l[j].id);
        }
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)b * k + j] = j < flen ? fin[j].id : -1;
            out_scores[(size_t)b * k + j] = j < flen ? fin[j].key : -INFINITY;
        }
    }
    free(fin); free(lists); free(lens); free(qinv);
    return 0;
}
