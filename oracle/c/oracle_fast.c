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
 * The row x query dot products use AVX2/FMA intrinsics (bf16 -> fp32 widening in registers, two FMA chains
 * per query), so one thread streams the corpus at close to its share of memory bandwidth.
 *
 * Build: oracle/Makefile (gcc -O3 -ffast-math -mavx2 -mfma -fopenmp).
 */
#define _GNU_SOURCE
#include <math.h>
#include <sched.h>
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

/* The timed CPU arm is meant to use every host thread it can; launchers such as torchrun export OMP_NUM_THREADS=1
 * to their workers, which the OpenMP runtime has already read by the time this library loads. */
void oracle_fast_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#define QB 4 /* queries scored per pass over a row */

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
/* 8 bf16 -> 8 fp32 (exact): zero-extend to 32 bits, shift into the high half */
static inline __m256 bf16x8_to_f32(const uint16_t* p) {
    __m128i h = _mm_loadu_si128((const __m128i*)p);
    return _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(h), 16));
}
static inline float hsum8(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    return _mm_cvtss_f32(lo);
}
/* dots of one bf16 row with nb (<= QB) fp32 queries; two independent FMA chains per query */
static inline void row_dots(const uint16_t* row, const float* q, int dim, int nb, float* out) {
    __m256 acc[QB][2];
    for (int j = 0; j < QB; ++j) acc[j][0] = acc[j][1] = _mm256_setzero_ps();
    for (int i = 0; i < dim; i += 16) {
        __m256 r0 = bf16x8_to_f32(row + i), r1 = bf16x8_to_f32(row + i + 8);
        for (int j = 0; j < nb; ++j) {
            const float* qb = q + (size_t)j * dim + i;
            acc[j][0] = _mm256_fmadd_ps(_mm256_loadu_ps(qb), r0, acc[j][0]);
            acc[j][1] = _mm256_fmadd_ps(_mm256_loadu_ps(qb + 8), r1, acc[j][1]);
        }
    }
    for (int j = 0; j < nb; ++j) out[j] = hsum8(_mm256_add_ps(acc[j][0], acc[j][1]));
}
#define HAVE_ROW_DOTS 1
#endif

/* corpus: bf16 bits [n_rows, dim]; inv_norm: fp32 [n_rows]; q: fp32 [n_q, dim] (any norm).
 * Cosine only.  out_*: [n_q, k], padded with -1 / -inf. */
int oracle_fast_scan_topk(const uint16_t* corpus, const float* inv_norm, int64_t n_rows, int dim,
                          const float* q, int n_q, int k, int64_t id_base,
                          float* out_scores, int64_t* out_ids) {
    if (k <= 0 || dim <= 0 || dim % 16) return -1;
    int n_thr = oracle_fast_threads();
    ent_t* lists = (ent_t*)malloc(sizeof(ent_t) * (size_t)k * n_q * n_thr);
    int* lens = (int*)calloc((size_t)n_q * n_thr, sizeof(int));
    float* qinv = (float*)malloc(sizeof(float) * (size_t)n_q);
    for (int b = 0; b < n_q; ++b) {
        double qq = 0;
        for (int i = 0; i < dim; ++i) qq += (double)q[(size_t)b * dim + i] * q[(size_t)b * dim + i];
        qinv[b] = qq > 0 ? (float)(1.0 / sqrt(qq)) : 0.f;
    }
    /* One worker per allowed CPU, pinned for the duration of the scan: left to the scheduler the workers of
     * this memory-bound loop were seen to pile up on a few CPUs (8 threads slower than 1 on the build box). */
    cpu_set_t allowed;
    int cpus[CPU_SETSIZE], ncpu = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
#pragma omp parallel num_threads(n_thr)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        if (ncpu > 0) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[tid % ncpu], &one);
            sched_setaffinity(0, sizeof(one), &one);
        }
#ifndef HAVE_ROW_DOTS
        float* rowf = (float*)malloc(sizeof(float) * (size_t)dim);
#endif
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n_rows; ++r) {
            const uint16_t* row = corpus + (size_t)r * dim;
#ifndef HAVE_ROW_DOTS
            for (int i = 0; i < dim; ++i) {
                uint32_t u = ((uint32_t)row[i]) << 16;
                memcpy(&rowf[i], &u, 4);
            }
#endif
            float inv = inv_norm[r];
            for (int b0 = 0; b0 < n_q; b0 += QB) {
                int nb = n_q - b0 < QB ? n_q - b0 : QB;
                float acc[QB] = {0};
#ifdef HAVE_ROW_DOTS
                row_dots(row, q + (size_t)b0 * dim, dim, nb, acc);
#else
                for (int j = 0; j < nb; ++j) {
                    const float* qb = q + (size_t)(b0 + j) * dim;
                    float a = 0.f;
#pragma omp simd reduction(+ : a)
                    for (int i = 0; i < dim; ++i) a += qb[i] * rowf[i];
                    acc[j] = a;
                }
#endif
                for (int j = 0; j < nb; ++j) {
                    int b = b0 + j;
                    list_insert(lists + ((size_t)tid * n_q + b) * k, &lens[(size_t)tid * n_q + b], k,
                                acc[j] * inv * qinv[b], r + id_base);
                }
            }
        }
#ifndef HAVE_ROW_DOTS
        free(rowf);
#endif
        if (ncpu > 0) sched_setaffinity(0, sizeof(allowed), &allowed);
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
