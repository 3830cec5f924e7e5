/*
 * CPU oracle for the bm25s query-time hot path, plain C (TEST INFRASTRUCTURE ONLY).
 *
 * A restatement of the reference algorithm, used (a) by tests/ as a fast checker at sizes where the numpy
 * oracle (oracle/oracle.py) is slow, and (b) by bench.py as the `cpu_baseline` / `--impl reference` arm
 * (kind "port"), threaded over queries the way the reference's numba backend is
 * (bm25s/numba/retrieve_utils.py:33, prange over queries). The product never links or loads this file.
 *
 * Parity status: PINNED - tests/test_oracle.py checks it against the golden fixtures generated from the
 * reference itself (oracle/make_golden.py) and against oracle/oracle.py.
 *
 * Reference lines followed (relative to the reference checkout):
 *   dense scorer ........ bm25s/scoring.py:345-368 (_compute_relevance_from_scores_jit_ready): zeros(N) float32,
 *                         then for each query token in order, scores[indices[j]] += data[j]
 *   weight mask ......... bm25s/__init__.py:610-612 (scores *= weight_mask, numpy order: BEFORE the shift)
 *   non-occurrence shift  bm25s/__init__.py:616-618 (scores += nonoccurrence_array[ids].sum()); the sum itself is
 *                         numpy's reduction and is passed in per query by the Python side
 *   empty query ......... bm25s/__init__.py:653-657 (all-zero scores)
 *   top-k ............... bm25s/selection.py:14-37 (_topk_numpy): partition the index vector around the k-th
 *                         largest score (np.argpartition == introselect), take the k last, sort descending
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* scoring.py:358-366 */
static void score_dense(const float *data, const int32_t *indices, const int64_t *indptr, int64_t n_docs,
                        const int32_t *q, int64_t nq, float *scores) {
    memset(scores, 0, (size_t)n_docs * sizeof(float));
    for (int64_t i = 0; i < nq; i++) {
        int64_t s = indptr[q[i]], e = indptr[q[i] + 1];
        for (int64_t j = s; j < e; j++) scores[indices[j]] += data[j];
    }
}

/* __init__.py:610-618. mask is the weight mask converted exactly to float64 (bool/int/float32/float64 all are):
 * float32 * mask evaluated in the promoted type and rounded back to float32 equals (float)((double)s * m). */
static void mask_shift(float *scores, int64_t n_docs, const double *mask, int has_shift, float shift) {
    if (mask)
        for (int64_t d = 0; d < n_docs; d++) scores[d] = (float)((double)scores[d] * mask[d]);
    if (has_shift)
        for (int64_t d = 0; d < n_docs; d++) scores[d] += shift;
}

int oracle_scores_dense(const float *data, const int32_t *indices, const int64_t *indptr, int64_t n_docs,
                        const int32_t *q, int64_t nq, const double *mask, int has_shift, float shift, float *out) {
    score_dense(data, indices, indptr, n_docs, q, nq, out);
    mask_shift(out, n_docs, mask, has_shift, shift);
    return 0;
}

/* selection.py:17: np.argpartition(scores, -k) -> quickselect on an index vector (median-of-3, 3-way partition). */
static void select_kth(const float *s, int64_t *idx, int64_t n, int64_t kth) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        float a = s[idx[lo]], b = s[idx[mid]], c = s[idx[hi]];
        float p = (a < b) ? ((b < c) ? b : ((a < c) ? c : a)) : ((a < c) ? a : ((b < c) ? c : b));
        int64_t lt = lo, i = lo, gt = hi;
        while (i <= gt) {
            float v = s[idx[i]];
            if (v < p) { int64_t t = idx[lt]; idx[lt] = idx[i]; idx[i] = t; lt++; i++; }
            else if (v > p) { int64_t t = idx[gt]; idx[gt] = idx[i]; idx[i] = t; gt--; }
            else i++;
        }
        if (kth < lt) hi = lt - 1;
        else if (kth > gt) lo = gt + 1;
        else return;
    }
}

typedef struct { float s; int64_t i; } pair_t;
static int cmp_desc(const void *a, const void *b) {
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i < y->i) ? -1 : (x->i > y->i);
}

/* selection.py:14-37 on one dense vector. scratch_idx has n entries. */
static void topk_dense(const float *scores, int64_t n, int64_t k, int sorted, int64_t *scratch_idx,
                       int64_t *out_ids, float *out_scores) {
    if (k <= 0) return;
    for (int64_t i = 0; i < n; i++) scratch_idx[i] = i;
    select_kth(scores, scratch_idx, n, n - k);
    pair_t *p = (pair_t *)malloc((size_t)k * sizeof(pair_t));
    for (int64_t j = 0; j < k; j++) { p[j].i = scratch_idx[n - k + j]; p[j].s = scores[p[j].i]; }
    if (sorted) qsort(p, (size_t)k, sizeof(pair_t), cmp_desc);
    for (int64_t j = 0; j < k; j++) { out_ids[j] = p[j].i; out_scores[j] = p[j].s; }
    free(p);
}

int oracle_topk(const float *scores, int64_t n, int64_t k, int sorted, int64_t *out_ids, float *out_scores) {
    int64_t *idx = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    if (!idx) return 4;
    topk_dense(scores, n, k, sorted, idx, out_ids, out_scores);
    free(idx);
    return 0;
}

/* Batch retrieve: __init__.py:889-917 (per-query _get_top_k_results, then stack to (Q, k)).
 * Threads pull queries from a shared counter (the reference's numba backend uses prange over queries). */
typedef struct {
    const float *data; const int32_t *indices; const int64_t *indptr; int64_t n_docs;
    const int32_t *q_flat; const int64_t *q_ptr; int64_t n_queries, k; int sorted;
    const float *shift; const double *mask; int64_t *out_ids; float *out_scores;
    atomic_llong next; atomic_int err;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    int64_t n = j->n_docs > 0 ? j->n_docs : 1;
    float *scores = (float *)malloc((size_t)n * sizeof(float));
    int64_t *idx = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    if (!scores || !idx) { atomic_store(&j->err, 4); free(scores); free(idx); return NULL; }
    for (;;) {
        int64_t qi = (int64_t)atomic_fetch_add(&j->next, 1);
        if (qi >= j->n_queries) break;
        int64_t nq = j->q_ptr[qi + 1] - j->q_ptr[qi];
        if (nq == 0) {
            memset(scores, 0, (size_t)j->n_docs * sizeof(float)); /* __init__.py:653-657 */
        } else {
            score_dense(j->data, j->indices, j->indptr, j->n_docs, j->q_flat + j->q_ptr[qi], nq, scores);
            mask_shift(scores, j->n_docs, j->mask, j->shift != NULL, j->shift ? j->shift[qi] : 0.0f);
        }
        topk_dense(scores, j->n_docs, j->k, j->sorted, idx, j->out_ids + qi * j->k, j->out_scores + qi * j->k);
    }
    free(scores);
    free(idx);
    return NULL;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_retrieve(const float *data, const int32_t *indices, const int64_t *indptr, int64_t n_docs,
                    const int32_t *q_flat, const int64_t *q_ptr, int64_t n_queries, int64_t k, int sorted,
                    const float *shift_per_query, const double *mask, int64_t *out_ids, float *out_scores,
                    int n_threads) {
    job_t j = {data, indices, indptr, n_docs, q_flat, q_ptr, n_queries, k, sorted,
               shift_per_query, mask, out_ids, out_scores, 0, 0};
    if (n_threads <= 0) n_threads = oracle_max_threads();
    if (n_threads > n_queries) n_threads = (int)(n_queries > 0 ? n_queries : 1);
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < n_threads; t++)
        if (pthread_create(&th[started], NULL, worker, &j) == 0) started++;
    worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return atomic_load(&j.err);
}
