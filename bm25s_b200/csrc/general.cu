// Dense "general" path: one CTA per query scores into a float32[n_docs] row of HBM scratch in the reference's
// exact order (token by token, scoring.py:358-366), applies the weight mask (bm25s/__init__.py:610-612), selects
// the k largest by radix select (selection.py:14-37) and adds the non-occurrence shift (__init__.py:616-618).
// It serves the inputs the fused streaming kernel (retrieve_fast.cu) declines: weight masks, queries with more
// than 15 tokens (kMaxTok), negative scores in the matrix, very large k; and it backs get_scores / selection.topk.
#include "common.cuh"
#include "select.cuh"

namespace bm25cu {

constexpr int kGeneralThreads = 1024;

// scores[d] = sum over tokens (in order) of M[d, t]; the row must be zero on entry. Column rows are unique,
// so one token's adds never collide; __syncthreads between tokens keeps the reference's summation order.
// A token id outside [0, n_vocab) is skipped and reported through `err` bit 0 (the host checks ids too; device entry
// points get theirs from the caller).
__device__ void dense_score_row(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                const int64_t *__restrict__ indptr, int32_t n_vocab, const int32_t *__restrict__ q, int nq,
                                float *row, unsigned int *err) {
    for (int t = 0; t < nq; ++t) {
        const int tok = q[t];
        if ((unsigned)tok >= (unsigned)n_vocab) {  // uniform over the CTA
            if (threadIdx.x == 0 && err) atomicOr(err, 1u);
            continue;
        }
        const int64_t s = indptr[tok], e = indptr[tok + 1];
        for (int64_t j = s + threadIdx.x; j < e; j += blockDim.x) row[indices[j]] += data[j];
        __syncthreads();
    }
}

__device__ void zero_row(float *row, int64_t n) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) row[i] = 0.0f;
    __syncthreads();
}

__global__ void __launch_bounds__(kGeneralThreads)
scores_dense_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                    const int64_t *__restrict__ indptr, int32_t n_vocab, const int32_t *__restrict__ q, int nq, int64_t n_docs,
                    const double *__restrict__ mask, int has_shift, float shift, float *out) {
    zero_row(out, n_docs);
    dense_score_row(data, indices, indptr, n_vocab, q, nq, out, nullptr);
    for (int64_t d = threadIdx.x; d < n_docs; d += blockDim.x) {
        float s = out[d];
        if (mask) s = (float)__dmul_rn((double)s, mask[d]);
        if (has_shift) s = __fadd_rn(s, shift);
        out[d] = s;
    }
}

// One CTA drains the general queue: query list[c], list[c + gridDim.x], ...
__global__ void __launch_bounds__(kGeneralThreads)
retrieve_general_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                        const int64_t *__restrict__ indptr, int32_t n_vocab, int64_t n_docs, int32_t doc_base, RetrieveArgs a,
                        Ctrl *ctrl, const unsigned int *__restrict__ list, float *scratch, uint32_t *tmp_key,
                        int32_t *tmp_id, int P) {
    __shared__ unsigned int hist[256];
    __shared__ SelectState<uint32_t> st;
    __shared__ unsigned int warp_cnt[33];
    __shared__ unsigned int gt_counter;
    const unsigned int count = ctrl->general_count;
    float *row = scratch + (size_t)blockIdx.x * (size_t)n_docs;
    uint32_t *key = tmp_key + (size_t)blockIdx.x * P;
    int32_t *id = tmp_id + (size_t)blockIdx.x * P;
    const int k = a.k;
    for (unsigned int w = blockIdx.x; w < count; w += gridDim.x) {
        const unsigned int q = list[w];
        const int64_t qs = a.q_ptr[q], qe = a.q_ptr[q + 1];
        const int nq = (int)(qe - qs);
        if (threadIdx.x == 0) {
            unsigned long long postings = 0;
            for (int t = 0; t < nq; ++t) {
                const int tok = a.q_tokens[qs + t];
                if ((unsigned)tok < (unsigned)n_vocab) postings += (unsigned long long)(indptr[tok + 1] - indptr[tok]);
            }
            atomicAdd(&ctrl->posting_count, postings);
        }
        zero_row(row, n_docs);
        dense_score_row(data, indices, indptr, n_vocab, a.q_tokens + qs, nq, row, &ctrl->error);
        if (a.mask && nq > 0) {
            for (int64_t d = threadIdx.x; d < n_docs; d += blockDim.x)
                row[d] = (float)__dmul_rn((double)row[d], a.mask[d]);
            __syncthreads();
        } else if (a.mask_bits && nq > 0) {  // 0/1 mask as a bitmap: times one is exact, times zero is zero
            for (int64_t d = threadIdx.x; d < n_docs; d += blockDim.x)
                if (!((a.mask_bits[d >> 5] >> (d & 31)) & 1u)) row[d] = 0.0f;
            __syncthreads();
        }
        const float shift = (a.shift && nq > 0) ? a.shift[q] : 0.0f;  // empty query: zeros, no shift (:653-657)
        const bool has_shift = (a.shift != nullptr) && nq > 0;
        const int kk = (int)((int64_t)k < n_docs ? k : n_docs);
        auto load = [&](long long i) { return f32_key(row[i]); };
        if (kk > 0) {
            uint32_t thr;
            long long need;
            block_radix_select<uint32_t>(load, n_docs, kk, hist, &st, f32_key(0.0f), thr, need);
            for (int i = threadIdx.x; i < P; i += blockDim.x) { key[i] = 0u; id[i] = 0x7fffffff; }
            __syncthreads();
            block_collect<uint32_t, int32_t>(load, n_docs, kk, thr, need, key, id, warp_cnt, &gt_counter);
            block_bitonic_sort_desc<uint32_t, int32_t>(key, id, P);
        }
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            int32_t oid = -1;
            float os = __int_as_float(0xff800000);  // -inf
            if (j < kk) {
                int32_t d = id[j];
                float s = row[d];
                if (has_shift) s = __fadd_rn(s, shift);
                oid = d + doc_base;
                os = s;
            }
            a.out_ids[(size_t)q * k + j] = oid;
            a.out_scores[(size_t)q * k + j] = os;
        }
        __syncthreads();
    }
}

int launch_scores_dense(bm25cu_index *ix, const int32_t *d_q, int32_t n_tokens, int has_shift, float shift,
                        const double *d_mask, float *d_out) {
    scores_dense_kernel<<<1, kGeneralThreads, 0, ix->stream>>>(ix->d_data, ix->d_indices, ix->d_indptr, ix->n_vocab, d_q, n_tokens,
                                                               ix->n_docs, d_mask, has_shift, shift, d_out);
    BM25CU_CHECK(cudaGetLastError());
    return BM25CU_OK;
}

int launch_retrieve_general(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, const unsigned int *d_general_list,
                            int *launches) {
    // scratch rows: bounded by ~1 GiB, by the SM count and by the number of queries
    const int64_t n = ix->n_docs > 0 ? ix->n_docs : 1;
    int rows = (int)((int64_t(1) << 30) / (n * 4));
    if (rows > ix->sm_count) rows = ix->sm_count;
    if (rows > a.n_queries) rows = a.n_queries;
    if (rows < 1) rows = 1;
    const int P = next_pow2(a.k < 1 ? 1 : (a.k > ix->n_docs ? (ix->n_docs > 0 ? ix->n_docs : 1) : a.k));
    int rc = ix->d_scratch.reserve((size_t)rows * n * 4 + (size_t)rows * P * 8);
    if (rc) return rc;
    float *scratch = ix->d_scratch.as<float>();
    uint32_t *tmp_key = reinterpret_cast<uint32_t *>(scratch + (size_t)rows * n);
    int32_t *tmp_id = reinterpret_cast<int32_t *>(tmp_key + (size_t)rows * P);
    retrieve_general_kernel<<<rows, kGeneralThreads, 0, ix->stream>>>(ix->d_data, ix->d_indices, ix->d_indptr, ix->n_vocab,
                                                                      ix->n_docs, ix->doc_base, a, d_ctrl,
                                                                      d_general_list, scratch, tmp_key, tmp_id, P);
    BM25CU_CHECK(cudaGetLastError());
    if (launches) ++*launches;
    return BM25CU_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Standalone top-k of a dense vector (selection.topk replacement, bm25s/selection.py:48-64)
// ---------------------------------------------------------------------------------------------------------
template <typename T, typename Key>
__device__ __forceinline__ Key key_of(T v);
template <>
__device__ __forceinline__ uint32_t key_of<float, uint32_t>(float v) { return f32_key(v); }
template <>
__device__ __forceinline__ uint64_t key_of<double, uint64_t>(double v) { return f64_key(v); }

template <typename T, typename Key>
__global__ void __launch_bounds__(kGeneralThreads)
topk_dense_kernel(const T *__restrict__ scores, long long n, int k, Key *tmp_key, long long *tmp_id, int P,
                  long long *out_ids, T *out_scores) {
    __shared__ unsigned int hist[256];
    __shared__ SelectState<Key> st;
    __shared__ unsigned int warp_cnt[33];
    __shared__ unsigned int gt_counter;
    auto load = [&](long long i) { return key_of<T, Key>(scores[i]); };
    Key thr;
    long long need;
    block_radix_select<Key>(load, n, k, hist, &st, key_of<T, Key>((T)0), thr, need);
    for (int i = threadIdx.x; i < P; i += blockDim.x) { tmp_key[i] = 0; tmp_id[i] = 0x7fffffffffffffffll; }
    __syncthreads();
    block_collect<Key, long long>(load, n, k, thr, need, tmp_key, tmp_id, warp_cnt, &gt_counter);
    block_bitonic_sort_desc<Key, long long>(tmp_key, tmp_id, P);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        long long d = tmp_id[j];
        out_ids[j] = d;
        out_scores[j] = scores[d];
    }
}

int topk_dense_device(const void *d_scores, int is_f64, int64_t n, int32_t k, int64_t *d_out_ids, void *d_out_scores,
                      cudaStream_t stream) {
    const int P = next_pow2(k);
    void *tmp = nullptr;
    BM25CU_CHECK(cudaMallocAsync(&tmp, (size_t)P * 16, stream));
    if (is_f64) {
        uint64_t *tk = (uint64_t *)tmp;
        long long *ti = (long long *)(tk + P);
        topk_dense_kernel<double, uint64_t><<<1, kGeneralThreads, 0, stream>>>(
            (const double *)d_scores, n, k, tk, ti, P, (long long *)d_out_ids, (double *)d_out_scores);
    } else {
        long long *ti = (long long *)tmp;
        uint32_t *tk = (uint32_t *)(ti + P);
        topk_dense_kernel<float, uint32_t><<<1, kGeneralThreads, 0, stream>>>(
            (const float *)d_scores, n, k, tk, ti, P, (long long *)d_out_ids, (float *)d_out_scores);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tmp, stream);
    BM25CU_CHECK(e);
    return BM25CU_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Merge of per-shard top-k lists (doc-range sharding): one CTA per query sorts n_parts * k pairs.
// ---------------------------------------------------------------------------------------------------------
// Small n_parts * k: one CTA bitonic-sorts all pairs.
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int32_t *__restrict__ ids, const float *__restrict__ scores, int n_parts, int n_queries, int k,
                  int P, const float *__restrict__ shift, int32_t *out_ids, float *out_scores) {
    extern __shared__ unsigned char smem_raw[];
    uint32_t *key = reinterpret_cast<uint32_t *>(smem_raw);
    int32_t *id = reinterpret_cast<int32_t *>(key + P);
    const int q = blockIdx.x;
    const int total = n_parts * k;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t kk = 0u;
        int32_t ii = 0x7fffffff;
        if (i < total) {
            const int p = i / k, j = i - p * k;
            const size_t src = ((size_t)p * n_queries + q) * k + j;
            int32_t d = ids[src];
            if (d >= 0) { kk = f32_key(scores[src]); ii = d; }
        }
        key[i] = kk;
        id[i] = ii;
    }
    block_bitonic_sort_desc<uint32_t, int32_t>(key, id, P);
    // scores are recovered from the keys (the key transform is a bijection on non-NaN floats)
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        uint32_t kk = key[j];
        int32_t d = id[j];
        float s;
        if (d == 0x7fffffff) { d = -1; s = __int_as_float(0xff800000); }
        else {
            uint32_t u = (kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk;
            s = __uint_as_float(u);
            if (shift != nullptr) s = __fadd_rn(s, shift[q]);  // BM25L / BM25+: after the merge, so the order does not depend on n_parts
        }
        out_ids[(size_t)q * k + j] = d;
        out_scores[(size_t)q * k + j] = s;
    }
}

// Larger k: the parts' rows (sorted by their producers) are rank-merged one after the other (select.cuh,
// block_merge_sorted_rows); same total order - score descending, id ascending.
__global__ void __launch_bounds__(256)
topk_merge_rows_kernel(const int32_t *__restrict__ ids, const float *__restrict__ scores, int n_parts, int n_queries, int k,
                       int kcap, const float *__restrict__ shift, int32_t *out_ids, float *out_scores) {
    extern __shared__ __align__(16) unsigned long long tm[];  // cur, nxt, row: kcap keys each
    const int q = blockIdx.x;
    auto load = [&](int r, int j) -> unsigned long long {
        const size_t src = ((size_t)r * n_queries + q) * k + j;
        const int32_t d = ids[src];
        return d >= 0 ? (((unsigned long long)f32_key(scores[src]) << 32) | (unsigned int)(~d)) : 0ull;
    };
    const unsigned long long *top = block_merge_sorted_rows(n_parts, k, kcap, tm, tm + kcap, tm + 2 * kcap, load);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long t = top[j];
        int32_t d = -1;
        float s = __int_as_float(0xff800000);
        if (t != 0ull) {
            const uint32_t kk = (uint32_t)(t >> 32);
            s = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
            if (shift != nullptr) s = __fadd_rn(s, shift[q]);
            d = (int32_t)~(unsigned int)(t & 0xffffffffu);
        }
        out_ids[(size_t)q * k + j] = d;
        out_scores[(size_t)q * k + j] = s;
    }
}

int topk_merge_device(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                      const float *shift, int32_t *out_ids, float *out_scores, cudaStream_t stream) {
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    const int kcap = next_pow2(k);
    if (n_parts * k > 256 && (size_t)3 * kcap * 8 <= 200 * 1024) {
        const size_t smem = (size_t)3 * kcap * 8;
        BM25CU_CHECK(cudaFuncSetAttribute(topk_merge_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_merge_rows_kernel<<<n_queries, 256, smem, stream>>>(ids, scores, n_parts, n_queries, k, kcap, shift, out_ids, out_scores);
        BM25CU_CHECK(cudaGetLastError());
        return BM25CU_OK;
    }
    const int P = next_pow2(n_parts * k);
    const size_t smem = (size_t)P * 8;
    BM25CU_REQUIRE(smem <= 200 * 1024, "bm25cu_topk_merge: n_parts * k too large for one CTA (max 25600 pairs)");
    BM25CU_CHECK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<n_queries, 256, smem, stream>>>(ids, scores, n_parts, n_queries, k, P, shift, out_ids, out_scores);
    BM25CU_CHECK(cudaGetLastError());
    return BM25CU_OK;
}

}  // namespace bm25cu
