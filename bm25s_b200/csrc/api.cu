// extern "C" entry points of the query path (declared in include/bm25cu.h).
#include <vector>

#include "common.cuh"

using namespace bm25cu;

namespace {

// K0: pull order of the batch - queries bucketed by log2 of their posting count, longest first, so that the persistent
// CTAs of K1 do not end on a long query (longest-processing-time-first; the order inside a bucket is arbitrary and does
// not affect results). One CTA; `bucket` is n_queries bytes of scratch behind the order array.
__global__ void __launch_bounds__(1024) order_queries_kernel(const int32_t *__restrict__ q_tokens, const int64_t *__restrict__ q_ptr,
                                                             int32_t n_queries, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                                             unsigned int *order, unsigned char *bucket) {
    __shared__ unsigned int hist[64], base[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int q = threadIdx.x; q < n_queries; q += blockDim.x) {
        unsigned long long cost = 0;
        for (int64_t i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
            const int tok = q_tokens[i];
            if ((unsigned)tok < (unsigned)n_vocab) cost += (unsigned long long)(indptr[tok + 1] - indptr[tok]);
        }
        const int bk = cost ? 64 - __clzll((long long)cost) : 0;  // 0 .. 63
        bucket[q] = (unsigned char)bk;
        atomicAdd(&hist[bk], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int bk = 63; bk >= 0; --bk) { base[bk] = acc; acc += hist[bk]; }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < n_queries; q += blockDim.x) order[atomicAdd(&base[bucket[q]], 1u)] = (unsigned int)q;
}

// A weight mask with all values in [0, 1] only lowers scores, so the pruning bounds of K1 stay valid and the mask is
// applied where a candidate is emitted; anything else (negative, > 1, NaN) sends the batch to the dense kernel.
__global__ void mask_check_kernel(const double *__restrict__ mask, int64_t n, Ctrl *ctrl) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = mask[i];
        if (!(m >= 0.0 && m <= 1.0)) bad = true;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) ctrl->mask_bad = 1u;
}

// The weight mask of a call: the argument if given, else the handle's resident mask (bm25cu_index_set_mask)
void resolve_mask(const bm25cu_index *ix, const double *call_mask_device, RetrieveArgs &a) {
    a.mask = call_mask_device;
    a.mask_bits = nullptr;
    if (call_mask_device == nullptr) {
        if (ix->rmask_kind == BM25CU_MASK_F64) a.mask = ix->d_rmask.as<double>();
        else if (ix->rmask_kind == BM25CU_MASK_BITS) a.mask_bits = ix->d_rmask.as<unsigned int>();
    }
}

int run_retrieve_on_device(bm25cu_index *ix, const RetrieveArgs &a, bm25cu_stats *stats, int64_t n_query_tokens,
                           bool time_total_already_started) {
    (void)time_total_already_started;
    int rc;
    if ((rc = ix->d_ctrl.reserve(sizeof(Ctrl)))) return rc;
    if ((rc = ix->d_general_list.reserve((size_t)(a.n_queries > 0 ? a.n_queries : 1) * sizeof(unsigned int)))) return rc;
    Ctrl *d_ctrl = ix->d_ctrl.as<Ctrl>();
    BM25CU_CHECK(cudaMemsetAsync(d_ctrl, 0, sizeof(Ctrl), ix->stream));
    int launches = 0;
    BM25CU_CHECK(cudaEventRecord(ix->ev[1], ix->stream));
    if (a.mask && ix->n_docs > 0) {
        mask_check_kernel<<<ix->sm_count * 2, 512, 0, ix->stream>>>(a.mask, ix->n_docs, d_ctrl);
        BM25CU_CHECK(cudaGetLastError());
        ++launches;
    }
    // Knob MS (default 1): queries go to the list-at-a-time kernel first (retrieve_ms.cu: plan, parts, merge); what it
    // hands on is served by the streaming kernel from a device-side list. A batch that kernel cannot take at all (k over
    // its top-list capacity, an index with negative scores, FORCE_GENERAL) skips it: no plan, no partial-result buffer.
    const bool ms = ix->knobs.get("MS", 1) != 0 && retrieve_ms_takes(ix, a);
    ix->ms_ran = false;
    ix->order_valid = false;
    if (ms) {
        if ((rc = ix->d_fb_list.reserve((size_t)a.n_queries * sizeof(unsigned int)))) return rc;
        if ((rc = launch_retrieve_ms(ix, a, d_ctrl, ix->d_fb_list.as<unsigned int>(), &launches))) return rc;
        BM25CU_CHECK(cudaEventRecord(ix->ev[5], ix->stream));
        ix->ms_ran = true;
        if ((rc = launch_retrieve_fast(ix, a, d_ctrl, ix->d_general_list.as<unsigned int>(), &launches,
                                       ix->d_fb_list.as<unsigned int>(), &d_ctrl->fb_count)))
            return rc;
    } else {
        if (a.n_queries >= 4 * ix->sm_count && ix->knobs.get("ORDER", 1) != 0) {  // small batches: every CTA gets at most a few queries
            if ((rc = ix->d_order.reserve((size_t)a.n_queries * 5))) return rc;
            order_queries_kernel<<<1, 1024, 0, ix->stream>>>(a.q_tokens, a.q_ptr, a.n_queries, ix->d_indptr, ix->n_vocab,
                                                             ix->d_order.as<unsigned int>(),
                                                             ix->d_order.as<unsigned char>() + (size_t)a.n_queries * 4);
            BM25CU_CHECK(cudaGetLastError());
            ix->order_valid = true;
            ++launches;
        }
        if ((rc = launch_retrieve_fast(ix, a, d_ctrl, ix->d_general_list.as<unsigned int>(), &launches))) return rc;
    }
    BM25CU_CHECK(cudaEventRecord(ix->ev[2], ix->stream));
    if ((rc = launch_retrieve_general(ix, a, d_ctrl, ix->d_general_list.as<unsigned int>(), &launches))) return rc;
    BM25CU_CHECK(cudaEventRecord(ix->ev[3], ix->stream));
    ix->last_launches = launches;
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->launches = launches;
    }
    return BM25CU_OK;
}

int finish_stats(bm25cu_index *ix, const RetrieveArgs &a, bm25cu_stats *stats, int64_t n_query_tokens, bool have_total) {
    Ctrl h;
    BM25CU_CHECK(cudaMemcpyAsync(&h, ix->d_ctrl.as<Ctrl>(), sizeof(Ctrl), cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    if (h.error & 1u) {
        set_error("The maximum token ID in the query is higher than the number of tokens in the index.");
        return BM25CU_EINVAL;
    }
    if (h.error & 8u) {
        set_error("internal error: a bounds check of the checked build failed in retrieve_ms.cu");
        return BM25CU_ECUDA;
    }
    if (h.error & 4u) {
        set_error("internal error: candidate buffer overflow in the fused kernel");
        return BM25CU_ECUDA;
    }
    if (h.error & 2u) {
        set_error("internal error: search did not converge in the fused kernel");
        return BM25CU_ECUDA;
    }
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ix->ev[1], ix->ev[3]);
        stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, ix->ev[1], ix->ev[2]);
        stats->fast_kernel_ms = ms;
        if (have_total) {
            cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[4]);
            stats->total_ms = ms;
        } else {
            stats->total_ms = stats->kernel_ms;
        }
        stats->posting_bytes = (int64_t)h.posting_count * 8;
        stats->algorithmic_bytes = stats->posting_bytes + 20 * n_query_tokens + (int64_t)8 * a.k * a.n_queries;
        stats->n_general = (int32_t)h.general_count;
        stats->n_fast = a.n_queries - (int32_t)h.general_count;
        if (ix->ms_ran) {
            cudaEventElapsedTime(&ms, ix->ev[1], ix->ev[5]);
            stats->ms_kernel_ms = ms;
        }
    }
    return BM25CU_OK;
}

}  // namespace

extern "C" {

int bm25cu_index_set_option(bm25cu_index *ix, const char *name, int32_t value) {
    BM25CU_REQUIRE(ix != nullptr && name != nullptr && *name, "bm25cu_index_set_option: null handle or name");
    ix->knobs.set(name, value);
    return BM25CU_OK;
}

int bm25cu_index_set_mask(bm25cu_index *ix, int32_t kind, const void *mask_host) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_index_set_mask: null handle");
    BM25CU_REQUIRE(kind == BM25CU_MASK_NONE || kind == BM25CU_MASK_BITS || kind == BM25CU_MASK_F64, "bm25cu_index_set_mask: unknown kind");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));  // no call in flight reads the old mask
    ix->rmask_kind = BM25CU_MASK_NONE;
    if (kind == BM25CU_MASK_NONE) return BM25CU_OK;
    BM25CU_REQUIRE(mask_host != nullptr, "bm25cu_index_set_mask: null mask");
    const size_t n = (size_t)(ix->n_docs > 0 ? ix->n_docs : 1);
    const size_t bytes = kind == BM25CU_MASK_BITS ? ((n + 31) / 32) * 4 : n * sizeof(double);
    int rc = ix->d_rmask.reserve(bytes);
    if (rc) return rc;
    BM25CU_CHECK(cudaMemcpyAsync(ix->d_rmask.p, mask_host, bytes, cudaMemcpyHostToDevice, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    ix->rmask_kind = kind;
    return BM25CU_OK;
}

int bm25cu_index_unset_option(bm25cu_index *ix, const char *name) {
    BM25CU_REQUIRE(ix != nullptr && name != nullptr && *name, "bm25cu_index_unset_option: null handle or name");
    ix->knobs.unset(name);
    return BM25CU_OK;
}

// diagnostics (declared in include/bm25cu.h): the per-column 10th / 100th / 1000th largest scores, f32[3][n_vocab]
int bm25cu_debug_colkth(bm25cu_index *ix, float *out) {
    BM25CU_REQUIRE(ix && out && ix->d_colkth, "bm25cu_debug_colkth: no table");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    BM25CU_CHECK(cudaMemcpy(out, ix->d_colkth, (size_t)3 * ix->n_vocab * sizeof(float), cudaMemcpyDeviceToHost));
    return BM25CU_OK;
}

// diagnostics: the phase counters of the last retrieve call (48 x uint64); zeros unless built with -DBM25CU_PROFILE
int bm25cu_debug_prof(bm25cu_index *ix, unsigned long long *out) {
    BM25CU_REQUIRE(ix && out && ix->d_ctrl.p, "bm25cu_debug_prof: no retrieve call yet");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    BM25CU_CHECK(cudaMemcpy(out, ix->d_ctrl.as<Ctrl>()->prof, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return BM25CU_OK;
}

int bm25cu_retrieve(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries, int32_t k,
                    int32_t sorted, const float *shift_per_query, const double *weight_mask, int32_t *out_ids,
                    float *out_scores, bm25cu_stats *stats) {
    (void)sorted;  // rows are always returned sorted (a valid "arbitrary" order for sorted=False, selection.py:33-35)
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_retrieve: null handle");
    BM25CU_REQUIRE(n_queries >= 0 && k >= 0, "bm25cu_retrieve: negative n_queries or k");
    BM25CU_REQUIRE(q_ptr != nullptr, "bm25cu_retrieve: null q_ptr");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(out_ids != nullptr && out_scores != nullptr, "bm25cu_retrieve: null output");
    const int64_t n_tok = q_ptr[n_queries];
    BM25CU_REQUIRE(q_ptr[0] == 0 && n_tok >= 0, "bm25cu_retrieve: q_ptr must start at 0");
    BM25CU_REQUIRE(n_tok == 0 || q_tokens != nullptr, "bm25cu_retrieve: null q_tokens");
    for (int32_t i = 0; i < n_queries; ++i)
        BM25CU_REQUIRE(q_ptr[i + 1] >= q_ptr[i], "bm25cu_retrieve: q_ptr must be non-decreasing");
    for (int64_t i = 0; i < n_tok; ++i) {
        if (q_tokens[i] < 0 || q_tokens[i] >= ix->n_vocab) {
            char buf[256];
            snprintf(buf, sizeof(buf),
                     "The maximum token ID in the query (%d) is higher than the number of tokens in the index.",
                     q_tokens[i]);
            set_error(buf);
            return BM25CU_EINVAL;
        }
    }
    BM25CU_CHECK(cudaSetDevice(ix->device));
    int rc;
    const size_t tok_bytes = (size_t)(n_tok > 0 ? n_tok : 1) * sizeof(int32_t);
    const size_t ptr_bytes = (size_t)(n_queries + 1) * sizeof(int64_t);
    const size_t shift_bytes = shift_per_query ? (size_t)n_queries * sizeof(float) : 0;
    const size_t out_elems = (size_t)n_queries * k;
    // pinned staging: [ptr | tokens | shift]
    const size_t off_tok = ptr_bytes, off_shift = off_tok + ((tok_bytes + 15) / 16) * 16;
    if ((rc = ix->h_in.reserve(off_shift + shift_bytes + 16))) return rc;
    if ((rc = ix->h_out.reserve(out_elems * 8))) return rc;
    if ((rc = ix->d_qtok.reserve(tok_bytes))) return rc;
    if ((rc = ix->d_qptr.reserve(ptr_bytes))) return rc;
    if ((rc = ix->d_shift.reserve(shift_bytes + 4))) return rc;
    if ((rc = ix->d_out_ids.reserve(out_elems * 4))) return rc;
    if ((rc = ix->d_out_scores.reserve(out_elems * 4))) return rc;
    if (weight_mask && (rc = ix->d_mask.reserve((size_t)(ix->n_docs > 0 ? ix->n_docs : 1) * sizeof(double)))) return rc;
    char *hin = ix->h_in.as<char>();
    std::memcpy(hin, q_ptr, ptr_bytes);
    if (n_tok) std::memcpy(hin + off_tok, q_tokens, (size_t)n_tok * sizeof(int32_t));
    if (shift_per_query) std::memcpy(hin + off_shift, shift_per_query, shift_bytes);

    BM25CU_CHECK(cudaEventRecord(ix->ev[0], ix->stream));
    BM25CU_CHECK(cudaMemcpyAsync(ix->d_qptr.p, hin, ptr_bytes, cudaMemcpyHostToDevice, ix->stream));
    if (n_tok) BM25CU_CHECK(cudaMemcpyAsync(ix->d_qtok.p, hin + off_tok, (size_t)n_tok * sizeof(int32_t), cudaMemcpyHostToDevice, ix->stream));
    if (shift_per_query) BM25CU_CHECK(cudaMemcpyAsync(ix->d_shift.p, hin + off_shift, shift_bytes, cudaMemcpyHostToDevice, ix->stream));
    if (weight_mask) BM25CU_CHECK(cudaMemcpyAsync(ix->d_mask.p, weight_mask, (size_t)ix->n_docs * sizeof(double), cudaMemcpyHostToDevice, ix->stream));

    RetrieveArgs a;
    a.q_tokens = ix->d_qtok.as<int32_t>();
    a.q_ptr = ix->d_qptr.as<int64_t>();
    a.n_queries = n_queries;
    a.k = k;
    a.shift = shift_per_query ? ix->d_shift.as<float>() : nullptr;
    resolve_mask(ix, weight_mask ? ix->d_mask.as<double>() : nullptr, a);
    a.out_ids = ix->d_out_ids.as<int32_t>();
    a.out_scores = ix->d_out_scores.as<float>();
    if ((rc = run_retrieve_on_device(ix, a, stats, n_tok, true))) return rc;
    char *hout = ix->h_out.as<char>();
    BM25CU_CHECK(cudaMemcpyAsync(hout, a.out_ids, out_elems * 4, cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaMemcpyAsync(hout + out_elems * 4, a.out_scores, out_elems * 4, cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaEventRecord(ix->ev[4], ix->stream));
    if ((rc = finish_stats(ix, a, stats, n_tok, true))) return rc;
    std::memcpy(out_ids, hout, out_elems * 4);
    std::memcpy(out_scores, hout + out_elems * 4, out_elems * 4);
    return BM25CU_OK;
}

int bm25cu_retrieve_device(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries,
                           int64_t n_query_tokens, int32_t k, const float *shift_per_query,
                           const double *weight_mask, int32_t *out_ids, float *out_scores, bm25cu_stats *stats) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_retrieve_device: null handle");
    BM25CU_REQUIRE(n_queries >= 0 && k >= 0, "bm25cu_retrieve_device: negative n_queries or k");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(q_ptr && out_ids && out_scores, "bm25cu_retrieve_device: null pointer");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    RetrieveArgs a;
    a.q_tokens = q_tokens;
    a.q_ptr = q_ptr;
    a.n_queries = n_queries;
    a.k = k;
    a.shift = shift_per_query;
    resolve_mask(ix, weight_mask, a);
    a.out_ids = out_ids;
    a.out_scores = out_scores;
    int rc;
    if ((rc = run_retrieve_on_device(ix, a, stats, n_query_tokens, false))) return rc;
    return finish_stats(ix, a, stats, n_query_tokens, false);
}

int bm25cu_retrieve_device_enqueue(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries,
                                   int64_t n_query_tokens, int32_t k, const float *shift_per_query,
                                   const double *weight_mask, int32_t *out_ids, float *out_scores) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_retrieve_device_enqueue: null handle");
    BM25CU_REQUIRE(n_queries >= 0 && k >= 0, "bm25cu_retrieve_device_enqueue: negative n_queries or k");
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(q_ptr && out_ids && out_scores, "bm25cu_retrieve_device_enqueue: null pointer");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    RetrieveArgs a;
    a.q_tokens = q_tokens;
    a.q_ptr = q_ptr;
    a.n_queries = n_queries;
    a.k = k;
    a.shift = shift_per_query;
    resolve_mask(ix, weight_mask, a);
    a.out_ids = out_ids;
    a.out_scores = out_scores;
    return run_retrieve_on_device(ix, a, nullptr, n_query_tokens, false);
}

int bm25cu_retrieve_collect(bm25cu_index *ix, int32_t n_queries, int64_t n_query_tokens, int32_t k, bm25cu_stats *stats) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_retrieve_collect: null handle");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_queries <= 0 || k <= 0) return BM25CU_OK;
    BM25CU_REQUIRE(ix->d_ctrl.p != nullptr, "bm25cu_retrieve_collect: no retrieve call has been enqueued");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    RetrieveArgs a{};
    a.n_queries = n_queries;
    a.k = k;
    const int launches = ix->last_launches;
    int rc = finish_stats(ix, a, stats, n_query_tokens, false);
    if (rc == BM25CU_OK && stats) stats->launches = launches;
    return rc;
}

int bm25cu_scores_dense(bm25cu_index *ix, const int32_t *q_tokens, int32_t n_tokens, int32_t has_shift, float shift,
                        const double *weight_mask, float *out) {
    BM25CU_REQUIRE(ix != nullptr && out != nullptr, "bm25cu_scores_dense: null pointer");
    BM25CU_REQUIRE(n_tokens >= 0, "bm25cu_scores_dense: negative n_tokens");
    for (int32_t i = 0; i < n_tokens; ++i) {
        if (q_tokens[i] < 0 || q_tokens[i] >= ix->n_vocab) {
            char buf[256];
            snprintf(buf, sizeof(buf),
                     "The maximum token ID in the query (%d) is higher than the number of tokens in the index.",
                     q_tokens[i]);
            set_error(buf);
            return BM25CU_EINVAL;
        }
    }
    BM25CU_CHECK(cudaSetDevice(ix->device));
    int rc;
    const size_t n = (size_t)(ix->n_docs > 0 ? ix->n_docs : 1);
    if ((rc = ix->d_qtok.reserve((size_t)(n_tokens > 0 ? n_tokens : 1) * 4))) return rc;
    if ((rc = ix->d_scratch.reserve(n * 4))) return rc;
    if (weight_mask && (rc = ix->d_mask.reserve(n * 8))) return rc;
    if (n_tokens) BM25CU_CHECK(cudaMemcpyAsync(ix->d_qtok.p, q_tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, ix->stream));
    if (weight_mask) BM25CU_CHECK(cudaMemcpyAsync(ix->d_mask.p, weight_mask, (size_t)ix->n_docs * 8, cudaMemcpyHostToDevice, ix->stream));
    if ((rc = launch_scores_dense(ix, ix->d_qtok.as<int32_t>(), n_tokens, has_shift, shift,
                                  weight_mask ? ix->d_mask.as<double>() : nullptr, ix->d_scratch.as<float>())))
        return rc;
    BM25CU_CHECK(cudaMemcpyAsync(out, ix->d_scratch.p, (size_t)ix->n_docs * 4, cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    return BM25CU_OK;
}

int bm25cu_topk_dense(const void *scores, int32_t is_f64, int64_t n, int32_t k, int32_t sorted, int64_t *out_ids,
                      void *out_scores, int32_t device) {
    (void)sorted;
    BM25CU_REQUIRE(n >= 0 && k >= 0, "bm25cu_topk_dense: negative size");
    BM25CU_REQUIRE(k <= n, "bm25cu_topk_dense: k larger than the number of scores");
    BM25CU_REQUIRE(k <= (1 << 20), "bm25cu_topk_dense: k > 2^20 is not supported");
    if (k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(scores && out_ids && out_scores, "bm25cu_topk_dense: null pointer");
    BM25CU_CHECK(cudaSetDevice(device));
    const size_t es = is_f64 ? 8 : 4;
    void *d_scores = nullptr, *d_os = nullptr;
    int64_t *d_ids = nullptr;
    BM25CU_CHECK(cudaMalloc(&d_scores, (size_t)n * es));
    cudaError_t e1 = cudaMalloc(&d_ids, (size_t)k * 8), e2 = cudaMalloc(&d_os, (size_t)k * es);
    int rc = BM25CU_OK;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("bm25cu_topk_dense: out of device memory");
        rc = BM25CU_ENOMEM;
    } else {
        cudaError_t e = cudaMemcpy(d_scores, scores, (size_t)n * es, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            rc = topk_dense_device(d_scores, is_f64, n, k, d_ids, d_os, 0);
            if (rc == BM25CU_OK) e = cudaStreamSynchronize(0);
            if (rc == BM25CU_OK && e == cudaSuccess) e = cudaMemcpy(out_ids, d_ids, (size_t)k * 8, cudaMemcpyDeviceToHost);
            if (rc == BM25CU_OK && e == cudaSuccess) e = cudaMemcpy(out_scores, d_os, (size_t)k * es, cudaMemcpyDeviceToHost);
        }
        if (e != cudaSuccess) {
            set_error(std::string("bm25cu_topk_dense: ") + cudaGetErrorString(e));
            rc = BM25CU_ECUDA;
        }
    }
    cudaFree(d_scores);
    cudaFree(d_ids);
    cudaFree(d_os);
    return rc;
}

int bm25cu_topk_merge_device(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                             const float *shift_per_query, int32_t *out_ids, float *out_scores, int32_t device, void *stream) {
    BM25CU_REQUIRE(n_parts >= 1 && n_queries >= 0 && k >= 0, "bm25cu_topk_merge: bad sizes");
    BM25CU_CHECK(cudaSetDevice(device));
    int rc = topk_merge_device(ids, scores, n_parts, n_queries, k, shift_per_query, out_ids, out_scores, (cudaStream_t)stream);
    if (rc) return rc;
    BM25CU_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return BM25CU_OK;
}

int bm25cu_topk_merge(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                      int32_t *out_ids, float *out_scores, int32_t device) {
    BM25CU_REQUIRE(n_parts >= 1 && n_queries >= 0 && k >= 0, "bm25cu_topk_merge: bad sizes");
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(ids && scores && out_ids && out_scores, "bm25cu_topk_merge: null pointer");
    BM25CU_CHECK(cudaSetDevice(device));
    const size_t in_elems = (size_t)n_parts * n_queries * k, out_elems = (size_t)n_queries * k;
    int32_t *d_ids = nullptr, *d_oi = nullptr;
    float *d_sc = nullptr, *d_os = nullptr;
    int rc = BM25CU_OK;
    cudaError_t e = cudaMalloc(&d_ids, in_elems * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_sc, in_elems * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_oi, out_elems * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_os, out_elems * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, in_elems * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_sc, scores, in_elems * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = topk_merge_device(d_ids, d_sc, n_parts, n_queries, k, nullptr, d_oi, d_os, 0);
        if (rc == BM25CU_OK) e = cudaStreamSynchronize(0);
        if (rc == BM25CU_OK && e == cudaSuccess) e = cudaMemcpy(out_ids, d_oi, out_elems * 4, cudaMemcpyDeviceToHost);
        if (rc == BM25CU_OK && e == cudaSuccess) e = cudaMemcpy(out_scores, d_os, out_elems * 4, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) {
        set_error(std::string("bm25cu_topk_merge: ") + cudaGetErrorString(e));
        rc = (e == cudaErrorMemoryAllocation) ? BM25CU_ENOMEM : BM25CU_ECUDA;
    }
    cudaFree(d_ids);
    cudaFree(d_sc);
    cudaFree(d_oi);
    cudaFree(d_os);
    return rc;
}

}  // extern "C"
