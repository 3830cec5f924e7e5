// Shared internals of libbm25cu.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bm25cu.h"

namespace bm25cu {

void set_error(const std::string &msg);

#define BM25CU_CHECK(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            char _buf[512];                                                                       \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                     cudaGetErrorString(_e));                                                     \
            bm25cu::set_error(_buf);                                                              \
            return (_e == cudaErrorMemoryAllocation) ? BM25CU_ENOMEM : BM25CU_ECUDA;              \
        }                                                                                         \
    } while (0)

#define BM25CU_REQUIRE(cond, msg)      \
    do {                               \
        if (!(cond)) {                 \
            bm25cu::set_error(msg);    \
            return BM25CU_EINVAL;      \
        }                              \
    } while (0)

// A growable device / pinned-host buffer owned by a handle.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return BM25CU_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        BM25CU_CHECK(cudaMalloc(&p, want));
        cap = want;
        return BM25CU_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return BM25CU_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        BM25CU_CHECK(cudaMallocHost(&p, want));
        cap = want;
        return BM25CU_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// Tuning / debugging knobs of a handle. Filled ONCE, when the handle is created, from the BM25CU_* environment
// variables; changed afterwards only through bm25cu_index_set_option. Names are stored without the BM25CU_ prefix.
struct Knobs {
    std::vector<std::pair<std::string, int>> kv;
    static const char *strip(const char *name) { return std::strncmp(name, "BM25CU_", 7) == 0 ? name + 7 : name; }
    void load_env();
    int find(const char *name) const {
        name = strip(name);
        for (size_t i = 0; i < kv.size(); ++i)
            if (kv[i].first == name) return (int)i;
        return -1;
    }
    bool has(const char *name) const { return find(name) >= 0; }
    int get(const char *name, int dflt) const { const int i = find(name); return i >= 0 ? kv[i].second : dflt; }
    void set(const char *name, int v) { const int i = find(name); if (i >= 0) kv[i].second = v; else kv.emplace_back(strip(name), v); }
    void unset(const char *name) { const int i = find(name); if (i >= 0) kv.erase(kv.begin() + i); }
};

constexpr int kPadElems = 64;  // every posting array is over-allocated by this many elements (bulk-copy tails)

}  // namespace bm25cu

// One doc-range shard of the CSC score matrix, resident in HBM (layout: DESIGN.md "Data layout").
struct bm25cu_index {
    int device = 0;
    int32_t n_vocab = 0;
    int32_t n_docs = 0;    // documents in this shard (local ids 0..n_docs-1)
    int32_t doc_base = 0;  // global id of local doc 0
    int64_t nnz = 0;
    float *d_data = nullptr;       // f32[nnz + pad]  scores, column (token) major, rows ascending
    int32_t *d_indices = nullptr;  // i32[nnz + pad]  local doc ids
    int64_t *d_indptr = nullptr;   // i64[n_vocab + 1]
    float *d_nonocc = nullptr;     // f32[n_vocab] or null (kept for download only)
    float *d_colmax = nullptr;     // f32[n_vocab] largest score of every column (upper bounds for pruning)
    float *d_colkth = nullptr;     // f32[3][n_vocab] 10th / 100th / 1000th largest score of every column (0: fewer postings);
                                   // a query's k-th best document scores at least the k-th largest posting of any of its tokens
    unsigned int *d_btab = nullptr;     // bucket tables of the long columns: posting offset of every 2^sh-document bucket
    long long *d_btab_off = nullptr;    // i64[n_vocab] start of a column's table in d_btab, -1: none (short column)
    unsigned char *d_btab_sh = nullptr; // u8[n_vocab] log2 of the bucket width in documents
    unsigned int *d_pbm = nullptr;      // presence bitmaps of the long columns (one bit per 2^s documents)
    long long *d_pbm_off = nullptr;     // i64[n_vocab] start (in 32-bit words) of a column's bitmap, -1: none
    unsigned char *d_pbm_sh = nullptr;  // u8[n_vocab] s
    void *d_stab = nullptr;             // slot tables of the long columns: 32-byte buckets of four (doc id, score) slots
    long long *d_stab_off = nullptr;    // i64[n_vocab] first bucket of a column's table, -1: none
    unsigned char *d_stab_sh = nullptr; // u8[n_vocab] log2 of the bucket width in documents
    bool stab_compact = false;          // the tables use the compact sector format: 16-bit doc offsets, five slots per sector (STAB_COMPACT)
    long long btab_len = 0, pbm_len = 0, stab_len = 0;  // entries of d_btab / words of d_pbm / buckets of d_stab (bounds checks of `make checked`)
    size_t aux_bytes = 0;               // bytes of the lookup structures derived at upload (bucket tables + presence bitmaps)
    bool nonneg = true;            // all data >= 0 and finite: the fused streaming kernel is exact
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    bool is_view = false;          // bm25cu_index_view: the posting arrays and lookup structures belong to another handle
    bm25cu_index *root = nullptr;  // a view: the handle that owns the arrays
    std::atomic<int> n_views{0};   // live views of this handle: 1 + n_views handles serve the shard, the caller means to keep
                                   // that many batches in flight (retrieve_ms.cu sizes its parts by it; knob IN_FLIGHT overrides)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // per-call workspaces
    bm25cu::DevBuf d_qtok, d_qptr, d_shift, d_mask, d_out_ids, d_out_scores, d_ctrl, d_general_list, d_scratch,
        d_spill, d_order, d_theta, d_fb_list, d_ms_items, d_ms_rows, d_ms_q, d_ms_partial;  // (d_rmask below is not per call)
    int last_launches = 0;         // kernels launched by the last retrieve call
    bool ms_ran = false;           // the last call ran retrieve_ms.cu (ev[5] is recorded behind it)
    bool order_valid = false;      // d_order holds the pull order of the current batch
    bm25cu::PinBuf h_in, h_out;
    int general_rows = 0;
    int rmask_kind = 0;            // resident weight mask (bm25cu_index_set_mask): 0 none, 1 bitmap, 2 float64; in d_rmask
    bm25cu::DevBuf d_rmask;
    bm25cu::Knobs knobs;           // BM25CU_* environment at creation + bm25cu_index_set_option
};

namespace bm25cu {

// ---- kernels / launchers implemented across the .cu files ------------------------------------------------
struct RetrieveArgs {
    const int32_t *q_tokens;  // device
    const int64_t *q_ptr;     // device
    int32_t n_queries;
    int32_t k;
    const float *shift;  // device or null
    const double *mask;  // device or null: float64 weight per document
    const unsigned int *mask_bits;  // device or null: a 0/1 weight mask as one bit per document (1 = keep); excludes `mask`
    int32_t *out_ids;    // device [n_queries * k]
    float *out_scores;   // device
};

// control block in device memory (zeroed before every call)
struct Ctrl {
    unsigned int fast_next;      // work-queue cursor of the fused kernel
    unsigned int general_count;  // number of queries routed to the general kernel
    unsigned int general_next;
    unsigned int error;          // bit 0: token id out of range; bits 1-2: internal invariants of the fused kernel
    unsigned long long posting_count;
    unsigned int mask_bad;       // weight mask has values outside [0, 1] (or NaN): every query takes the general kernel
    unsigned int ms_next;        // work-queue cursor of the list-at-a-time kernel (retrieve_ms.cu)
    unsigned int fb_count;       // queries that kernel handed on to the streaming kernel
    unsigned int ms_items;       // work items (query parts) of that kernel
    unsigned int ms_rows;        // rows handed out in its partial-result buffer
    unsigned int ms_hist[64];    // parts per log2(cost) bucket, and the scatter cursors of the buckets (ms_plan_*_kernel)
    unsigned int ms_cursor[64];
    unsigned long long prof[48];  // phase cycle counters / event counts of the fused kernel (BM25CU_PROFILE builds)
};

// qlist/qcount non-null: serve only the *qcount queries listed in qlist (what retrieve_ms.cu handed on)
int launch_retrieve_fast(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_general_list,
                         int *launches, const unsigned int *qlist = nullptr, const unsigned int *qcount = nullptr);
bool retrieve_ms_takes(const bm25cu_index *ix, const RetrieveArgs &a);  // false: the batch skips retrieve_ms.cu altogether
int launch_retrieve_ms(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_fb_list, int *launches);
int launch_retrieve_general(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl,
                            const unsigned int *d_general_list, int *launches);
int launch_scores_dense(bm25cu_index *ix, const int32_t *d_q, int32_t n_tokens, int has_shift, float shift,
                        const double *d_mask, float *d_out);
int topk_dense_device(const void *d_scores, int is_f64, int64_t n, int32_t k, int64_t *d_out_ids, void *d_out_scores,
                      cudaStream_t stream);
int topk_merge_device(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                      const float *shift_or_null, int32_t *out_ids, float *out_scores, cudaStream_t stream);
int finalize_index(bm25cu_index *ix);          // device attributes, events, non-negativity scan
int alloc_postings(bm25cu_index *ix, int64_t nnz);  // d_data / d_indices with bulk-copy padding
// structural checks of a CSC on the device (indptr monotone, ids in range, rows strictly ascending): BM25CU_EINVAL if not
int validate_csc(const int32_t *d_indices, const int64_t *d_indptr, int32_t n_vocab, int64_t nnz, int32_t n_docs, cudaStream_t stream);

}  // namespace bm25cu
