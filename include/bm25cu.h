/*
 * bm25cu.h - C ABI of libbm25cu.so: the B200 (sm_100a) implementation of the bm25s query-time hot path.
 *
 * The reference (xhluca/bm25s) is pure Python and has no FFI; its seam for this path is the string
 * `BM25(backend=...)` (bm25s/__init__.py:154,223-234) branched on in `BM25.retrieve`
 * (bm25s/__init__.py:847). Each entry point below names the reference code it replaces; the ctypes
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status: 0 ok, 1 invalid argument, 2 CUDA error, 4 out of memory;
 *     bm25cu_last_error() returns a thread-local message for the last non-zero status;
 *   - no exceptions cross the ABI, all pointers are plain host pointers unless the name says `_device`;
 *   - host buffers must be C-contiguous; the caller keeps ownership of everything it passes in;
 *   - a handle (`bm25cu_index`) owns one doc-range shard on ONE device with one stream; it is not
 *     thread-safe (serialise calls on the same handle), different handles are independent;
 *   - document ids in results are GLOBAL ids (local id + doc_lo of the shard).
 */
#ifndef BM25CU_H
#define BM25CU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BM25CU_OK 0
#define BM25CU_EINVAL 1
#define BM25CU_ECUDA 2
#define BM25CU_ENOMEM 4

/* tfc / idf variants, bm25s/scoring.py:162-175 and :222-235 */
#define BM25CU_ROBERTSON 0
#define BM25CU_LUCENE 1
#define BM25CU_ATIRE 2
#define BM25CU_BM25L 3
#define BM25CU_BM25PLUS 4

typedef struct bm25cu_index bm25cu_index;

/* Timing / accounting of the last retrieve call on a handle (all device work, CUDA events). */
typedef struct bm25cu_stats {
    float kernel_ms;           /* device time of the scoring+top-k kernels of the call (CUDA events)        */
    float fast_kernel_ms;      /* ... of the fused streaming kernel alone                                   */
    float total_ms;            /* device time of the whole call incl. H2D / D2H copies                      */
    int64_t posting_bytes;     /* algorithmic bytes: 8 * sum over query tokens of df(t)       */
    int64_t algorithmic_bytes; /* posting_bytes + 20 * sum|q| + 8 * k * n_queries                            */
    int32_t n_fast;            /* queries served by the fused streaming kernel                              */
    int32_t n_general;         /* queries served by the dense general kernel (exotic inputs) */
    int32_t launches;          /* kernels launched by the call                                              */
    float ms_kernel_ms;        /* ... of the list-at-a-time kernel alone (part of fast_kernel_ms)            */
} bm25cu_stats;

const char *bm25cu_last_error(void);
int bm25cu_version(void);
int bm25cu_device_count(int32_t *n);

/*
 * Upload a CSC score matrix (the reference's `self.scores` dict, bm25s/__init__.py:432-438 / :1121-1127)
 * to `device`, keeping only the rows (documents) in [doc_lo, doc_hi) - the doc-range shard of this handle.
 * Pass doc_lo = 0, doc_hi = n_docs for the whole matrix. Arrays are copied; memmaps are fine.
 * `nonocc` (bm25s/__init__.py:370-384) may be NULL; it is only stored so that download() can return it.
 */
int bm25cu_index_upload(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz,
                        int32_t n_vocab, int32_t n_docs, const float *nonocc, int32_t device, int32_t doc_lo,
                        int32_t doc_hi, bm25cu_index **out);

/* Same, but data / indices / indptr / nonocc are DEVICE pointers on `device` (tensor interop; copied device to device). */
int bm25cu_index_upload_device(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz,
                               int32_t n_vocab, int32_t n_docs, const float *nonocc, int32_t device, int32_t doc_lo,
                               int32_t doc_hi, bm25cu_index **out);

/* A second handle on the SAME resident shard (no copy of the index): it shares the posting arrays and lookup structures
 * of `src` and owns a stream, events and per-call workspaces of its own, so calls on the view and on `src` may be in
 * flight at the same time - several batches on several streams overlap one batch's tail with the next one's start
 * (bench.py --lanes). The number of handles on a shard (1 + live views) is taken as the number of batches the caller keeps
 * in flight, and the scoring kernel cuts a batch into coarser parts accordingly (option "IN_FLIGHT" overrides; the rows do
 * not depend on it). A view has no resident mask of its own. Free it before `src`. */
int bm25cu_index_view(const bm25cu_index *src, bm25cu_index **out);

/* Run the handle's device work on a caller-owned CUDA stream (cudaStream_t, e.g. torch's current stream) so that
 * the caller can bracket calls with its own events. Pass the stream's raw handle. */
int bm25cu_index_set_stream(bm25cu_index *ix, void *stream);

/*
 * Build the index on the GPU from flat token ids - replaces BM25.build_index_from_ids
 * (bm25s/__init__.py:326-438; scoring.py:28-57 df, :60-73 idf, :76-112 nonoccurrence, :115-159 tfc,
 * :246-309 score triples, :371-432 CSC assembly). `tokens`/`doc_ptr` describe the documents of THIS shard
 * (local doc i = tokens[doc_ptr[i]:doc_ptr[i+1]]), whose global ids start at doc_lo.
 * Two-step form for doc-sharded builds: begin() counts, the caller sums df / n_docs / total length over
 * shards (an allreduce), finish() scores and assembles. bm25cu_index_build() is begin+finish for one shard.
 */
typedef struct bm25cu_builder bm25cu_builder;
int bm25cu_build_begin(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs_local, int32_t n_vocab,
                       int32_t device, int32_t tokens_on_device, bm25cu_builder **out);
/* device pointer to the shard-local document frequencies int32[n_vocab] (sum it across shards in place) */
int bm25cu_build_df_device(bm25cu_builder *b, int32_t **df_device);
int bm25cu_build_df_get(bm25cu_builder *b, int32_t *df_host);
int bm25cu_build_df_set(bm25cu_builder *b, const int32_t *df_host);
int bm25cu_build_finish(bm25cu_builder *b, int64_t n_docs_global, int64_t total_len_global, int32_t tfc_method,
                        int32_t idf_method, double k1, double bparam, double delta, int32_t doc_lo,
                        bm25cu_index **out);
int bm25cu_build_free(bm25cu_builder *b);
int bm25cu_index_build(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs, int32_t n_vocab,
                       int32_t tfc_method, int32_t idf_method, double k1, double bparam, double delta,
                       int32_t device, int32_t tokens_on_device, bm25cu_index **out);

/* Tuning / debugging knobs of a handle (DESIGN.md lists them). The BM25CU_<NAME> environment variables are read ONCE,
 * when a handle is created; afterwards a knob changes only through these calls. `name` with or without the BM25CU_
 * prefix, e.g. bm25cu_index_set_option(ix, "MS", 0) serves every query with the streaming kernel alone. */
int bm25cu_index_set_option(bm25cu_index *ix, const char *name, int32_t value);
int bm25cu_index_unset_option(bm25cu_index *ix, const char *name);

/* Resident weight mask of a handle (bm25s/__init__.py:610-612, 833-845): uploaded once, applied by every later retrieve
 * call whose weight_mask argument is NULL, until replaced or cleared - a filter used call after call costs nothing per
 * call. BM25CU_MASK_BITS: a 0/1 mask as uint32 words, bit (d % 32) of word (d / 32) = keep shard-local document d
 * (multiplying by 1 or by 0 is exact, so one bit per document carries the mask; its documents are dropped before any
 * lookup). BM25CU_MASK_F64: float64[n_docs_local], as the weight_mask argument. BM25CU_MASK_NONE clears. */
#define BM25CU_MASK_NONE 0
#define BM25CU_MASK_BITS 1
#define BM25CU_MASK_F64 2
int bm25cu_index_set_mask(bm25cu_index *ix, int32_t kind, const void *mask_host);

/* Diagnostics. _colkth: the upload-time table of the 10th / 100th / 1000th largest score of every column,
 * float32[3][n_vocab] (start values of the pruning threshold). _prof: 48 uint64 phase counters of the last retrieve call
 * (all zero unless the library was built with -DBM25CU_PROFILE: `make prof`). */
int bm25cu_debug_colkth(bm25cu_index *ix, float *out);
int bm25cu_debug_prof(bm25cu_index *ix, unsigned long long *out);

/* Sizes of a shard, then a copy of it back into caller-allocated numpy arrays (so that `self.scores`
 * and BM25.save, bm25s/__init__.py:1010-1017, stay reference-identical). indices are shard-local ids. */
int bm25cu_index_sizes(const bm25cu_index *ix, int64_t *nnz, int32_t *n_vocab, int32_t *n_docs_local,
                       int32_t *doc_lo, int32_t *has_nonocc);
int bm25cu_index_download(const bm25cu_index *ix, float *data, int32_t *indices, int64_t *indptr, float *nonocc);

/*
 * Batched scoring + top-k: replaces _retrieve_internal_jitted_parallel / _retrieve_numba_functional
 * (bm25s/numba/retrieve_utils.py:13-61, :64-153) and, per query, get_scores_from_ids + selection.topk
 * (bm25s/__init__.py:581-620, bm25s/selection.py:14-64). Queries use the flat layout of
 * retrieve_utils.py:114-115: token ids int32[sum|q|] + pointers int64[n_queries + 1].
 *   shift_per_query : float32[n_queries] or NULL - nonoccurrence_array[ids].sum() per query (BM25L/BM25+,
 *                     bm25s/__init__.py:616-618), added to every score AFTER the mask (numpy-backend order);
 *                     it is not added for empty queries (bm25s/__init__.py:653-657).
 *   weight_mask     : float64[n_docs_local] or NULL - the mask converted exactly to float64; scores become
 *                     (float)((double)score * mask[d]) as in `scores *= weight_mask` (bm25s/__init__.py:610-612).
 *   out_ids/out_scores : int32 / float32 [n_queries * k], each row sorted by score descending (ties by id
 *                     ascending); when the shard has fewer than k documents the tail is id -1, score -inf.
 * Token ids >= n_vocab return BM25CU_EINVAL ("maximum token ID", bm25s/__init__.py:593-599).
 */
int bm25cu_retrieve(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries, int32_t k,
                    int32_t sorted, const float *shift_per_query, const double *weight_mask, int32_t *out_ids,
                    float *out_scores, bm25cu_stats *stats);

/* Same, all six array arguments are DEVICE pointers on the handle's device (tensor interop, no copies);
 * work is enqueued on the handle's stream and the call returns after synchronising it. */
int bm25cu_retrieve_device(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries,
                           int64_t n_query_tokens, int32_t k, const float *shift_per_query,
                           const double *weight_mask, int32_t *out_ids, float *out_scores, bm25cu_stats *stats);

/* bm25cu_retrieve_device in two halves, for callers that chain more device work behind the scoring kernels (the
 * exchange below): _enqueue launches the kernels on the handle's stream and returns at once, nothing is synchronised;
 * _collect synchronises the stream, reports the errors of the enqueued call (token id out of range, internal) and
 * fills `stats` (may be NULL) - call it before the outputs are read on the host. */
int bm25cu_retrieve_device_enqueue(bm25cu_index *ix, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries,
                                   int64_t n_query_tokens, int32_t k, const float *shift_per_query,
                                   const double *weight_mask, int32_t *out_ids, float *out_scores);
int bm25cu_retrieve_collect(bm25cu_index *ix, int32_t n_queries, int64_t n_query_tokens, int32_t k, bm25cu_stats *stats);

/* Dense scores of one query: replaces BM25.get_scores_from_ids (bm25s/__init__.py:581-620). out: float32[n_docs_local] */
int bm25cu_scores_dense(bm25cu_index *ix, const int32_t *q_tokens, int32_t n_tokens, int32_t has_shift, float shift,
                        const double *weight_mask, float *out);

/* k largest of a 1-D host vector: replaces selection.topk (bm25s/selection.py:48-64). scores float32 or float64
 * (is_f64), out_scores same dtype as the input, out_ids int64. */
int bm25cu_topk_dense(const void *scores, int32_t is_f64, int64_t n, int32_t k, int32_t sorted, int64_t *out_ids,
                      void *out_scores, int32_t device);

/* Merge n_parts partial top-k lists per query (doc-range shards) into the global top-k. Inputs are
 * [n_parts][n_queries][k] (part-major), outputs [n_queries][k]. No reference equivalent (the reference is
 * single-process); order is score descending, id ascending, independent of n_parts. */
int bm25cu_topk_merge(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                      int32_t *out_ids, float *out_scores, int32_t device);
/* _device: all pointers are device pointers; `shift_per_query` (float32[n_queries] or NULL) is the BM25L / BM25+ shift,
 * added AFTER the merge: shards score WITHOUT it, so the merged order - and the rows - do not depend on n_parts. */
int bm25cu_topk_merge_device(const int32_t *ids, const float *scores, int32_t n_parts, int32_t n_queries, int32_t k,
                             const float *shift_per_query, int32_t *out_ids, float *out_scores, int32_t device,
                             void *stream);

/*
 * Exchange of the shards' top-k rows over NVLink peer memory, fused with the merge (doc-range sharding over the GPUs of
 * one box, one process per GPU). No reference equivalent (the reference is single-process); it stands where an
 * all_gather + bm25cu_topk_merge_device would. Protocol: every rank creates an exchange for at most max_elems =
 * n_queries * k result elements, publishes its bm25cu_exchange_handle (bm25cu_exchange_handle_size() opaque bytes,
 * a CUDA IPC handle) to the others by any means (torch.distributed, MPI, a file), connects with the world handles
 * (rank-major, its own slot is ignored) and, after a barrier, calls bm25cu_exchange_merge once per batch on every rank:
 * the local rows (device pointers, as written by bm25cu_retrieve_device WITHOUT a shift) are stored into every peer's
 * buffer, and the merged rows [n_queries][k] (score descending, id ascending; the BM25L / BM25+ shift, if given, added
 * after the merge) appear in out_* on every rank. The push kernel runs on `stream`, the merge kernel - which waits for the
 * slowest shard - on a side stream of the exchange, so `stream` is free for the scoring kernels of the next batch;
 * nothing is synchronised. bm25cu_exchange_join makes a stream wait for the last merge (device-side), bm25cu_exchange_check
 * does that, synchronises the stream and reports a peer that never delivered: call one of them before out_* is read.
 */
typedef struct bm25cu_exchange bm25cu_exchange;
int bm25cu_exchange_create(int32_t device, int32_t rank, int32_t world, int64_t max_elems, bm25cu_exchange **out);
int bm25cu_exchange_handle_size(void);
int bm25cu_exchange_handle(bm25cu_exchange *x, void *handle_out);
int bm25cu_exchange_connect(bm25cu_exchange *x, const void *handles);
int bm25cu_exchange_merge(bm25cu_exchange *x, const int32_t *ids_device, const float *scores_device, int32_t n_queries,
                          int32_t k, const float *shift_per_query_device, int32_t *out_ids_device,
                          float *out_scores_device, void *stream);
int bm25cu_exchange_check(bm25cu_exchange *x, void *stream);
int bm25cu_exchange_join(bm25cu_exchange *x, void *stream);
int bm25cu_exchange_free(bm25cu_exchange *x);

int bm25cu_index_free(bm25cu_index *ix);

/*
 * One index over SEVERAL GPUs of the box, driven from one process - the multi-device form of the entry points above
 * (the reference's seam is one constructor string, bm25s/__init__.py:154 / :847; the Python mirror passes
 * `BM25(backend="cuda", devices=[0, 1, ...])`). Documents are sharded by contiguous doc-id ranges, shard r =
 * [n_docs * r / n, n_docs * (r + 1) / n) on devices[r]; only a shard's rows travel to its device (_upload) or are built
 * there (_build: shard-local counting, document frequencies summed on the host, shard-local scoring). _retrieve has the
 * contract of bm25cu_retrieve: every shard scores the whole batch (the calls are enqueued by one host thread per
 * device), the shards' rows are copied to devices[0] over NVLink, merged there (score descending, id ascending; the
 * shift added after the merge, so the rows do not depend on the number of shards) and returned. `weight_mask` is the
 * whole float64[n_docs] mask. _download returns the whole matrix in the reference's layout (global ids), _scores_dense
 * fills float32[n_docs]. A group is not thread-safe. No torch, no process group.
 */
typedef struct bm25cu_group bm25cu_group;
int bm25cu_group_upload(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz, int32_t n_vocab,
                        int32_t n_docs, const float *nonocc, const int32_t *devices, int32_t n_devices,
                        bm25cu_group **out);
int bm25cu_group_build(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs, int32_t n_vocab,
                       int32_t tfc_method, int32_t idf_method, double k1, double bparam, double delta,
                       const int32_t *devices, int32_t n_devices, bm25cu_group **out);
int bm25cu_group_sizes(const bm25cu_group *g, int64_t *nnz, int32_t *n_vocab, int32_t *n_docs, int32_t *n_shards,
                       int32_t *has_nonocc);
int bm25cu_group_shard(bm25cu_group *g, int32_t i, bm25cu_index **ix); /* borrowed: freed with the group */
int bm25cu_group_set_option(bm25cu_group *g, const char *name, int32_t value);
int bm25cu_group_unset_option(bm25cu_group *g, const char *name);
int bm25cu_group_set_mask(bm25cu_group *g, int32_t kind, const void *mask_host); /* the WHOLE mask; shards take their ranges */
int bm25cu_group_download(const bm25cu_group *g, float *data, int32_t *indices, int64_t *indptr, float *nonocc);
int bm25cu_group_retrieve(bm25cu_group *g, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries, int32_t k,
                          const float *shift_per_query, const double *weight_mask, int32_t *out_ids,
                          float *out_scores, bm25cu_stats *stats);
int bm25cu_group_scores_dense(bm25cu_group *g, const int32_t *q_tokens, int32_t n_tokens, int32_t has_shift,
                              float shift, const double *weight_mask, float *out);
int bm25cu_group_free(bm25cu_group *g);

#ifdef __cplusplus
}
#endif
#endif /* BM25CU_H */
