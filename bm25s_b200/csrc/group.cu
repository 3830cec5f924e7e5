// bm25cu_group: one index sharded by contiguous doc-id ranges over several GPUs of the box, driven from ONE process
// (the reference's seam is one constructor string, bm25s/__init__.py:154, 847 - so is this: the Python
// mirror passes `devices=[...]`). One bm25cu_index shard per device; a call fans out over a small pool of host threads
// (one per device: the query upload and the kernel launches of the shards are enqueued side by side), every shard scores
// the whole batch on its documents, the shards' top-k rows travel to the first device with peer copies over NVLink, the
// merge kernel there unites them (score descending, id ascending; BM25L / BM25+ shift added after the merge) and the rows
// come back to the host. No torch, no process group: `torch.distributed` + the peer-memory exchange of exchange.cu stay
// the form for one process per GPU.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace bm25cu {

// N persistent host threads; run(fn) executes fn(r) on thread r for r = 0 .. n - 1 and returns when all are done.
class WorkerPool {
  public:
    explicit WorkerPool(int n) : n_(n), status_((size_t)n, BM25CU_OK), errors_((size_t)n) {
        for (int r = 0; r < n; ++r) threads_.emplace_back([this, r]() { loop(r); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // returns the first non-zero status; the matching message becomes the caller thread's last error
    int run(const std::function<int(int)> &fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            pending_ = n_;
            ++generation_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this]() { return pending_ == 0; });
        fn_ = nullptr;
        for (int r = 0; r < n_; ++r)
            if (status_[(size_t)r] != BM25CU_OK) {
                set_error(errors_[(size_t)r]);
                return status_[(size_t)r];
            }
        return BM25CU_OK;
    }

  private:
    void loop(int r) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<int(int)> *fn = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&]() { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                fn = fn_;
            }
            int rc = BM25CU_OK;
            if (fn) {
                rc = (*fn)(r);
                status_[(size_t)r] = rc;
                errors_[(size_t)r] = rc != BM25CU_OK ? std::string(bm25cu_last_error()) : std::string();  // thread-local message of this worker
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<int(int)> *fn_ = nullptr;
    unsigned long long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    std::vector<int> status_;
    std::vector<std::string> errors_;
};

}  // namespace bm25cu

using namespace bm25cu;

struct bm25cu_group {
    int n = 0;
    int32_t n_docs = 0, n_vocab = 0;
    std::vector<bm25cu_index *> shard;
    std::vector<int> device;
    std::vector<int32_t> lo, hi;
    std::vector<cudaEvent_t> done;         // per shard: its rows have arrived on the first device
    DevBuf g_ids, g_scores, m_ids, m_scores, d_shift;  // on device[0]: gathered rows [n][Q][k], merged rows, shift
    PinBuf h_in, h_out;
    WorkerPool *pool = nullptr;
    bool has_nonocc = false;
};

namespace {

void group_destroy(bm25cu_group *g) {
    if (!g) return;
    delete g->pool;
    for (int r = 0; r < (int)g->shard.size(); ++r) {
        if (g->done.size() > (size_t)r && g->done[(size_t)r]) {
            cudaSetDevice(g->device[(size_t)r]);
            cudaEventDestroy(g->done[(size_t)r]);
        }
        if (g->shard[(size_t)r]) bm25cu_index_free(g->shard[(size_t)r]);
    }
    if (!g->device.empty()) cudaSetDevice(g->device[0]);
    g->g_ids.release();
    g->g_scores.release();
    g->m_ids.release();
    g->m_scores.release();
    g->d_shift.release();
    g->h_in.release();
    g->h_out.release();
    delete g;
}

int group_prepare(bm25cu_group *g, const int32_t *devices, int32_t n_devices, int32_t n_docs, int32_t n_vocab) {
    g->n = n_devices;
    g->n_docs = n_docs;
    g->n_vocab = n_vocab;
    g->shard.assign((size_t)n_devices, nullptr);
    g->done.assign((size_t)n_devices, nullptr);
    for (int r = 0; r < n_devices; ++r) {
        g->device.push_back(devices[r]);
        g->lo.push_back((int32_t)((long long)n_docs * r / n_devices));
        g->hi.push_back((int32_t)((long long)n_docs * (r + 1) / n_devices));
    }
    // direct NVLink copies between the shards' devices and the first one (without it the peer copies are staged)
    for (int r = 1; r < n_devices; ++r) {
        int can = 0;
        if (devices[r] != devices[0] && cudaDeviceCanAccessPeer(&can, devices[0], devices[r]) == cudaSuccess && can) {
            cudaSetDevice(devices[0]);
            if (cudaDeviceEnablePeerAccess(devices[r], 0) != cudaSuccess) cudaGetLastError();
            cudaSetDevice(devices[r]);
            if (cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();
        }
    }
    g->pool = new WorkerPool(n_devices);
    return BM25CU_OK;
}

int group_finish(bm25cu_group *g) {
    for (int r = 0; r < g->n; ++r) {
        BM25CU_CHECK(cudaSetDevice(g->device[(size_t)r]));
        BM25CU_CHECK(cudaEventCreateWithFlags(&g->done[(size_t)r], cudaEventDisableTiming));
    }
    g->has_nonocc = g->shard[0]->d_nonocc != nullptr;
    return BM25CU_OK;
}

int check_devices(const int32_t *devices, int32_t n_devices) {
    BM25CU_REQUIRE(devices != nullptr && n_devices >= 1 && n_devices <= 64, "bm25cu_group: 1 .. 64 devices");
    int count = 0;
    BM25CU_CHECK(cudaGetDeviceCount(&count));
    // a device may be listed more than once: several shards on one GPU (how the one-GPU tests exercise this path)
    for (int r = 0; r < n_devices; ++r) BM25CU_REQUIRE(devices[r] >= 0 && devices[r] < count, "bm25cu_group: no such device");
    return BM25CU_OK;
}

// Host-side cut of a CSC to the documents [lo, hi): only these rows travel to the shard's device. Per column the slice
// is found by two bound searches in the ascending row ids; slices are gathered into two pinned staging buffers and
// copied while the next one fills. `indices` are rebased to shard-local ids on the way.
int upload_shard_rows(const float *data, const int32_t *indices, const int64_t *indptr, int32_t n_vocab, int32_t n_docs,
                      const float *nonocc, int device, int32_t lo, int32_t hi, bm25cu_index **out) {
    *out = nullptr;
    BM25CU_CHECK(cudaSetDevice(device));
    std::vector<int64_t> start((size_t)n_vocab), sptr((size_t)n_vocab + 1);
    sptr[0] = 0;
    for (int32_t t = 0; t < n_vocab; ++t) {
        const int32_t *b = indices + indptr[t], *e = indices + indptr[t + 1];
        const int32_t *a = (lo <= 0) ? b : std::lower_bound(b, e, lo);
        const int32_t *z = (hi >= n_docs) ? e : std::lower_bound(a, e, hi);
        start[(size_t)t] = a - indices;
        sptr[(size_t)t + 1] = sptr[(size_t)t] + (z - a);
    }
    const int64_t nnz = sptr[(size_t)n_vocab];
    bm25cu_index *ix = new bm25cu_index();
    ix->device = device;
    ix->n_vocab = n_vocab;
    ix->n_docs = hi - lo;
    ix->doc_base = lo;
    ix->nnz = nnz;
    char *stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    const size_t kChunk = (size_t)1 << 21;  // postings per staging buffer (8 MB of scores or ids)
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            if (stage[b]) cudaFreeHost(stage[b]);
            if (ev[b]) cudaEventDestroy(ev[b]);
        }
    };
    auto fail = [&](int code) { cleanup(); bm25cu_index_free(ix); return code; };
#define GTRY(expr)                                                                                           \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            set_error(std::string("bm25cu_group_upload: ") + #expr + ": " + cudaGetErrorString(_e));         \
            return fail(_e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA);                     \
        }                                                                                                    \
    } while (0)
    GTRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    int rc = alloc_postings(ix, nnz);
    if (rc) return fail(rc);
    GTRY(cudaMalloc(&ix->d_indptr, (size_t)(n_vocab + 1) * sizeof(int64_t)));
    GTRY(cudaMemcpyAsync(ix->d_indptr, sptr.data(), (size_t)(n_vocab + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    if (nonocc) {
        GTRY(cudaMalloc(&ix->d_nonocc, (size_t)(n_vocab > 0 ? n_vocab : 1) * sizeof(float)));
        GTRY(cudaMemcpyAsync(ix->d_nonocc, nonocc, (size_t)n_vocab * sizeof(float), cudaMemcpyHostToDevice, ix->stream));
    }
    for (int b = 0; b < 2; ++b) {
        GTRY(cudaHostAlloc((void **)&stage[b], kChunk * 4, cudaHostAllocDefault));
        GTRY(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
    }
    // two passes over the columns: scores, then ids (rebased); a staging buffer is refilled only after its copy has ended
    int used[2] = {0, 0};
    int64_t i = 0;
    for (int pass = 0; pass < 2 && nnz > 0; ++pass) {
        int32_t t = 0;
        int64_t in_col = 0;  // postings of column t already taken
        for (int64_t off = 0; off < nnz; off += (int64_t)kChunk, ++i) {
            const int b = (int)(i & 1);
            const int64_t n = std::min<int64_t>((int64_t)kChunk, nnz - off);
            if (used[b]) GTRY(cudaEventSynchronize(ev[b]));
            int64_t filled = 0;
            while (filled < n) {
                const int64_t col_n = sptr[(size_t)t + 1] - sptr[(size_t)t];
                if (in_col >= col_n) { ++t; in_col = 0; continue; }
                const int64_t take = std::min(col_n - in_col, n - filled);
                const int64_t src = start[(size_t)t] + in_col;
                if (pass == 0) {
                    std::memcpy(stage[b] + filled * 4, data + src, (size_t)take * 4);
                } else {
                    int32_t *dst = reinterpret_cast<int32_t *>(stage[b]) + filled;
                    for (int64_t j = 0; j < take; ++j) dst[j] = indices[src + j] - lo;
                }
                filled += take;
                in_col += take;
            }
            void *dst_dev = pass == 0 ? (void *)(ix->d_data + off) : (void *)(ix->d_indices + off);
            GTRY(cudaMemcpyAsync(dst_dev, stage[b], (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
            GTRY(cudaEventRecord(ev[b], ix->stream));
            used[b] = 1;
        }
    }
    GTRY(cudaStreamSynchronize(ix->stream));
#undef GTRY
    cleanup();
    // the same structural checks as a whole-matrix upload (a matrix with unsorted rows was cut at the wrong places)
    if ((rc = validate_csc(ix->d_indices, ix->d_indptr, n_vocab, nnz, ix->n_docs, ix->stream))) {
        bm25cu_index_free(ix);
        return rc;
    }
    if ((rc = finalize_index(ix))) {
        bm25cu_index_free(ix);
        return rc;
    }
    *out = ix;
    return BM25CU_OK;
}

}  // namespace

extern "C" {

int bm25cu_group_upload(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz, int32_t n_vocab,
                        int32_t n_docs, const float *nonocc, const int32_t *devices, int32_t n_devices, bm25cu_group **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_group_upload: null output handle");
    *out = nullptr;
    int rc = check_devices(devices, n_devices);
    if (rc) return rc;
    BM25CU_REQUIRE(nnz >= 0 && n_vocab >= 0 && n_docs >= 0 && indptr != nullptr, "bm25cu_group_upload: bad sizes / null indptr");
    BM25CU_REQUIRE(nnz == 0 || (data != nullptr && indices != nullptr), "bm25cu_group_upload: null data/indices");
    BM25CU_REQUIRE(indptr[0] == 0 && indptr[n_vocab] == nnz, "bm25cu_group_upload: indptr does not span [0, nnz]");
    for (int32_t t = 0; t < n_vocab; ++t)  // the host-side cut walks the columns: their bounds must be sane before it does
        BM25CU_REQUIRE(indptr[t] <= indptr[t + 1], "bm25cu_group_upload: indptr is not non-decreasing");
    bm25cu_group *g = new bm25cu_group();
    group_prepare(g, devices, n_devices, n_docs, n_vocab);
    if (n_devices == 1) {
        rc = bm25cu_index_upload(data, indices, indptr, nnz, n_vocab, n_docs, nonocc, devices[0], 0, n_docs, &g->shard[0]);
    } else {
        rc = g->pool->run([&](int r) {
            return upload_shard_rows(data, indices, indptr, n_vocab, n_docs, nonocc, g->device[(size_t)r], g->lo[(size_t)r],
                                     g->hi[(size_t)r], &g->shard[(size_t)r]);
        });
    }
    if (rc == BM25CU_OK) rc = group_finish(g);
    if (rc) {
        const std::string msg = bm25cu_last_error();
        group_destroy(g);
        set_error(msg);
        return rc;
    }
    *out = g;
    return BM25CU_OK;
}

int bm25cu_group_build(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs, int32_t n_vocab, int32_t tfc_method,
                       int32_t idf_method, double k1, double bparam, double delta, const int32_t *devices, int32_t n_devices,
                       bm25cu_group **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_group_build: null output handle");
    *out = nullptr;
    int rc = check_devices(devices, n_devices);
    if (rc) return rc;
    BM25CU_REQUIRE(n_docs >= 0 && n_vocab >= 0 && doc_ptr != nullptr, "bm25cu_group_build: bad sizes / null doc_ptr");
    bm25cu_group *g = new bm25cu_group();
    group_prepare(g, devices, n_devices, n_docs, n_vocab);
    std::vector<bm25cu_builder *> bld((size_t)n_devices, nullptr);
    std::vector<std::vector<int64_t>> ptr_local((size_t)n_devices);
    std::vector<std::vector<int32_t>> df((size_t)n_devices);
    // shard-local counting (one host thread per device) ...
    rc = g->pool->run([&](int r) {
        const int32_t lo = g->lo[(size_t)r], hi = g->hi[(size_t)r];
        auto &pl = ptr_local[(size_t)r];
        pl.resize((size_t)(hi - lo) + 1);
        for (int32_t d = lo; d <= hi; ++d) pl[(size_t)(d - lo)] = doc_ptr[d] - doc_ptr[lo];
        int rc2 = bm25cu_build_begin(tokens ? tokens + doc_ptr[lo] : nullptr, pl.data(), hi - lo, n_vocab, g->device[(size_t)r], 0, &bld[(size_t)r]);
        if (rc2) return rc2;
        df[(size_t)r].resize((size_t)(n_vocab > 0 ? n_vocab : 1));
        return n_devices > 1 ? bm25cu_build_df_get(bld[(size_t)r], df[(size_t)r].data()) : BM25CU_OK;
    });
    // ... the document frequencies summed over the shards on the host (V integers per shard) ...
    if (rc == BM25CU_OK && n_devices > 1) {
        for (int r = 1; r < n_devices; ++r)
            for (int32_t t = 0; t < n_vocab; ++t) df[0][(size_t)t] += df[(size_t)r][(size_t)t];
    }
    // ... and shard-local scoring with the global statistics
    if (rc == BM25CU_OK) {
        const int64_t total_len = doc_ptr[n_docs] - doc_ptr[0];
        rc = g->pool->run([&](int r) {
            int rc2 = n_devices > 1 ? bm25cu_build_df_set(bld[(size_t)r], df[0].data()) : BM25CU_OK;
            if (rc2) return rc2;
            return bm25cu_build_finish(bld[(size_t)r], n_docs, total_len, tfc_method, idf_method, k1, bparam, delta, g->lo[(size_t)r],
                                       &g->shard[(size_t)r]);
        });
    }
    const std::string msg = rc ? bm25cu_last_error() : "";
    for (auto *b : bld)
        if (b) bm25cu_build_free(b);
    if (rc == BM25CU_OK) rc = group_finish(g);
    if (rc) {
        group_destroy(g);
        if (!msg.empty()) set_error(msg);
        return rc;
    }
    *out = g;
    return BM25CU_OK;
}

int bm25cu_group_sizes(const bm25cu_group *g, int64_t *nnz, int32_t *n_vocab, int32_t *n_docs, int32_t *n_shards, int32_t *has_nonocc) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_sizes: null handle");
    int64_t total = 0;
    for (auto *ix : g->shard) total += ix->nnz;
    if (nnz) *nnz = total;
    if (n_vocab) *n_vocab = g->n_vocab;
    if (n_docs) *n_docs = g->n_docs;
    if (n_shards) *n_shards = g->n;
    if (has_nonocc) *has_nonocc = g->has_nonocc;
    return BM25CU_OK;
}

int bm25cu_group_shard(bm25cu_group *g, int32_t i, bm25cu_index **ix) {
    BM25CU_REQUIRE(g != nullptr && ix != nullptr && i >= 0 && i < g->n, "bm25cu_group_shard: bad arguments");
    *ix = g->shard[(size_t)i];
    return BM25CU_OK;
}

int bm25cu_group_set_option(bm25cu_group *g, const char *name, int32_t value) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_set_option: null handle");
    for (auto *ix : g->shard) {
        int rc = bm25cu_index_set_option(ix, name, value);
        if (rc) return rc;
    }
    return BM25CU_OK;
}

int bm25cu_group_set_mask(bm25cu_group *g, int32_t kind, const void *mask_host) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_set_mask: null handle");
    BM25CU_REQUIRE(kind == BM25CU_MASK_NONE || mask_host != nullptr, "bm25cu_group_set_mask: null mask");
    for (int r = 0; r < g->n; ++r) {
        bm25cu_index *ix = g->shard[(size_t)r];
        int rc;
        if (kind == BM25CU_MASK_BITS && g->n > 1) {
            // the shard's bits [lo, hi) of the whole bitmap, shifted down to bit 0
            const uint32_t *w = static_cast<const uint32_t *>(mask_host);
            const int64_t lo = ix->doc_base, n = ix->n_docs;
            std::vector<uint32_t> local((size_t)((n + 31) / 32 + 1), 0u);
            const int sh = (int)(lo & 31);
            const int64_t w0 = lo >> 5, n_src = ((int64_t)g->n_docs + 31) / 32;
            for (int64_t i = 0; i < (n + 31) / 32; ++i) {
                const uint32_t a = (w0 + i < n_src) ? w[w0 + i] : 0u, b = (w0 + i + 1 < n_src) ? w[w0 + i + 1] : 0u;
                local[(size_t)i] = sh ? ((a >> sh) | (b << (32 - sh))) : a;
            }
            rc = bm25cu_index_set_mask(ix, kind, local.data());
        } else if (kind == BM25CU_MASK_F64) {
            rc = bm25cu_index_set_mask(ix, kind, static_cast<const double *>(mask_host) + ix->doc_base);
        } else {
            rc = bm25cu_index_set_mask(ix, kind, mask_host);
        }
        if (rc) return rc;
    }
    return BM25CU_OK;
}

int bm25cu_group_unset_option(bm25cu_group *g, const char *name) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_unset_option: null handle");
    for (auto *ix : g->shard) {
        int rc = bm25cu_index_unset_option(ix, name);
        if (rc) return rc;
    }
    return BM25CU_OK;
}

// The whole matrix back in the reference's layout (global document ids): column t = the shards' slices of t in shard order.
int bm25cu_group_download(const bm25cu_group *g, float *data, int32_t *indices, int64_t *indptr, float *nonocc) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_download: null handle");
    if (g->n == 1) return bm25cu_index_download(g->shard[0], data, indices, indptr, nonocc);
    const int32_t V = g->n_vocab;
    std::vector<std::vector<int64_t>> sp((size_t)g->n);
    for (int r = 0; r < g->n; ++r) {
        sp[(size_t)r].resize((size_t)V + 1);
        int rc = bm25cu_index_download(g->shard[(size_t)r], nullptr, nullptr, sp[(size_t)r].data(), r == 0 ? nonocc : nullptr);
        if (rc) return rc;
    }
    std::vector<int64_t> gp((size_t)V + 1);
    gp[0] = 0;
    for (int32_t t = 0; t < V; ++t) {
        int64_t n = 0;
        for (int r = 0; r < g->n; ++r) n += sp[(size_t)r][(size_t)t + 1] - sp[(size_t)r][(size_t)t];
        gp[(size_t)t + 1] = gp[(size_t)t] + n;
    }
    if (indptr) std::memcpy(indptr, gp.data(), ((size_t)V + 1) * sizeof(int64_t));
    if (!data && !indices) return BM25CU_OK;
    std::vector<int64_t> before((size_t)V, 0);  // postings of the lower shards in front of a column's slice
    for (int r = 0; r < g->n; ++r) {
        const bm25cu_index *ix = g->shard[(size_t)r];
        std::vector<float> sd((size_t)(ix->nnz > 0 ? ix->nnz : 1));
        std::vector<int32_t> si((size_t)(ix->nnz > 0 ? ix->nnz : 1));
        int rc = bm25cu_index_download(ix, data ? sd.data() : nullptr, indices ? si.data() : nullptr, nullptr, nullptr);
        if (rc) return rc;
        const auto &p = sp[(size_t)r];
        for (int32_t t = 0; t < V; ++t) {
            const int64_t n = p[(size_t)t + 1] - p[(size_t)t], dst = gp[(size_t)t] + before[(size_t)t];
            if (n == 0) continue;
            if (data) std::memcpy(data + dst, sd.data() + p[(size_t)t], (size_t)n * sizeof(float));
            if (indices)
                for (int64_t j = 0; j < n; ++j) indices[dst + j] = si[(size_t)(p[(size_t)t] + j)] + ix->doc_base;
            before[(size_t)t] += n;
        }
    }
    return BM25CU_OK;
}

int bm25cu_group_retrieve(bm25cu_group *g, const int32_t *q_tokens, const int64_t *q_ptr, int32_t n_queries, int32_t k,
                          const float *shift_per_query, const double *weight_mask, int32_t *out_ids, float *out_scores,
                          bm25cu_stats *stats) {
    BM25CU_REQUIRE(g != nullptr, "bm25cu_group_retrieve: null handle");
    if (g->n == 1)
        return bm25cu_retrieve(g->shard[0], q_tokens, q_ptr, n_queries, k, 1, shift_per_query, weight_mask, out_ids, out_scores, stats);
    BM25CU_REQUIRE(n_queries >= 0 && k >= 0 && q_ptr != nullptr, "bm25cu_group_retrieve: bad sizes / null q_ptr");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    BM25CU_REQUIRE(out_ids != nullptr && out_scores != nullptr, "bm25cu_group_retrieve: null output");
    const int64_t n_tok = q_ptr[n_queries];
    BM25CU_REQUIRE(q_ptr[0] == 0 && n_tok >= 0 && (n_tok == 0 || q_tokens != nullptr), "bm25cu_group_retrieve: bad q_ptr / null q_tokens");
    for (int32_t i = 0; i < n_queries; ++i) BM25CU_REQUIRE(q_ptr[i + 1] >= q_ptr[i], "bm25cu_group_retrieve: q_ptr must be non-decreasing");
    for (int64_t i = 0; i < n_tok; ++i) {
        if (q_tokens[i] < 0 || q_tokens[i] >= g->n_vocab) {
            char buf[256];
            snprintf(buf, sizeof(buf), "The maximum token ID in the query (%d) is higher than the number of tokens in the index.", q_tokens[i]);
            set_error(buf);
            return BM25CU_EINVAL;
        }
    }
    const int root = g->device[0];
    const size_t rows = (size_t)n_queries * k;
    const size_t tok_bytes = (size_t)(n_tok > 0 ? n_tok : 1) * sizeof(int32_t), ptr_bytes = (size_t)(n_queries + 1) * sizeof(int64_t);
    const size_t off_tok = ptr_bytes, off_shift = off_tok + ((tok_bytes + 15) / 16) * 16;
    int rc;
    BM25CU_CHECK(cudaSetDevice(root));
    if ((rc = g->h_in.reserve(off_shift + (size_t)n_queries * 4 + 16))) return rc;
    if ((rc = g->h_out.reserve(rows * 8))) return rc;
    if ((rc = g->g_ids.reserve((size_t)g->n * rows * 4))) return rc;
    if ((rc = g->g_scores.reserve((size_t)g->n * rows * 4))) return rc;
    if ((rc = g->m_ids.reserve(rows * 4))) return rc;
    if ((rc = g->m_scores.reserve(rows * 4))) return rc;
    if (shift_per_query && (rc = g->d_shift.reserve((size_t)n_queries * 4))) return rc;
    char *hin = g->h_in.as<char>();
    std::memcpy(hin, q_ptr, ptr_bytes);
    if (n_tok) std::memcpy(hin + off_tok, q_tokens, (size_t)n_tok * sizeof(int32_t));
    if (shift_per_query) std::memcpy(hin + off_shift, shift_per_query, (size_t)n_queries * 4);
    const bool timing = g->shard[0]->knobs.get("TIME", 0) != 0;  // debugging: host time of the phases of a call
    auto tick = []() { return std::chrono::steady_clock::now(); };
    auto t0 = tick();
    // every shard: queries up, scoring kernels (no shift: it is added after the merge), rows over to the first device
    rc = g->pool->run([&](int r) {
        bm25cu_index *ix = g->shard[(size_t)r];
        BM25CU_CHECK(cudaSetDevice(ix->device));
        int rc2;
        if ((rc2 = ix->d_qtok.reserve(tok_bytes))) return rc2;
        if ((rc2 = ix->d_qptr.reserve(ptr_bytes))) return rc2;
        if ((rc2 = ix->d_out_ids.reserve(rows * 4))) return rc2;
        if ((rc2 = ix->d_out_scores.reserve(rows * 4))) return rc2;
        BM25CU_CHECK(cudaMemcpyAsync(ix->d_qptr.p, hin, ptr_bytes, cudaMemcpyHostToDevice, ix->stream));
        if (n_tok) BM25CU_CHECK(cudaMemcpyAsync(ix->d_qtok.p, hin + off_tok, (size_t)n_tok * 4, cudaMemcpyHostToDevice, ix->stream));
        const double *d_mask = nullptr;
        if (weight_mask) {
            if ((rc2 = ix->d_mask.reserve((size_t)(ix->n_docs > 0 ? ix->n_docs : 1) * sizeof(double)))) return rc2;
            BM25CU_CHECK(cudaMemcpyAsync(ix->d_mask.p, weight_mask + ix->doc_base, (size_t)ix->n_docs * sizeof(double),
                                         cudaMemcpyHostToDevice, ix->stream));
            d_mask = ix->d_mask.as<double>();
        }
        if ((rc2 = bm25cu_retrieve_device_enqueue(ix, ix->d_qtok.as<int32_t>(), ix->d_qptr.as<int64_t>(), n_queries, n_tok, k, nullptr,
                                                  d_mask, ix->d_out_ids.as<int32_t>(), ix->d_out_scores.as<float>())))
            return rc2;
        BM25CU_CHECK(cudaMemcpyPeerAsync(g->g_ids.as<int32_t>() + (size_t)r * rows, root, ix->d_out_ids.p, ix->device, rows * 4, ix->stream));
        BM25CU_CHECK(cudaMemcpyPeerAsync(g->g_scores.as<float>() + (size_t)r * rows, root, ix->d_out_scores.p, ix->device, rows * 4, ix->stream));
        BM25CU_CHECK(cudaEventRecord(g->done[(size_t)r], ix->stream));
        return BM25CU_OK;
    });
    if (rc) return rc;
    auto t1 = tick();
    // the first device: wait for every shard's rows, merge, rows down
    BM25CU_CHECK(cudaSetDevice(root));
    cudaStream_t st = g->shard[0]->stream;
    for (int r = 1; r < g->n; ++r) BM25CU_CHECK(cudaStreamWaitEvent(st, g->done[(size_t)r], 0));
    if (shift_per_query)
        BM25CU_CHECK(cudaMemcpyAsync(g->d_shift.p, hin + off_shift, (size_t)n_queries * 4, cudaMemcpyHostToDevice, st));
    if ((rc = topk_merge_device(g->g_ids.as<int32_t>(), g->g_scores.as<float>(), g->n, n_queries, k,
                                shift_per_query ? g->d_shift.as<float>() : nullptr, g->m_ids.as<int32_t>(), g->m_scores.as<float>(), st)))
        return rc;
    char *hout = g->h_out.as<char>();
    BM25CU_CHECK(cudaMemcpyAsync(hout, g->m_ids.p, rows * 4, cudaMemcpyDeviceToHost, st));
    BM25CU_CHECK(cudaMemcpyAsync(hout + rows * 4, g->m_scores.p, rows * 4, cudaMemcpyDeviceToHost, st));
    auto t2 = tick();
    // errors / stats of the shards (synchronises their streams), then the first device's stream
    bm25cu_stats acc;
    std::memset(&acc, 0, sizeof(acc));
    for (int r = 0; r < g->n; ++r) {
        bm25cu_stats s1;
        if ((rc = bm25cu_retrieve_collect(g->shard[(size_t)r], n_queries, n_tok, k, &s1))) return rc;
        acc.kernel_ms = std::max(acc.kernel_ms, s1.kernel_ms);
        acc.fast_kernel_ms = std::max(acc.fast_kernel_ms, s1.fast_kernel_ms);
        acc.ms_kernel_ms = std::max(acc.ms_kernel_ms, s1.ms_kernel_ms);
        acc.total_ms = std::max(acc.total_ms, s1.total_ms);
        acc.posting_bytes += s1.posting_bytes;
        acc.n_general = std::max(acc.n_general, s1.n_general);
        acc.launches += s1.launches;
    }
    BM25CU_CHECK(cudaSetDevice(root));
    BM25CU_CHECK(cudaStreamSynchronize(st));
    if (timing) {
        auto t3 = tick();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[bm25cu group] enqueue on %d shards %.3f ms, merge enqueue %.3f ms, collect + sync %.3f ms (kernels %.3f ms)\n", g->n,
                ms(t0, t1), ms(t1, t2), ms(t2, t3), acc.kernel_ms);
    }
    acc.launches += 1;  // the merge
    acc.n_fast = n_queries - acc.n_general;
    acc.algorithmic_bytes = acc.posting_bytes + 20 * n_tok + (int64_t)8 * k * n_queries;
    if (stats) *stats = acc;
    std::memcpy(out_ids, hout, rows * 4);
    std::memcpy(out_scores, hout + rows * 4, rows * 4);
    return BM25CU_OK;
}

// Dense scores of one query over the whole matrix: every shard fills its document range of `out`.
int bm25cu_group_scores_dense(bm25cu_group *g, const int32_t *q_tokens, int32_t n_tokens, int32_t has_shift, float shift,
                              const double *weight_mask, float *out) {
    BM25CU_REQUIRE(g != nullptr && out != nullptr, "bm25cu_group_scores_dense: null pointer");
    return g->pool->run([&](int r) {
        const bm25cu_index *ix = g->shard[(size_t)r];
        return bm25cu_scores_dense(g->shard[(size_t)r], q_tokens, n_tokens, has_shift, shift,
                                   weight_mask ? weight_mask + ix->doc_base : nullptr, out + ix->doc_base);
    });
}

int bm25cu_group_free(bm25cu_group *g) {
    group_destroy(g);
    return BM25CU_OK;
}

}  // extern "C"
