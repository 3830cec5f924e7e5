// GPU index build (K4-K8): replaces BM25.build_index_from_ids (bm25s/__init__.py:326-438).
//
//   pairs .......... one warp per document writes (token, doc) for every token occurrence
//   sort ........... stable LSD radix sort on the token id (8-bit digits); documents were ascending on input, so the
//                    result is ordered by (token, doc) - the CSC order of scoring.py:412-432
//   unique + tf .... run-length encoding of equal (token, doc) pairs = per-document Counter (scoring.py:238-243)
//   df / indptr .... column boundaries of the unique pairs (doc frequencies, scoring.py:28-57; cumsum :419-421)
//   idf / nonocc ... float64 libm log on the host for the V distinct columns (scoring.py:60-112, :178-219), exactly
//                    the reference's math.log, rounded to float32 once
//   scores ......... tfc variants in float64 with the reference's operation order and numpy >= 2 promotions
//                    (scoring.py:115-159, :292-307); compiled with -fmad=false so nothing is contracted into FMAs
// Compiled with -fmad=false (Makefile) because bit-parity of `data` depends on un-fused IEEE arithmetic.
#include <cmath>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

using namespace bm25cu;

struct bm25cu_builder {
    int device = 0;
    int32_t n_docs = 0;   // local
    int32_t n_vocab = 0;
    int64_t n_tokens = 0;
    int64_t nnz = 0;
    cudaStream_t stream = nullptr;
    int64_t *d_doc_ptr = nullptr;  // i64[n_docs + 1]
    uint32_t *d_utok = nullptr;    // u32[nnz] unique pairs, (token, doc) ascending
    uint32_t *d_udoc = nullptr;    // u32[nnz]
    uint32_t *d_utf = nullptr;     // u32[nnz] term frequency
    int64_t *d_indptr = nullptr;   // i64[n_vocab + 1] (local)
    int32_t *d_df = nullptr;       // i32[n_vocab] local document frequencies (summed across shards by the caller)
};

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // keys per block
constexpr int kSortWarps = kSortThreads / 32;

__global__ void make_pairs_kernel(const int32_t *__restrict__ tokens, const int64_t *__restrict__ doc_ptr, int32_t n_docs,
                                  int32_t n_vocab, uint32_t *tok_out, uint32_t *doc_out, unsigned int *err) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_docs) return;
    const int64_t s = doc_ptr[warp], e = doc_ptr[warp + 1];
    for (int64_t i = s + lane; i < e; i += 32) {
        const int32_t t = tokens[i];
        if ((uint32_t)t >= (uint32_t)n_vocab) atomicOr(err, 1u);
        tok_out[i] = (uint32_t)t;
        doc_out[i] = (uint32_t)warp;
    }
}

// ---- one pass of the LSD radix sort: per-tile digit histograms (digit-major layout for the scan)
__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, unsigned int *hist /*[256][n_tiles]*/,
                 int64_t n_tiles) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// ---- stable scatter: warp w of the tile owns items [w * 512, (w + 1) * 512) and walks them in order, 32 at a time
__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t *keys_out,
                    uint32_t *vals_out, int64_t n, int shift, const long long *__restrict__ offsets /*[256][n_tiles]*/,
                    int64_t n_tiles) {
    __shared__ unsigned int warp_cnt[kSortWarps][256];  // per-warp digit counts, then exclusive bases
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortTile / kSortWarps);
    uint32_t k[kSortItems], v[kSortItems];
    unsigned short rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + (int64_t)r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : 0xFFFFFFFFu;
        v[r] = valid ? vals_in[idx] : 0u;
        const unsigned int digit = (k[r] >> shift) & 0xFFu;
        // lanes with the same digit (invalid lanes are matched among themselves through bit 8)
        const unsigned int peers = __match_any_sync(0xffffffffu, valid ? digit : 256u);
        const unsigned int before = __popc(peers & ((1u << lane) - 1u));
        unsigned int run = 0;
        if (valid) run = warp_cnt[warp][digit];
        rank[r] = (unsigned short)(run + before);
        __syncwarp();
        if (valid && before == 0) warp_cnt[warp][digit] = run + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan across the warps of the tile, per digit; the tile's global base of each digit comes from the scan
    __shared__ long long tile_base[256];
    {
        const int d = threadIdx.x;  // 256 threads == 256 digits
        unsigned int run = 0;
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned int c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
        tile_base[d] = offsets[(int64_t)d * n_tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + (int64_t)r * 32 + lane;
        if (idx < n) {
            const unsigned int digit = (k[r] >> shift) & 0xFFu;
            const long long pos = tile_base[digit] + warp_cnt[warp][digit] + rank[r];
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
    }
}

__global__ void head_flags_kernel(const uint32_t *__restrict__ tok, const uint32_t *__restrict__ doc, int64_t n,
                                  unsigned char *flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0 || tok[i] != tok[i - 1] || doc[i] != doc[i - 1]) ? 1 : 0;
}

// unique pairs + run starts (pos[i] = number of heads before i)
__global__ void compact_unique_kernel(const uint32_t *__restrict__ tok, const uint32_t *__restrict__ doc,
                                      const unsigned char *__restrict__ flags, const long long *__restrict__ pos,
                                      int64_t n, uint32_t *utok, uint32_t *udoc, long long *run_start) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) {
            const long long p = pos[i];
            utok[p] = tok[i];
            udoc[p] = doc[i];
            run_start[p] = i;
        }
}

__global__ void run_length_kernel(const long long *__restrict__ run_start, int64_t nnz, int64_t n, uint32_t *utf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        utf[i] = (uint32_t)((i + 1 < nnz ? run_start[i + 1] : n) - run_start[i]);
}

// indptr[t] = first unique pair whose token is >= t
__global__ void column_starts_kernel(const uint32_t *__restrict__ utok, int64_t nnz, int32_t n_vocab, int64_t *indptr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t prev = (i == 0) ? -1 : (int64_t)utok[i - 1];
        const int64_t cur = (i == nnz) ? (int64_t)n_vocab : (int64_t)utok[i];
        for (int64_t t = prev + 1; t <= cur; ++t) indptr[t] = i;
    }
}

__global__ void diff_kernel(const int64_t *__restrict__ indptr, int32_t n_vocab, int32_t *df) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_vocab) df[t] = (int32_t)(indptr[t + 1] - indptr[t]);
}

// Term-frequency component, float64, the reference's operation order (scoring.py:115-159) under numpy >= 2
// promotion rules: tf is a float32 array, l_avg an np.float64, k1/b/delta Python floats. The only float32 operation
// is `tf * (k1 + 1)` (float32 array times weak Python scalar) in atire and bm25+.
__device__ __forceinline__ double tfc_value(int method, float tf32, double l_d, double l_avg, double k1, double b,
                                            double delta) {
    const double tf = (double)tf32;
    switch (method) {
        case BM25CU_ROBERTSON:
        case BM25CU_LUCENE:
            return tf / (k1 * ((1.0 - b) + b * l_d / l_avg) + tf);
        case BM25CU_ATIRE: {
            const float num = tf32 * (float)(k1 + 1.0);
            return (double)num / (tf + k1 * (1.0 - b + b * l_d / l_avg));
        }
        case BM25CU_BM25L: {
            const double c = tf / (1.0 - b + b * l_d / l_avg);
            return ((k1 + 1.0) * (c + delta)) / (k1 + c + delta);
        }
        default: {  // BM25CU_BM25PLUS
            const float num = (float)(k1 + 1.0) * tf32;
            const double den = k1 * (1.0 - b + b * l_d / l_avg) + tf;
            return ((double)num / den) + delta;
        }
    }
}

__global__ void score_kernel(const uint32_t *__restrict__ utok, const uint32_t *__restrict__ udoc,
                             const uint32_t *__restrict__ utf, const int64_t *__restrict__ doc_ptr,
                             const float *__restrict__ idf32, const float *__restrict__ nonocc32, int64_t nnz, int method,
                             double l_avg, double k1, double b, double delta, float *data, int32_t *indices) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t t = utok[i], d = udoc[i];
        const double l_d = (double)(doc_ptr[d + 1] - doc_ptr[d]);
        const double tfc = tfc_value(method, (float)utf[i], l_d, l_avg, k1, b, delta);
        double s = (double)idf32[t] * tfc;  // idf_array is float32, tfc float64 -> float64 product (scoring.py:292)
        if (nonocc32) s = s - (double)nonocc32[t];  // scores_doc -= nonoccurrence_array[voc_ind_doc] (scoring.py:296-297)
        data[i] = (float)s;
        indices[i] = (int32_t)d;
    }
}

double idf_host(int method, double df, double N) {
    switch (method) {
        case BM25CU_ROBERTSON: {
            double inner = (N - df + 0.5) / (df + 0.5);
            if (inner < 1.0) inner = 1.0;
            return std::log(inner);
        }
        case BM25CU_LUCENE: return std::log(1.0 + (N - df + 0.5) / (df + 0.5));
        case BM25CU_ATIRE: return std::log(N / df);
        case BM25CU_BM25L: return std::log((N + 1.0) / (df + 0.5));
        default: return std::log((N + 1.0) / df);
    }
}

// tfc at tf = 0 with l_d = l_avg, the Python-scalar evaluation of bm25s/__init__.py:370-384 / scoring.py:106-110
double tfc0_host(int method, double l_avg, double k1, double b, double delta) {
    if (method == BM25CU_BM25L) {
        const double c = 0.0 / (1.0 - b + b * l_avg / l_avg);
        return ((k1 + 1.0) * (c + delta)) / (k1 + c + delta);
    }
    const double num = (k1 + 1.0) * 0.0;
    const double den = k1 * (1.0 - b + b * l_avg / l_avg) + 0.0;
    return (num / den) + delta;
}

void free_builder(bm25cu_builder *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->d_doc_ptr);
    cudaFree(b->d_utok);
    cudaFree(b->d_udoc);
    cudaFree(b->d_utf);
    cudaFree(b->d_indptr);
    cudaFree(b->d_df);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int ceil_log2(uint32_t x) {
    int b = 0;
    while ((1ull << b) < x) ++b;
    return b;
}

}  // namespace

extern "C" {

int bm25cu_build_begin(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs_local, int32_t n_vocab,
                       int32_t device, int32_t tokens_on_device, bm25cu_builder **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_build_begin: null output");
    *out = nullptr;
    BM25CU_REQUIRE(n_docs_local >= 0 && n_vocab >= 0, "bm25cu_build_begin: negative size");
    BM25CU_REQUIRE(doc_ptr != nullptr, "bm25cu_build_begin: null doc_ptr");
    BM25CU_CHECK(cudaSetDevice(device));
    bm25cu_builder *b = new bm25cu_builder();
    b->device = device;
    b->n_docs = n_docs_local;
    b->n_vocab = n_vocab;
    uint32_t *tokA = nullptr, *tokB = nullptr, *docA = nullptr, *docB = nullptr;
    int32_t *d_tokens_tmp = nullptr;
    unsigned int *d_hist = nullptr, *d_err = nullptr;
    long long *d_off = nullptr, *d_scan_tmp = nullptr, *d_pos = nullptr, *d_run = nullptr;
    unsigned char *d_flags = nullptr;
    auto cleanup_tmp = [&]() {
        cudaFree(tokA); cudaFree(tokB); cudaFree(docA); cudaFree(docB); cudaFree(d_tokens_tmp); cudaFree(d_hist);
        cudaFree(d_err); cudaFree(d_off); cudaFree(d_scan_tmp); cudaFree(d_pos); cudaFree(d_run); cudaFree(d_flags);
    };
#define BTRY(expr)                                                                                   \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            set_error(std::string("bm25cu_build: ") + #expr + ": " + cudaGetErrorString(_e));        \
            cleanup_tmp();                                                                           \
            free_builder(b);                                                                         \
            return _e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA;                   \
        }                                                                                            \
    } while (0)
    BTRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    cudaStream_t st = b->stream;
    BTRY(cudaMalloc(&b->d_doc_ptr, (size_t)(n_docs_local + 1) * sizeof(int64_t)));
    BTRY(cudaMemcpyAsync(b->d_doc_ptr, doc_ptr, (size_t)(n_docs_local + 1) * sizeof(int64_t), cudaMemcpyDefault, st));
    int64_t n = 0;
    BTRY(cudaMemcpyAsync(&n, b->d_doc_ptr + n_docs_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BTRY(cudaStreamSynchronize(st));
    if (n < 0 || (n > 0 && tokens == nullptr)) {
        set_error("bm25cu_build_begin: bad doc_ptr / null tokens");
        cleanup_tmp();
        free_builder(b);
        return BM25CU_EINVAL;
    }
    b->n_tokens = n;
    const size_t ne = (size_t)(n > 0 ? n : 1);
    const int32_t *d_tokens = tokens;
    if (!tokens_on_device && n > 0) {
        BTRY(cudaMalloc(&d_tokens_tmp, ne * sizeof(int32_t)));
        BTRY(cudaMemcpyAsync(d_tokens_tmp, tokens, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        d_tokens = d_tokens_tmp;
    }
    BTRY(cudaMalloc(&tokA, ne * 4));
    BTRY(cudaMalloc(&docA, ne * 4));
    BTRY(cudaMalloc(&d_err, sizeof(unsigned int)));
    BTRY(cudaMemsetAsync(d_err, 0, sizeof(unsigned int), st));
    if (n > 0 && n_docs_local > 0) {
        const int64_t threads = (int64_t)n_docs_local * 32;
        make_pairs_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_tokens, b->d_doc_ptr, n_docs_local, n_vocab, tokA, docA, d_err);
        BTRY(cudaGetLastError());
    }
    unsigned int err = 0;
    BTRY(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, st));
    BTRY(cudaStreamSynchronize(st));
    if (d_tokens_tmp) { cudaFree(d_tokens_tmp); d_tokens_tmp = nullptr; }
    if (err) {
        set_error("bm25cu_build_begin: token id outside [0, n_vocab)");
        cleanup_tmp();
        free_builder(b);
        return BM25CU_EINVAL;
    }
    // ---- stable LSD radix sort by token id
    const int bits = ceil_log2((uint32_t)(n_vocab > 1 ? n_vocab : 2));
    const int passes = (bits + 7) / 8;
    const int64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    if (n > 0) {
        BTRY(cudaMalloc(&tokB, ne * 4));
        BTRY(cudaMalloc(&docB, ne * 4));
        BTRY(cudaMalloc(&d_hist, (size_t)256 * n_tiles * sizeof(unsigned int)));
        BTRY(cudaMalloc(&d_off, ((size_t)256 * n_tiles + 1) * sizeof(long long)));
        BTRY(cudaMalloc(&d_scan_tmp, scan_tmp_elems(256 * n_tiles) * sizeof(long long)));
        for (int pass = 0; pass < passes; ++pass) {
            const int shift = 8 * pass;
            sort_hist_kernel<<<(unsigned)n_tiles, kSortThreads, 0, st>>>(tokA, n, shift, d_hist, n_tiles);
            BTRY(cudaGetLastError());
            BTRY(exclusive_scan<unsigned int>(d_hist, 256 * n_tiles, d_off, d_scan_tmp, st));
            sort_scatter_kernel<<<(unsigned)n_tiles, kSortThreads, 0, st>>>(tokA, docA, tokB, docB, n, shift, d_off, n_tiles);
            BTRY(cudaGetLastError());
            std::swap(tokA, tokB);
            std::swap(docA, docB);
        }
        cudaFree(tokB); tokB = nullptr;
        cudaFree(docB); docB = nullptr;
        cudaFree(d_hist); d_hist = nullptr;
        cudaFree(d_off); d_off = nullptr;
        cudaFree(d_scan_tmp); d_scan_tmp = nullptr;
    }
    // ---- unique (token, doc) pairs with term frequencies
    int64_t nnz = 0;
    if (n > 0) {
        BTRY(cudaMalloc(&d_flags, ne));
        BTRY(cudaMalloc(&d_pos, (ne + 1) * sizeof(long long)));
        BTRY(cudaMalloc(&d_scan_tmp, scan_tmp_elems(n) * sizeof(long long)));
        head_flags_kernel<<<1184, 256, 0, st>>>(tokA, docA, n, d_flags);
        BTRY(cudaGetLastError());
        BTRY(exclusive_scan<unsigned char>(d_flags, n, d_pos, d_scan_tmp, st));
        BTRY(cudaMemcpyAsync(&nnz, d_pos + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        BTRY(cudaStreamSynchronize(st));
    }
    b->nnz = nnz;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    BTRY(cudaMalloc(&b->d_utok, nz * 4));
    BTRY(cudaMalloc(&b->d_udoc, nz * 4));
    BTRY(cudaMalloc(&b->d_utf, nz * 4));
    BTRY(cudaMalloc(&b->d_indptr, (size_t)(n_vocab + 1) * sizeof(int64_t)));
    BTRY(cudaMalloc(&b->d_df, (size_t)(n_vocab > 0 ? n_vocab : 1) * sizeof(int32_t)));
    if (n > 0) {
        BTRY(cudaMalloc(&d_run, nz * sizeof(long long)));
        compact_unique_kernel<<<1184, 256, 0, st>>>(tokA, docA, d_flags, d_pos, n, b->d_utok, b->d_udoc, d_run);
        BTRY(cudaGetLastError());
        run_length_kernel<<<1184, 256, 0, st>>>(d_run, nnz, n, b->d_utf);
        BTRY(cudaGetLastError());
    }
    column_starts_kernel<<<1184, 256, 0, st>>>(b->d_utok, nnz, n_vocab, b->d_indptr);
    BTRY(cudaGetLastError());
    if (n_vocab > 0) {
        diff_kernel<<<(n_vocab + 255) / 256, 256, 0, st>>>(b->d_indptr, n_vocab, b->d_df);
        BTRY(cudaGetLastError());
    }
    BTRY(cudaStreamSynchronize(st));
    cleanup_tmp();
#undef BTRY
    *out = b;
    return BM25CU_OK;
}

int bm25cu_build_df_device(bm25cu_builder *b, int32_t **df_device) {
    BM25CU_REQUIRE(b && df_device, "bm25cu_build_df_device: null pointer");
    *df_device = b->d_df;
    return BM25CU_OK;
}

int bm25cu_build_df_get(bm25cu_builder *b, int32_t *df_host) {
    BM25CU_REQUIRE(b && df_host, "bm25cu_build_df_get: null pointer");
    BM25CU_CHECK(cudaSetDevice(b->device));
    BM25CU_CHECK(cudaMemcpy(df_host, b->d_df, (size_t)b->n_vocab * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return BM25CU_OK;
}

int bm25cu_build_df_set(bm25cu_builder *b, const int32_t *df_host) {
    BM25CU_REQUIRE(b && df_host, "bm25cu_build_df_set: null pointer");
    BM25CU_CHECK(cudaSetDevice(b->device));
    BM25CU_CHECK(cudaMemcpy(b->d_df, df_host, (size_t)b->n_vocab * sizeof(int32_t), cudaMemcpyHostToDevice));
    return BM25CU_OK;
}

int bm25cu_build_finish(bm25cu_builder *b, int64_t n_docs_global, int64_t total_len_global, int32_t tfc_method,
                        int32_t idf_method, double k1, double bparam, double delta, int32_t doc_lo,
                        bm25cu_index **out) {
    BM25CU_REQUIRE(b && out, "bm25cu_build_finish: null pointer");
    *out = nullptr;
    BM25CU_REQUIRE(tfc_method >= 0 && tfc_method <= 4 && idf_method >= 0 && idf_method <= 4,
                   "bm25cu_build_finish: unknown method");
    BM25CU_REQUIRE(n_docs_global >= b->n_docs && doc_lo >= 0, "bm25cu_build_finish: bad global sizes");
    BM25CU_CHECK(cudaSetDevice(b->device));
    const int32_t V = b->n_vocab;
    // ---- idf / nonoccurrence tables on the host: libm log in float64, one float32 rounding (scoring.py:60-112)
    std::vector<int32_t> df((size_t)(V > 0 ? V : 1));
    if (V > 0) BM25CU_CHECK(cudaMemcpy(df.data(), b->d_df, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const bool has_nonocc = (tfc_method == BM25CU_BM25L || tfc_method == BM25CU_BM25PLUS);
    // np.array(lens).mean(): exact integer sum in float64 divided by the count (bm25s/__init__.py:357)
    const double l_avg = n_docs_global > 0 ? (double)total_len_global / (double)n_docs_global : 0.0;
    const double tfc0 = has_nonocc ? tfc0_host(tfc_method, l_avg, k1, bparam, delta) : 0.0;
    std::vector<float> idf32((size_t)(V > 0 ? V : 1), 0.0f), nonocc32;
    if (has_nonocc) nonocc32.assign((size_t)(V > 0 ? V : 1), 0.0f);
    for (int32_t t = 0; t < V; ++t) {
        if (df[t] != 0) {
            const double idf = idf_host(idf_method, (double)df[t], (double)n_docs_global);
            idf32[t] = (float)idf;
            if (has_nonocc) nonocc32[t] = (float)(idf * tfc0);  // unrounded float64 idf (scoring.py:106-110)
        }
    }
    bm25cu_index *ix = new bm25cu_index();
    ix->device = b->device;
    ix->n_vocab = V;
    ix->n_docs = b->n_docs;
    ix->doc_base = doc_lo;
    ix->nnz = b->nnz;
    float *d_idf = nullptr;
    int rc = BM25CU_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) rc = alloc_postings(ix, b->nnz);
    if (e == cudaSuccess && rc == BM25CU_OK) e = cudaMalloc(&ix->d_indptr, (size_t)(V + 1) * sizeof(int64_t));
    if (e == cudaSuccess && rc == BM25CU_OK) e = cudaMalloc(&d_idf, (size_t)(V > 0 ? V : 1) * sizeof(float));
    if (e == cudaSuccess && rc == BM25CU_OK && has_nonocc) e = cudaMalloc(&ix->d_nonocc, (size_t)(V > 0 ? V : 1) * sizeof(float));
    if (e == cudaSuccess && rc == BM25CU_OK) {
        e = cudaMemcpyAsync(ix->d_indptr, b->d_indptr, (size_t)(V + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, ix->stream);
        if (e == cudaSuccess && V > 0) e = cudaMemcpyAsync(d_idf, idf32.data(), (size_t)V * sizeof(float), cudaMemcpyHostToDevice, ix->stream);
        if (e == cudaSuccess && has_nonocc && V > 0)
            e = cudaMemcpyAsync(ix->d_nonocc, nonocc32.data(), (size_t)V * sizeof(float), cudaMemcpyHostToDevice, ix->stream);
        if (e == cudaSuccess && b->nnz > 0) {
            score_kernel<<<1184, 256, 0, ix->stream>>>(b->d_utok, b->d_udoc, b->d_utf, b->d_doc_ptr, d_idf,
                                                      has_nonocc ? ix->d_nonocc : nullptr, b->nnz, tfc_method, l_avg, k1,
                                                      bparam, delta, ix->d_data, ix->d_indices);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    }
    cudaFree(d_idf);
    if (e != cudaSuccess || rc != BM25CU_OK) {
        if (e != cudaSuccess) set_error(std::string("bm25cu_build_finish: ") + cudaGetErrorString(e));
        bm25cu_index_free(ix);
        return rc != BM25CU_OK ? rc : (e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA);
    }
    if ((rc = finalize_index(ix))) {
        bm25cu_index_free(ix);
        return rc;
    }
    *out = ix;
    return BM25CU_OK;
}

int bm25cu_build_free(bm25cu_builder *b) {
    free_builder(b);
    return BM25CU_OK;
}

int bm25cu_index_build(const int32_t *tokens, const int64_t *doc_ptr, int32_t n_docs, int32_t n_vocab,
                       int32_t tfc_method, int32_t idf_method, double k1, double bparam, double delta,
                       int32_t device, int32_t tokens_on_device, bm25cu_index **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_index_build: null output");
    *out = nullptr;
    bm25cu_builder *b = nullptr;
    int rc = bm25cu_build_begin(tokens, doc_ptr, n_docs, n_vocab, device, tokens_on_device, &b);
    if (rc) return rc;
    rc = bm25cu_build_finish(b, n_docs, b->n_tokens, tfc_method, idf_method, k1, bparam, delta, 0, out);
    free_builder(b);
    return rc;
}

}  // extern "C"
