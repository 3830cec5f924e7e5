// Index residency: upload of the reference's CSC arrays (bm25s/__init__.py:432-438, :1121-1127) into HBM,
// doc-range shard extraction on the device, download, free.
#include "common.cuh"
#include "select.cuh"
#include <chrono>
#include <mutex>
#include <thread>
#include <cstdlib>

#include "scan.cuh"

namespace bm25cu {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

extern "C" char **environ;
void Knobs::load_env() {
    for (char **e = environ; e && *e; ++e) {
        if (std::strncmp(*e, "BM25CU_", 7) != 0) continue;
        const char *eq = std::strchr(*e, '=');
        if (!eq || !eq[1]) continue;
        set(std::string(*e + 7, (size_t)(eq - (*e + 7))).c_str(), atoi(eq + 1));
    }
}

// per column: [first posting with doc >= doc_lo, first posting with doc >= doc_hi)
__global__ void shard_bounds_kernel(const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr,
                                    int32_t n_vocab, int32_t doc_lo, int32_t doc_hi, int64_t *start, int64_t *count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vocab) return;
    const int64_t s = indptr[t], e = indptr[t + 1];
    int64_t lo = s, hi = e;
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (indices[m] < doc_lo) lo = m + 1; else hi = m;
    }
    const int64_t a = lo;
    hi = e;
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (indices[m] < doc_hi) lo = m + 1; else hi = m;
    }
    start[t] = a;
    count[t] = lo - a;
}

// one warp per column copies its in-shard slice, rebasing doc ids to shard-local ids
__global__ void shard_copy_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                  const int64_t *__restrict__ start, const int64_t *__restrict__ new_indptr,
                                  int32_t n_vocab, int32_t doc_lo, float *out_data, int32_t *out_indices) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const int64_t src = start[warp], dst = new_indptr[warp], cnt = new_indptr[warp + 1] - dst;
    for (int64_t i = lane; i < cnt; i += 32) {
        out_data[dst + i] = data[src + i];
        out_indices[dst + i] = indices[src + i] - doc_lo;
    }
}

// Structural validation of an uploaded CSC (one warp per column): flags[0] |= 1 indptr not non-decreasing inside
// [0, nnz], |= 2 a document id outside [0, n_docs), |= 4 rows of a column not strictly ascending. The kernels rely on
// all three (lower bounds, bucket tables, presence bitmaps, scatter into row[indices[j]]); the reference would raise
// IndexError or return garbage for such a matrix, here the upload is refused.
__global__ void validate_csc_kernel(const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                    int64_t nnz, int32_t n_docs, unsigned int *flags) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const int64_t s = indptr[warp], e = indptr[warp + 1];
    if (s < 0 || e < s || e > nnz) {
        if (lane == 0) atomicOr(flags, 1u);
        return;
    }
    unsigned int bad = 0u;
    for (int64_t i = s + lane; i < e; i += 32) {
        const int32_t d = indices[i];
        if ((unsigned)d >= (unsigned)n_docs) bad |= 2u;
        if (i + 1 < e && indices[i + 1] <= d) bad |= 4u;
    }
    if (bad) atomicOr(flags, bad);
}

int validate_csc(const int32_t *d_indices, const int64_t *d_indptr, int32_t n_vocab, int64_t nnz, int32_t n_docs,
                 cudaStream_t stream) {
    if (n_vocab <= 0) return BM25CU_OK;
    unsigned int *d_flag = nullptr, flag = 0;
    BM25CU_CHECK(cudaMalloc(&d_flag, sizeof(unsigned int)));
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(unsigned int), stream);
    if (e == cudaSuccess) {
        const int64_t threads = (int64_t)n_vocab * 32;
        validate_csc_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_indices, d_indptr, n_vocab, nnz, n_docs, d_flag);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_flag);
    BM25CU_CHECK(e);
    if (flag & 1u) { set_error("bm25cu_index_upload: indptr is not non-decreasing within [0, nnz]"); return BM25CU_EINVAL; }
    if (flag & 2u) { set_error("bm25cu_index_upload: indices holds a document id outside [0, n_docs)"); return BM25CU_EINVAL; }
    if (flag & 4u) { set_error("bm25cu_index_upload: the rows of a column are not strictly ascending"); return BM25CU_EINVAL; }
    return BM25CU_OK;
}

// flags[0] |= 1 if any score is negative, NaN or infinite
__global__ void check_nonneg_kernel(const float *__restrict__ data, int64_t nnz, unsigned int *flags) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        float v = data[i];
        bad |= !(v >= 0.0f && v < __int_as_float(0x7f800000));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
}

// colmax[t] = max of column t (0 for an empty column); one warp per column
__global__ void column_max_kernel(const float *__restrict__ data, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                  float *colmax) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const int64_t s = indptr[warp], e = indptr[warp + 1];
    float m = 0.0f;
    for (int64_t i = s + lane; i < e; i += 32) m = fmaxf(m, data[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) colmax[warp] = m;
}

// colkth[r][t] = the 10th (r = 0), 100th (r = 1), 1000th (r = 2) largest score of column t, 0 if the column is shorter.
// One warp per column: MSB-first radix select on order-preserving keys with a per-warp shared-memory histogram.
__global__ void __launch_bounds__(256) column_kth_kernel(const float *__restrict__ data, const int64_t *__restrict__ indptr,
                                                         int32_t n_vocab, float *colkth) {
    __shared__ unsigned int hist_all[8][256];
    const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= n_vocab) return;
    unsigned int *hist = hist_all[threadIdx.x >> 5];
    const int64_t s = indptr[col], e = indptr[col + 1];
    const int64_t n = e - s;
    const int ranks[3] = {10, 100, 1000};
    for (int r = 0; r < 3; ++r) {
        float out = 0.0f;
        if (n >= ranks[r]) {
            uint32_t prefix = 0, mask = 0;
            unsigned int need = (unsigned int)ranks[r];
            for (int pass = 3; pass >= 0; --pass) {
                const int shift = 8 * pass;
#pragma unroll
                for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0;
                __syncwarp();
                for (int64_t i = s + lane; i < e; i += 32) {
                    const uint32_t key = f32_key(data[i]);
                    if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
                }
                __syncwarp();
                // lane L owns bins 8 L .. 8 L + 7; the wanted bin is found from the top
                unsigned int cnt[8], lane_sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { cnt[j] = hist[lane * 8 + j]; lane_sum += cnt[j]; }
                unsigned int above = lane_sum;  // inclusive suffix sum over lanes >= L
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_down_sync(0xffffffffu, above, o); if (lane + o < 32) above += y; }
                above -= lane_sum;  // elements in bins of higher lanes
                const bool here = above < need && need <= above + lane_sum;
                int sel = 0;
                unsigned int acc = above;
                if (here) {
                    for (int j = 7; j >= 0; --j) {
                        if (acc + cnt[j] >= need) { sel = lane * 8 + j; break; }
                        acc += cnt[j];
                    }
                }
                const int src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
                sel = __shfl_sync(0xffffffffu, sel, src);
                acc = __shfl_sync(0xffffffffu, acc, src);
                prefix |= (uint32_t)sel << shift;
                mask |= 0xFFu << shift;
                need -= acc;
                __syncwarp();
            }
            out = __uint_as_float((prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix);
        }
        if (lane == 0) colkth[(size_t)r * n_vocab + col] = out;
    }
}

// ---- bucket tables of the long columns (lookups of retrieve_ms.cu) ------------------------------------------------
// A column with df >= kBtabMinDf gets a table over buckets of 2^sh consecutive document ids, sh chosen so that a bucket
// holds 8..16 postings on average: tab[b] = number of postings with doc < b * 2^sh, (n_docs >> sh) + 2 entries.
// A lookup of document d reads tab[d >> sh], tab[(d >> sh) + 1] and searches that slice of `indices` (one or two
// sectors) instead of log2(df) scattered ones. Size: at most nnz / 4 bytes. Derived at upload, not part of the index.
constexpr int kBtabMinDf = 16;

__device__ __forceinline__ int btab_shift_of(long long df, int n_docs, int avg) {
    int sh = 0;
    while (sh < 31 && ((df << sh) < (long long)avg * (long long)n_docs)) ++sh;  // smallest sh with df * 2^sh >= avg * n_docs
    return sh;
}

__global__ void btab_size_kernel(const int64_t *__restrict__ indptr, int32_t n_vocab, int32_t n_docs, int avg, long long *sizes,
                                 unsigned char *shifts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vocab) return;
    const long long df = indptr[t + 1] - indptr[t];
    long long sz = 0;
    int sh = 0;
    if (df >= kBtabMinDf) {
        sh = btab_shift_of(df, n_docs, avg);
        sz = (long long)(n_docs >> sh) + 2;
    }
    sizes[t] = sz;
    shifts[t] = (unsigned char)sh;
}

// one warp per column: every lane finds the lower bound of its bucket boundaries
__global__ void btab_fill_kernel(const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                 int32_t n_docs, const unsigned char *__restrict__ shifts, long long *offs /* in: scan, out: -1 for none */,
                                 unsigned int *tab) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const long long beg = indptr[warp], df = indptr[warp + 1] - beg;
    if (df < kBtabMinDf) {
        __syncwarp();
        if (lane == 0) offs[warp] = -1;
        return;
    }
    const int sh = shifts[warp];
    const long long nb = (long long)(n_docs >> sh) + 2;
    const long long off = offs[warp];
    const int32_t *col = indices + beg;
    for (long long b = lane; b < nb; b += 32) {
        const long long bound = b << sh;  // first document of bucket b
        long long lo = 0, hi = df;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((long long)col[mid] < bound) lo = mid + 1; else hi = mid;
        }
        tab[off + b] = (unsigned int)lo;
    }
}

static int build_bucket_tables(bm25cu_index *ix) {
    if (ix->n_vocab <= 0 || ix->nnz <= 0 || ix->n_docs <= 0) return BM25CU_OK;
    if (ix->knobs.get("BTAB", 1) == 0) return BM25CU_OK;
    long long *d_sizes = nullptr, *d_tmp = nullptr;
    const int V = ix->n_vocab;
    BM25CU_CHECK(cudaMalloc(&d_sizes, (size_t)V * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&d_tmp, scan_tmp_elems(V) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_btab_off, (size_t)(V + 1) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_btab_sh, (size_t)V));
    int avg = ix->knobs.get("BTAB_AVG", 8);  // a bucket holds avg .. 2 avg postings on average
    if (avg < 1) avg = 1;
    btab_size_kernel<<<(V + 255) / 256, 256, 0, ix->stream>>>(ix->d_indptr, V, ix->n_docs, avg, d_sizes, ix->d_btab_sh);
    BM25CU_CHECK(cudaGetLastError());
    BM25CU_CHECK(exclusive_scan<long long>(d_sizes, V, ix->d_btab_off, d_tmp, ix->stream));
    long long total = 0;
    BM25CU_CHECK(cudaMemcpyAsync(&total, ix->d_btab_off + V, sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    cudaFree(d_sizes);
    cudaFree(d_tmp);
    BM25CU_CHECK(cudaMalloc(&ix->d_btab, (size_t)(total > 0 ? total : 1) * sizeof(unsigned int)));
    const long long threads = (long long)V * 32;
    btab_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(ix->d_indices, ix->d_indptr, V, ix->n_docs,
                                                                                ix->d_btab_sh, ix->d_btab_off, ix->d_btab);
    BM25CU_CHECK(cudaGetLastError());
    ix->aux_bytes += (size_t)total * 4;
    ix->btab_len = total;
    return BM25CU_OK;
}

// ---- presence bitmaps of the long columns (negative lookups of retrieve_ms.cu in one load) ------------------------
// A column with df >= kBtabMinDf gets one bit per 2^s consecutive documents, "a posting of this column falls in here",
// with s the largest shift that keeps at least 16 bits per posting (PBM_BITS; s = 0, an exact bitmap, for columns denser than
// n_docs / 8). A document that is not in the column - the usual outcome of a lookup - is recognised by one load with
// probability >= exp(-1/16). Size: at most 4 bytes per posting (1.4 GB at the MS-MARCO shape). Derived at upload, not part of the saved index.
__device__ __forceinline__ int pbm_shift_of(long long df, int n_docs, int bits_per_posting) {
    int s = 0;
    while (s < 30 && (((long long)bits_per_posting * df) << (s + 1)) <= (long long)n_docs) ++s;
    return s;
}

__global__ void pbm_size_kernel(const int64_t *__restrict__ indptr, int32_t n_vocab, int32_t n_docs, int bits_per_posting,
                                long long *sizes, unsigned char *shifts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vocab) return;
    const long long df = indptr[t + 1] - indptr[t];
    long long sz = 0;
    int sh = 0;
    if (df >= kBtabMinDf) {
        sh = pbm_shift_of(df, n_docs, bits_per_posting);
        sz = ((((long long)(n_docs - 1) >> sh) + 32) >> 5) + 1;  // 32-bit words
    }
    sizes[t] = sz;
    shifts[t] = (unsigned char)sh;
}

__global__ void pbm_fill_kernel(const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                const unsigned char *__restrict__ shifts, long long *offs, unsigned int *bits) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const long long beg = indptr[warp], df = indptr[warp + 1] - beg;
    if (df < kBtabMinDf) {
        __syncwarp();
        if (lane == 0) offs[warp] = -1;
        return;
    }
    const int sh = shifts[warp];
    unsigned int *b = bits + offs[warp];
    for (long long i = lane; i < df; i += 32) {
        const unsigned int x = (unsigned int)indices[beg + i] >> sh;
        atomicOr(&b[x >> 5], 1u << (x & 31u));
    }
}

static int build_presence_bitmaps(bm25cu_index *ix) {
    if (ix->n_vocab <= 0 || ix->nnz <= 0 || ix->n_docs <= 0) return BM25CU_OK;
    if (ix->knobs.get("PBM", 1) == 0) return BM25CU_OK;
    long long *d_sizes = nullptr, *d_tmp = nullptr;
    const int V = ix->n_vocab;
    BM25CU_CHECK(cudaMalloc(&d_sizes, (size_t)V * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&d_tmp, scan_tmp_elems(V) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_pbm_off, (size_t)(V + 1) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_pbm_sh, (size_t)V));
    int bits = ix->knobs.get("PBM_BITS", 16);  // at least this many bits per posting (8: 2.34 ms, 16: 2.20 ms, 32: 2.23 ms per MS-MARCO batch)
    if (bits < 1) bits = 1;
    pbm_size_kernel<<<(V + 255) / 256, 256, 0, ix->stream>>>(ix->d_indptr, V, ix->n_docs, bits, d_sizes, ix->d_pbm_sh);
    BM25CU_CHECK(cudaGetLastError());
    BM25CU_CHECK(exclusive_scan<long long>(d_sizes, V, ix->d_pbm_off, d_tmp, ix->stream));
    long long total = 0;
    BM25CU_CHECK(cudaMemcpyAsync(&total, ix->d_pbm_off + V, sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    cudaFree(d_sizes);
    cudaFree(d_tmp);
    BM25CU_CHECK(cudaMalloc(&ix->d_pbm, (size_t)(total > 0 ? total : 1) * sizeof(unsigned int)));
    BM25CU_CHECK(cudaMemsetAsync(ix->d_pbm, 0, (size_t)(total > 0 ? total : 1) * sizeof(unsigned int), ix->stream));
    const long long threads = (long long)V * 32;
    pbm_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(ix->d_indices, ix->d_indptr, V, ix->d_pbm_sh,
                                                                               ix->d_pbm_off, ix->d_pbm);
    BM25CU_CHECK(cudaGetLastError());
    ix->aux_bytes += (size_t)total * 4;
    ix->pbm_len = total;
    return BM25CU_OK;
}

// ---- slot tables of the long columns (positive lookups of retrieve_ms.cu in ONE load) ------------------------------
// A column with df >= kBtabMinDf gets a direct-mapped table over buckets of 2^s consecutive document ids, s the largest
// shift that keeps the average bucket at `load` / 10 postings or fewer (default 3.0: between 1.5 and 3 per bucket).
// A bucket is 64 bytes: two 32-byte sectors, one per HALF of the bucket's document range, of four (doc id, score) slots
// each; empty slots hold id 0xffffffff. A document is looked up in the sector of its half - one load. A half with more
// than four postings keeps four and lends the rest to the free slots of the other sector (filled from the top): bit 31 of
// slot 0's id says "look at the other sector, too" (the same 64-byte line); if they did not fit there either, bit 31 of
// slot 1's id says "look in the column itself" (about one lookup in a hundred). Replaces the chain bucket table -> id
// slice -> score. Derived at upload, not part of the saved index; 21 .. 43 bytes per posting (HBM capacity traded for
// latency; knob STAB=0 drops the tables, STAB_LOAD sizes them).
constexpr int kStabSlots = 8;
constexpr int kStabScan = 24;  // postings of one bucket looked at by the fill kernel; longer buckets overflow to the column

__device__ __forceinline__ int stab_shift_of(long long df, int n_docs, int load10) {
    int s = 1;  // a bucket spans at least two documents: one per half
    while (s < 30 && ((df * 10) << (s + 1)) <= (long long)load10 * (long long)n_docs) ++s;
    return s;
}

__global__ void stab_size_kernel(const int64_t *__restrict__ indptr, int32_t n_vocab, int32_t n_docs, int load10, int compact,
                                 long long *sizes, unsigned char *shifts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_vocab) return;
    const long long df = indptr[t + 1] - indptr[t];
    long long sz = 0;
    int sh = 0;
    if (df >= kBtabMinDf) {
        sh = stab_shift_of(df, n_docs, load10);
        // compact format: buckets wider than 16-bit offsets reach get no table (rare columns, searched through their bucket tables)
        if (!compact || sh <= 16) sz = ((long long)(n_docs - 1) >> sh) + 1;  // buckets
    }
    sizes[t] = sz;
    shifts[t] = (unsigned char)sh;
}

// One warp per column, one lane per posting. Rows of a column ascend, so the postings of a bucket are neighbours and the
// postings of its lower half come first: a lane finds its bucket's extent, the two halves' counts and its own rank by
// looking at most kStabScan postings back and ahead, and derives its slot (or none) and the flags from those alone.
// Two sector formats, one per index (knob STAB_COMPACT):
//   wide    four slots, uint2 {doc id, score bits}; bit 31 of slot 0's id = "look at the other sector, too", bit 31 of slot
//           1's id = "look in the column itself"
//   compact (lossless compression by tile-local 16-bit doc offsets, applied where the bytes are) a doc id is
//           stored as its 16-bit offset inside the bucket (the bucket number is the rest of the id), which makes room for
//           FIVE slots per 32-byte sector: words 0..2 = offsets 0..4 and a 16-bit meta field, words 3..7 = the five scores.
//           meta is stored inverted (an untouched sector, all ones, decodes to "empty"): bits 0..4 = slot valid, bit 5 =
//           "look at the other sector, too", bit 6 = "look in the column itself".
__global__ void stab_fill_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices, const int64_t *__restrict__ indptr,
                                 int32_t n_vocab, const unsigned char *__restrict__ shifts, long long *offs /* in: scan, out: -1 for none */,
                                 const long long *__restrict__ sizes, unsigned int *slots, int compact) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_vocab) return;
    const long long beg = indptr[warp], df = indptr[warp + 1] - beg;
    if (sizes[warp] == 0) {  // short column, or buckets wider than 16-bit offsets reach: no table
        __syncwarp();
        if (lane == 0) offs[warp] = -1;
        return;
    }
    const int sh = shifts[warp];
    unsigned int *tab = slots + offs[warp] * 16;  // 16 words per bucket, sector h = words 8 h .. 8 h + 7
    const int32_t *col = indices + beg;
    for (long long i = lane; i < df; i += 32) {
        const int32_t d = col[i];
        const unsigned b = (unsigned)d >> sh;
        const int h = (int)(((unsigned)d >> (sh - 1)) & 1u);
        int before = 0, before_same = 0;  // postings of this bucket in front of this one; those of its own half
        while (before < kStabScan && i - before - 1 >= 0 && ((unsigned)col[i - before - 1] >> sh) == b) {
            if ((int)(((unsigned)col[i - before - 1] >> (sh - 1)) & 1u) == h) ++before_same;
            ++before;
        }
        int after = 0, after_same = 0;
        while (after < kStabScan && i + after + 1 < df && ((unsigned)col[i + after + 1] >> sh) == b) {
            if ((int)(((unsigned)col[i + after + 1] >> (sh - 1)) & 1u) == h) ++after_same;
            ++after;
        }
        const int S = compact ? 5 : 4;                                 // slots per sector
        const bool huge = before >= kStabScan || after >= kStabScan;  // extent unknown: everything past S per half overflows to the column
        const int mine = before_same + 1 + after_same;                 // postings of this half
        const int other = before + after + 1 - mine;                   // postings of the other half
        const int rank = before_same;
        const int room = other >= S ? 0 : S - other;                   // free slots of the other sector
        const int extra = mine > S ? mine - S : 0;
        const bool fits = !huge && extra <= room;
        unsigned int *bk = tab + (size_t)b * 16;
        const unsigned sbits = __float_as_uint(data[beg + i]);
        if (!compact) {
            uint2 *sec = reinterpret_cast<uint2 *>(bk);
            unsigned id = (unsigned)d;
            if (rank < 4) {
                if (rank == 0 && extra > 0) id |= 0x80000000u;           // look at the other sector, too
                if (rank == 1 && extra > 0 && !fits) id |= 0x80000000u;  // ... and in the column itself
                sec[h * 4 + rank] = make_uint2(id, sbits);
            } else if (!huge && rank - 4 < room) {
                sec[(1 - h) * 4 + 3 - (rank - 4)] = make_uint2(id, sbits);  // lent slot, from the top
            }
            continue;
        }
        const unsigned short off16 = (unsigned short)((unsigned)d & ((1u << sh) - 1u));
        unsigned short *own16 = reinterpret_cast<unsigned short *>(bk + h * 8), *sib16 = reinterpret_cast<unsigned short *>(bk + (1 - h) * 8);
        if (rank < 5) {
            own16[rank] = off16;
            bk[h * 8 + 3 + rank] = sbits;
        } else if (!huge && rank - 5 < room) {
            const int slot = 4 - (rank - 5);  // lent slot of the other sector, from the top
            sib16[slot] = off16;
            bk[(1 - h) * 8 + 3 + slot] = sbits;
        }
        if (rank == 0) {
            // the first posting of a half writes the sector's meta field: its own slots, the slots lent to it, the flags
            const int lent_in = huge ? 0 : min(max(other - 5, 0), 5 - min(mine, 5));
            unsigned meta = ((1u << min(mine, 5)) - 1u) | (((1u << lent_in) - 1u) << (5 - lent_in));
            if (extra > 0) meta |= 32u;
            if (extra > 0 && !fits) meta |= 64u;
            own16[5] = (unsigned short)~meta;
            if (other == 0 && extra > 0 && !huge) {  // the other half has no posting of its own: its meta is written here, too
                const int lent = min(extra, 5);
                sib16[5] = (unsigned short)~(((1u << lent) - 1u) << (5 - lent));
            }
        }
    }
}

static int build_slot_tables(bm25cu_index *ix) {
    if (ix->n_vocab <= 0 || ix->nnz <= 0 || ix->n_docs <= 0) return BM25CU_OK;
    if (ix->knobs.get("STAB", 1) == 0 || ix->d_btab == nullptr) return BM25CU_OK;  // overflowing buckets need the bucket tables
    long long *d_sizes = nullptr, *d_tmp = nullptr;
    const int V = ix->n_vocab;
    BM25CU_CHECK(cudaMalloc(&d_sizes, (size_t)V * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&d_tmp, scan_tmp_elems(V) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_stab_off, (size_t)(V + 1) * sizeof(long long)));
    BM25CU_CHECK(cudaMalloc(&ix->d_stab_sh, (size_t)V));
    int load10 = ix->knobs.get("STAB_LOAD", 30);  // average postings per bucket, times ten, at most
    if (load10 < 10) load10 = 10;
    if (load10 > 80) load10 = 80;
    ix->stab_compact = ix->knobs.get("STAB_COMPACT", 0) != 0;
    stab_size_kernel<<<(V + 255) / 256, 256, 0, ix->stream>>>(ix->d_indptr, V, ix->n_docs, load10, ix->stab_compact ? 1 : 0, d_sizes, ix->d_stab_sh);
    BM25CU_CHECK(cudaGetLastError());
    BM25CU_CHECK(exclusive_scan<long long>(d_sizes, V, ix->d_stab_off, d_tmp, ix->stream));
    long long total = 0;
    BM25CU_CHECK(cudaMemcpyAsync(&total, ix->d_stab_off + V, sizeof(long long), cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    cudaFree(d_tmp);
    const size_t bytes = (size_t)(total > 0 ? total : 1) * kStabSlots * 8;
    cudaError_t e = cudaMalloc(&ix->d_stab, bytes);
    if (e != cudaSuccess) {  // not enough HBM for the tables: lookups take the bucket-table path
        cudaGetLastError();
        cudaFree(d_sizes);
        cudaFree(ix->d_stab_off); ix->d_stab_off = nullptr;
        cudaFree(ix->d_stab_sh); ix->d_stab_sh = nullptr;
        ix->d_stab = nullptr;
        return BM25CU_OK;
    }
    BM25CU_CHECK(cudaMemsetAsync(ix->d_stab, 0xff, bytes, ix->stream));
    const long long threads = (long long)V * 32;
    stab_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(ix->d_data, ix->d_indices, ix->d_indptr, V, ix->d_stab_sh,
                                                                                ix->d_stab_off, d_sizes, reinterpret_cast<unsigned int *>(ix->d_stab),
                                                                                ix->stab_compact ? 1 : 0);
    BM25CU_CHECK(cudaGetLastError());
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    cudaFree(d_sizes);
    ix->aux_bytes += bytes;
    ix->stab_len = total;
    return BM25CU_OK;
}

int finalize_index(bm25cu_index *ix) {
    ix->knobs.load_env();  // the one place the BM25CU_* environment is read for this handle
    cudaDeviceProp prop;
    BM25CU_CHECK(cudaGetDeviceProperties(&prop, ix->device));
    ix->sm_count = prop.multiProcessorCount;
    ix->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (int i = 0; i < 6; ++i) BM25CU_CHECK(cudaEventCreate(&ix->ev[i]));
    unsigned int *d_flag = nullptr;
    BM25CU_CHECK(cudaMalloc(&d_flag, sizeof(unsigned int)));
    BM25CU_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(unsigned int), ix->stream));
    if (ix->nnz > 0) {
        check_nonneg_kernel<<<ix->sm_count * 4, 256, 0, ix->stream>>>(ix->d_data, ix->nnz, d_flag);
        BM25CU_CHECK(cudaGetLastError());
    }
    BM25CU_CHECK(cudaMalloc(&ix->d_colmax, (size_t)(ix->n_vocab > 0 ? ix->n_vocab : 1) * sizeof(float)));
    if (ix->n_vocab > 0) {
        const int64_t threads = (int64_t)ix->n_vocab * 32;
        column_max_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(ix->d_data, ix->d_indptr, ix->n_vocab, ix->d_colmax);
        BM25CU_CHECK(cudaGetLastError());
    }
    BM25CU_CHECK(cudaMalloc(&ix->d_colkth, (size_t)3 * (ix->n_vocab > 0 ? ix->n_vocab : 1) * sizeof(float)));
    if (ix->n_vocab > 0) {
        const int64_t threads = (int64_t)ix->n_vocab * 32;
        column_kth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(ix->d_data, ix->d_indptr, ix->n_vocab, ix->d_colkth);
        BM25CU_CHECK(cudaGetLastError());
    }
    {
        const bool timing = ix->knobs.get("TIME", 0) != 0;  // debugging: host time of the lookup-structure builds
        auto now = [&]() { if (timing) cudaStreamSynchronize(ix->stream); return std::chrono::steady_clock::now(); };
        auto t0 = now();
        int rc_bt = build_bucket_tables(ix);
        if (rc_bt) return rc_bt;
        auto t1 = now();
        if ((rc_bt = build_presence_bitmaps(ix))) return rc_bt;
        auto t2 = now();
        if ((rc_bt = build_slot_tables(ix))) return rc_bt;
        auto t3 = now();
        if (timing)
            fprintf(stderr, "[bm25cu] bucket tables %.1f ms, presence bitmaps %.1f ms, slot tables %.1f ms, %.1f MB\n",
                    std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(t2 - t1).count(),
                    std::chrono::duration<double, std::milli>(t3 - t2).count(), ix->aux_bytes / 1e6);
    }
    unsigned int flag = 0;
    BM25CU_CHECK(cudaMemcpyAsync(&flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost, ix->stream));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    cudaFree(d_flag);
    ix->nonneg = (flag == 0);
    return BM25CU_OK;
}

int alloc_postings(bm25cu_index *ix, int64_t nnz) {
    BM25CU_CHECK(cudaMalloc(&ix->d_data, (size_t)(nnz + kPadElems) * sizeof(float)));
    BM25CU_CHECK(cudaMalloc(&ix->d_indices, (size_t)(nnz + kPadElems) * sizeof(int32_t)));
    BM25CU_CHECK(cudaMemsetAsync(ix->d_data + nnz, 0, kPadElems * sizeof(float), ix->stream));
    BM25CU_CHECK(cudaMemsetAsync(ix->d_indices + nnz, 0x7f, kPadElems * sizeof(int32_t), ix->stream));
    return BM25CU_OK;
}

static void destroy(bm25cu_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->is_view) {  // a view owns its workspaces, stream and events only
        if (ix->root) ix->root->n_views.fetch_sub(1);
        ix->d_data = nullptr; ix->d_indices = nullptr; ix->d_indptr = nullptr; ix->d_nonocc = nullptr; ix->d_colmax = nullptr;
        ix->d_colkth = nullptr; ix->d_btab = nullptr; ix->d_btab_off = nullptr; ix->d_btab_sh = nullptr; ix->d_pbm = nullptr;
        ix->d_pbm_off = nullptr; ix->d_pbm_sh = nullptr; ix->d_stab = nullptr; ix->d_stab_off = nullptr; ix->d_stab_sh = nullptr;
    }
    cudaFree(ix->d_data);
    cudaFree(ix->d_indices);
    cudaFree(ix->d_indptr);
    cudaFree(ix->d_nonocc);
    cudaFree(ix->d_colmax);
    cudaFree(ix->d_colkth);
    cudaFree(ix->d_btab);
    cudaFree(ix->d_btab_off);
    cudaFree(ix->d_btab_sh);
    cudaFree(ix->d_pbm);
    cudaFree(ix->d_pbm_off);
    cudaFree(ix->d_pbm_sh);
    cudaFree(ix->d_stab);
    cudaFree(ix->d_stab_off);
    cudaFree(ix->d_stab_sh);
    bm25cu::DevBuf *bufs[] = {&ix->d_qtok, &ix->d_qptr, &ix->d_shift, &ix->d_mask, &ix->d_out_ids, &ix->d_out_scores,
                              &ix->d_ctrl, &ix->d_general_list, &ix->d_scratch, &ix->d_spill, &ix->d_order, &ix->d_theta, &ix->d_fb_list, &ix->d_ms_items, &ix->d_ms_rows, &ix->d_ms_q, &ix->d_ms_partial};
    for (auto *b : bufs) b->release();
    ix->d_rmask.release();
    ix->h_in.release();
    ix->h_out.release();
    for (int i = 0; i < 6; ++i)
        if (ix->ev[i]) cudaEventDestroy(ix->ev[i]);
    if (ix->stream && ix->owns_stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

}  // namespace bm25cu

using namespace bm25cu;

extern "C" {

const char *bm25cu_last_error(void) { return g_last_error.c_str(); }
int bm25cu_version(void) { return 100; }

int bm25cu_device_count(int32_t *n) {
    BM25CU_REQUIRE(n != nullptr, "bm25cu_device_count: null output");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *n = c;
    return BM25CU_OK;
}

// ---- host -> device copies through a pinned staging ring -----------------------------------------
// cudaMemcpy from pageable memory (numpy arrays, .npy memmaps) is staged by the driver on one thread at 8-10 GB/s.
// Here kStageThreads host threads each own two pinned buffers and a stream: a thread copies its chunks into pinned
// memory (page faults of a memmap are taken in parallel, too) while the DMA engine drains the other buffer.
// The ring (kStageThreads * 2 * kStageChunk bytes of pinned memory) is allocated once per process and device.
constexpr int kStageThreads = 4;
constexpr size_t kStageChunk = (size_t)8 << 20;

struct StageRing {
    int device = -1;
    char *buf[kStageThreads][2] = {};
    cudaStream_t st[kStageThreads] = {};
    cudaEvent_t ev[kStageThreads][2] = {};
    bool ok = false;
};
static std::mutex g_stage_mu;
static StageRing g_stage;

static bool stage_ring_ready(int device) {
    if (g_stage.ok && g_stage.device == device) return true;
    if (g_stage.ok) return false;  // one ring per process: uploads to another device take the plain path
    for (int t = 0; t < kStageThreads; ++t) {
        if (cudaStreamCreateWithFlags(&g_stage.st[t], cudaStreamNonBlocking) != cudaSuccess) return false;
        for (int b = 0; b < 2; ++b) {
            if (cudaHostAlloc((void **)&g_stage.buf[t][b], kStageChunk, cudaHostAllocDefault) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&g_stage.ev[t][b], cudaEventDisableTiming) != cudaSuccess) return false;
        }
    }
    g_stage.device = device;
    g_stage.ok = true;
    return true;
}

// Synchronous: returns when `bytes` bytes of host memory at `src` are in device memory at `dst`.
static cudaError_t staged_h2d(void *dst, const void *src, size_t bytes, int device, bool staged) {
    const bool want = staged && bytes >= 4 * kStageChunk;
    std::unique_lock<std::mutex> lock(g_stage_mu);
    if (!want || !stage_ring_ready(device)) {
        lock.unlock();
        cudaGetLastError();
        return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    }
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    cudaError_t errs[kStageThreads];
    std::thread th[kStageThreads];
    for (int t = 0; t < kStageThreads; ++t) {
        errs[t] = cudaSuccess;
        th[t] = std::thread([&, t]() {
            cudaError_t e = cudaSetDevice(device);
            int used[2] = {0, 0};
            for (size_t c = t, i = 0; c < n_chunks && e == cudaSuccess; c += kStageThreads, ++i) {
                const int b = (int)(i & 1);
                const size_t off = c * kStageChunk, n = bytes - off < kStageChunk ? bytes - off : kStageChunk;
                if (used[b]) e = cudaEventSynchronize(g_stage.ev[t][b]);  // the DMA out of this buffer has finished
                if (e != cudaSuccess) break;
                std::memcpy(g_stage.buf[t][b], static_cast<const char *>(src) + off, n);
                e = cudaMemcpyAsync(static_cast<char *>(dst) + off, g_stage.buf[t][b], n, cudaMemcpyHostToDevice, g_stage.st[t]);
                if (e == cudaSuccess) e = cudaEventRecord(g_stage.ev[t][b], g_stage.st[t]);
                used[b] = 1;
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stage.st[t]);
            errs[t] = e;
        });
    }
    cudaError_t res = cudaSuccess;
    for (int t = 0; t < kStageThreads; ++t) {
        th[t].join();
        if (errs[t] != cudaSuccess) res = errs[t];
    }
    return res;
}

static int upload_impl(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz,
                       int32_t n_vocab, int32_t n_docs, const float *nonocc, int32_t device, int32_t doc_lo,
                       int32_t doc_hi, bool src_on_device, bm25cu_index **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_index_upload: null output handle");
    *out = nullptr;
    BM25CU_REQUIRE(nnz >= 0 && n_vocab >= 0 && n_docs >= 0, "bm25cu_index_upload: negative size");
    BM25CU_REQUIRE(indptr != nullptr, "bm25cu_index_upload: null indptr");
    BM25CU_REQUIRE(nnz == 0 || (data != nullptr && indices != nullptr), "bm25cu_index_upload: null data/indices");
    BM25CU_REQUIRE(0 <= doc_lo && doc_lo <= doc_hi && doc_hi <= n_docs, "bm25cu_index_upload: bad doc range");
    if (!src_on_device)
        BM25CU_REQUIRE(indptr[0] == 0 && indptr[n_vocab] == nnz, "bm25cu_index_upload: indptr does not span [0, nnz]");
    BM25CU_CHECK(cudaSetDevice(device));
    bm25cu_index *ix = new bm25cu_index();
    ix->device = device;
    ix->n_vocab = n_vocab;
    ix->n_docs = doc_hi - doc_lo;
    ix->doc_base = doc_lo;
    const char *sv = getenv("BM25CU_STAGED_UPLOAD");  // read once per upload (the handle's knob table is filled in finalize_index)
    const bool staged = !(sv && *sv && atoi(sv) == 0);
    int rc = BM25CU_OK;
    auto fail = [&](int code) { destroy(ix); return code; };
#define TRY(expr)                                                       \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) {                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
            return fail(_e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA); \
        }                                                               \
    } while (0)
    TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    TRY(cudaMalloc(&ix->d_indptr, (size_t)(n_vocab + 1) * sizeof(int64_t)));
    if (nonocc) {
        TRY(cudaMalloc(&ix->d_nonocc, (size_t)(n_vocab > 0 ? n_vocab : 1) * sizeof(float)));
        TRY(cudaMemcpyAsync(ix->d_nonocc, nonocc, (size_t)n_vocab * sizeof(float), cudaMemcpyDefault, ix->stream));
    }
    const bool whole = (doc_lo == 0 && doc_hi == n_docs);
    if (whole) {
        ix->nnz = nnz;
        if ((rc = alloc_postings(ix, nnz))) return fail(rc);
        TRY(cudaMemcpyAsync(ix->d_indptr, indptr, (size_t)(n_vocab + 1) * sizeof(int64_t), cudaMemcpyDefault, ix->stream));
        if (nnz && src_on_device) {
            TRY(cudaMemcpyAsync(ix->d_data, data, (size_t)nnz * sizeof(float), cudaMemcpyDefault, ix->stream));
            TRY(cudaMemcpyAsync(ix->d_indices, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyDefault, ix->stream));
        } else if (nnz) {
            TRY(staged_h2d(ix->d_data, data, (size_t)nnz * sizeof(float), device, staged));
            TRY(staged_h2d(ix->d_indices, indices, (size_t)nnz * sizeof(int32_t), device, staged));
        }
        if ((rc = validate_csc(ix->d_indices, ix->d_indptr, n_vocab, nnz, n_docs, ix->stream))) return fail(rc);
    } else {
        // stage the full matrix on the device, cut out the doc range there
        float *f_data = nullptr;
        int32_t *f_indices = nullptr;
        int64_t *f_indptr = nullptr, *d_start = nullptr, *d_count = nullptr, *d_tmp = nullptr;
        auto cleanup = [&]() { cudaFree(f_data); cudaFree(f_indices); cudaFree(f_indptr); cudaFree(d_start); cudaFree(d_count); cudaFree(d_tmp); };
#define TRY2(expr)                                                      \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) {                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
            cleanup();                                                  \
            return fail(_e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA); \
        }                                                               \
    } while (0)
        TRY2(cudaMalloc(&f_data, (size_t)(nnz + 1) * sizeof(float)));
        TRY2(cudaMalloc(&f_indices, (size_t)(nnz + 1) * sizeof(int32_t)));
        TRY2(cudaMalloc(&f_indptr, (size_t)(n_vocab + 1) * sizeof(int64_t)));
        TRY2(cudaMalloc(&d_start, (size_t)(n_vocab + 1) * sizeof(int64_t)));
        TRY2(cudaMalloc(&d_count, (size_t)(n_vocab + 1) * sizeof(int64_t)));
        TRY2(cudaMalloc(&d_tmp, scan_tmp_elems(n_vocab) * sizeof(long long)));
        TRY2(cudaMemcpyAsync(f_indptr, indptr, (size_t)(n_vocab + 1) * sizeof(int64_t), cudaMemcpyDefault, ix->stream));
        if (nnz) {
            if (src_on_device) {
                TRY2(cudaMemcpyAsync(f_data, data, (size_t)nnz * sizeof(float), cudaMemcpyDefault, ix->stream));
                TRY2(cudaMemcpyAsync(f_indices, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyDefault, ix->stream));
            } else {
                TRY2(staged_h2d(f_data, data, (size_t)nnz * sizeof(float), device, staged));
                TRY2(staged_h2d(f_indices, indices, (size_t)nnz * sizeof(int32_t), device, staged));
            }
        }
        if ((rc = validate_csc(f_indices, f_indptr, n_vocab, nnz, n_docs, ix->stream))) { cleanup(); return fail(rc); }
        if (n_vocab > 0) {
            shard_bounds_kernel<<<(n_vocab + 255) / 256, 256, 0, ix->stream>>>(f_indices, f_indptr, n_vocab, doc_lo, doc_hi, d_start, d_count);
            TRY2(cudaGetLastError());
        }
        TRY2(exclusive_scan<long long>((const long long *)d_count, n_vocab, (long long *)ix->d_indptr, (long long *)d_tmp, ix->stream));
        int64_t new_nnz = 0;
        TRY2(cudaMemcpyAsync(&new_nnz, ix->d_indptr + n_vocab, sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
        TRY2(cudaStreamSynchronize(ix->stream));
        ix->nnz = new_nnz;
        if ((rc = alloc_postings(ix, new_nnz))) { cleanup(); return fail(rc); }
        if (n_vocab > 0 && new_nnz > 0) {
            const int64_t threads = (int64_t)n_vocab * 32;
            shard_copy_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ix->stream>>>(f_data, f_indices, d_start, ix->d_indptr, n_vocab, doc_lo, ix->d_data, ix->d_indices);
            TRY2(cudaGetLastError());
        }
        TRY2(cudaStreamSynchronize(ix->stream));
        cleanup();
#undef TRY2
    }
    TRY(cudaStreamSynchronize(ix->stream));
#undef TRY
    if ((rc = finalize_index(ix))) return fail(rc);
    *out = ix;
    return BM25CU_OK;
}

int bm25cu_index_upload(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz,
                        int32_t n_vocab, int32_t n_docs, const float *nonocc, int32_t device, int32_t doc_lo,
                        int32_t doc_hi, bm25cu_index **out) {
    return upload_impl(data, indices, indptr, nnz, n_vocab, n_docs, nonocc, device, doc_lo, doc_hi, false, out);
}

int bm25cu_index_upload_device(const float *data, const int32_t *indices, const int64_t *indptr, int64_t nnz,
                               int32_t n_vocab, int32_t n_docs, const float *nonocc, int32_t device, int32_t doc_lo,
                               int32_t doc_hi, bm25cu_index **out) {
    return upload_impl(data, indices, indptr, nnz, n_vocab, n_docs, nonocc, device, doc_lo, doc_hi, true, out);
}

// A second handle on the SAME resident shard: it shares the posting arrays and lookup structures of `src` (read-only) and
// owns a stream, events and per-call workspaces of its own, so that calls on the view and on `src` may be in flight at
// the same time (several batches on several streams). Free the view before `src`.
int bm25cu_index_view(const bm25cu_index *src, bm25cu_index **out) {
    BM25CU_REQUIRE(src != nullptr && out != nullptr, "bm25cu_index_view: null pointer");
    *out = nullptr;
    BM25CU_CHECK(cudaSetDevice(src->device));
    bm25cu_index *v = new bm25cu_index();
    v->is_view = true;
    v->root = src->root ? src->root : const_cast<bm25cu_index *>(src);
    v->root->n_views.fetch_add(1);
    v->device = src->device;
    v->n_vocab = src->n_vocab;
    v->n_docs = src->n_docs;
    v->doc_base = src->doc_base;
    v->nnz = src->nnz;
    v->d_data = src->d_data;
    v->d_indices = src->d_indices;
    v->d_indptr = src->d_indptr;
    v->d_nonocc = src->d_nonocc;
    v->d_colmax = src->d_colmax;
    v->d_colkth = src->d_colkth;
    v->d_btab = src->d_btab;
    v->d_btab_off = src->d_btab_off;
    v->d_btab_sh = src->d_btab_sh;
    v->d_pbm = src->d_pbm;
    v->d_pbm_off = src->d_pbm_off;
    v->d_pbm_sh = src->d_pbm_sh;
    v->d_stab = src->d_stab;
    v->d_stab_off = src->d_stab_off;
    v->d_stab_sh = src->d_stab_sh;
    v->stab_compact = src->stab_compact;
    v->btab_len = src->btab_len;
    v->pbm_len = src->pbm_len;
    v->stab_len = src->stab_len;
    v->aux_bytes = 0;
    v->nonneg = src->nonneg;
    v->sm_count = src->sm_count;
    v->max_smem_optin = src->max_smem_optin;
    v->knobs = src->knobs;
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&v->ev[i]);
    if (e != cudaSuccess) {
        set_error(std::string("bm25cu_index_view: ") + cudaGetErrorString(e));
        destroy(v);
        return BM25CU_ECUDA;
    }
    *out = v;
    return BM25CU_OK;
}

int bm25cu_index_set_stream(bm25cu_index *ix, void *stream) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_index_set_stream: null handle");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    BM25CU_CHECK(cudaStreamSynchronize(ix->stream));
    if (ix->owns_stream && ix->stream) cudaStreamDestroy(ix->stream);
    ix->stream = (cudaStream_t)stream;
    ix->owns_stream = false;
    return BM25CU_OK;
}

int bm25cu_index_sizes(const bm25cu_index *ix, int64_t *nnz, int32_t *n_vocab, int32_t *n_docs_local,
                       int32_t *doc_lo, int32_t *has_nonocc) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_index_sizes: null handle");
    if (nnz) *nnz = ix->nnz;
    if (n_vocab) *n_vocab = ix->n_vocab;
    if (n_docs_local) *n_docs_local = ix->n_docs;
    if (doc_lo) *doc_lo = ix->doc_base;
    if (has_nonocc) *has_nonocc = ix->d_nonocc != nullptr;
    return BM25CU_OK;
}

int bm25cu_index_download(const bm25cu_index *ix, float *data, int32_t *indices, int64_t *indptr, float *nonocc) {
    BM25CU_REQUIRE(ix != nullptr, "bm25cu_index_download: null handle");
    BM25CU_CHECK(cudaSetDevice(ix->device));
    if (data && ix->nnz) BM25CU_CHECK(cudaMemcpy(data, ix->d_data, (size_t)ix->nnz * sizeof(float), cudaMemcpyDeviceToHost));
    if (indices && ix->nnz) BM25CU_CHECK(cudaMemcpy(indices, ix->d_indices, (size_t)ix->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (indptr) BM25CU_CHECK(cudaMemcpy(indptr, ix->d_indptr, (size_t)(ix->n_vocab + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (nonocc) {
        BM25CU_REQUIRE(ix->d_nonocc != nullptr, "bm25cu_index_download: index has no nonoccurrence array");
        BM25CU_CHECK(cudaMemcpy(nonocc, ix->d_nonocc, (size_t)ix->n_vocab * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return BM25CU_OK;
}

int bm25cu_index_free(bm25cu_index *ix) {
    destroy(ix);
    return BM25CU_OK;
}

}  // extern "C"
