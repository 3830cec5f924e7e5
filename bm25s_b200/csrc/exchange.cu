// Exchange of the shards' top-k rows over NVLink peer memory (doc-range sharding), fused with the merge.
//
// The reference is single-process: there is nothing to replace; this is the one exchange step the doc-range sharding of
// DESIGN.md section 5 adds. Instead of an NCCL all_gather followed by a merge kernel, every rank
//   1. PUSHES its local rows (ids int32[Q][k], scores float32[Q][k]) straight into slot [rank] of a buffer that lives in
//      the memory of EVERY peer (plain stores through pointers opened with cudaIpcOpenMemHandle: NVLink / NVSwitch peer
//      writes), fences them system-wide and raises a per-source flag on every peer to the call's epoch;
//   2. waits until the world flags reached the epoch (a one-warp kernel of its own, exchange_wait_kernel: no kernel with a
//      full grid ever spins) and runs the MERGE kernel on its own copy: a warp (k <= 32) or a CTA per query merges the
//      world rows (same order as bm25cu_topk_merge: score descending, id ascending).
// For k > 32 the rows are big (k = 1000: 56 MB per rank and batch) and merging ALL queries on EVERY rank would move and
// sort everything world times: there the exchange is SLICED - rank p receives only the rows of its slice of the queries
// (push), merges them and stores the merged rows into a "gathered" buffer on every peer (merge), and a small kernel copies
// the gathered rows out once every rank's slice has arrived: 2/world of the all-to-all traffic, 1/world of the merge work.
// Two kernels (three when sliced) plus the one-warp waits, no host round trip, no collective library call. The push runs on the caller's stream, the merge on a side
// stream of the exchange (behind an event): waiting for the slowest shard and merging overlap the scoring kernels of the
// caller's NEXT batch; bm25cu_exchange_check / _join make the caller's stream wait for the merge. Buffers are
// double-buffered by epoch parity; a rank's merge of call e ends by raising a "consumed" flag on every peer, and the
// push of call e + 2 - the next user of that parity - waits for all peers' flags, so nobody overwrites rows still being
// read (the wait is over before it starts unless a peer is a whole call behind).
#include "common.cuh"
#include "select.cuh"

namespace {

constexpr int kMaxWorld = 16;
constexpr size_t kFlagBytes = 512;  // unsigned int flags[2][kMaxWorld] (rows arrived), consumed[kMaxWorld] at word 64, padded
constexpr int kConsumedWord = 64;
constexpr int kGatheredWord = 96;   // gathered[2][kMaxWorld]: rank r's merged slice of the epoch has arrived (sliced form)

struct ExchangeView {
    int32_t *ids[2];   // [world][cap]
    float *scores[2];  // [world][cap]
    int32_t *g_ids[2];   // [cap] merged rows of all queries, every rank writes its slice (sliced form)
    float *g_scores[2];  // [cap]
    unsigned int *flags;  // [2][kMaxWorld]
};

__host__ __device__ inline ExchangeView view_of(void *base, int world, long long cap) {
    ExchangeView v;
    char *p = static_cast<char *>(base);
    v.flags = reinterpret_cast<unsigned int *>(p);
    p += kFlagBytes;
    for (int b = 0; b < 2; ++b) {
        v.ids[b] = reinterpret_cast<int32_t *>(p);
        p += (size_t)world * cap * 4;
        v.scores[b] = reinterpret_cast<float *>(p);
        p += (size_t)world * cap * 4;
    }
    for (int b = 0; b < 2; ++b) {
        v.g_ids[b] = reinterpret_cast<int32_t *>(p);
        p += (size_t)cap * 4;
        v.g_scores[b] = reinterpret_cast<float *>(p);
        p += (size_t)cap * 4;
    }
    return v;
}

// the queries rank p merges in the sliced form: [slice_lo(p), slice_lo(p + 1))
__host__ __device__ inline int slice_lo(int n_queries, int world, int p) { return (int)((long long)n_queries * p / world); }

struct PeerPtrs {
    void *p[kMaxWorld];
};

__global__ void __launch_bounds__(256) exchange_push_kernel(const int32_t *__restrict__ ids, const float *__restrict__ scores,
                                                            long long n, PeerPtrs peers, int rank, int world, long long cap,
                                                            int parity, unsigned int epoch, unsigned int *done_counter,
                                                            unsigned int *err, int sliced_queries, int k) {
    for (int pr = 0; pr < world; ++pr) {
        const ExchangeView v = view_of(peers.p[pr], world, cap);
        int32_t *di = v.ids[parity] + (size_t)rank * cap;
        float *ds = v.scores[parity] + (size_t)rank * cap;
        // sliced form (sliced_queries > 0): a peer gets the rows of the queries it merges, nothing else
        const long long lo = sliced_queries > 0 ? (long long)slice_lo(sliced_queries, world, pr) * k : 0;
        const long long hi = sliced_queries > 0 ? (long long)slice_lo(sliced_queries, world, pr + 1) * k : n;
        for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
            di[i] = ids[i];
            ds[i] = scores[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {  // last block: every block's rows are fenced - raise the flags
            *done_counter = 0u;
            __threadfence_system();
            for (int pr = 0; pr < world; ++pr) {
                const ExchangeView v = view_of(peers.p[pr], world, cap);
                volatile unsigned int *f = v.flags + parity * kMaxWorld + rank;
                *f = epoch;
            }
            __threadfence_system();
        }
    }
}

// The ONLY kernel that waits for another GPU: one warp, lane r polls the flag rank r raises in this rank's memory until it
// reached `epoch` (10 s without it sets `errbit`). It runs in the stream ahead of the kernel that needs the peers' data,
// so a wait never holds more than one warp of the GPU: kernels with full grids that spin can fill the SMs with waiting
// CTAs while the scoring kernel of another batch in flight - the one whose rows the PEERS wait for - finds no slot
// (seen with three batches in flight on two GPUs), a deadlock across GPUs no single stream order prevents.
__global__ void __launch_bounds__(32) exchange_wait_kernel(const unsigned int *flags, int world, unsigned int epoch, unsigned int *err,
                                                           unsigned int errbit) {
    if (threadIdx.x < world) {
        const volatile unsigned int *f = flags + threadIdx.x;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*f - epoch) < 0) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 10000000000ull) { atomicOr(err, errbit); break; }
        }
    }
    __threadfence_system();
}

// End of a merge kernel: the last block to finish tells every peer that this rank is done reading the rows of `epoch`.
__device__ __forceinline__ void raise_consumed(PeerPtrs peers, int rank, int world, long long cap, unsigned int epoch,
                                               unsigned int *merge_counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(merge_counter, 1u);
        if (prev == gridDim.x - 1) {
            *merge_counter = 0u;
            __threadfence_system();
            for (int pr = 0; pr < world; ++pr) {
                const ExchangeView v = view_of(peers.p[pr], world, cap);
                volatile unsigned int *c = v.flags + kConsumedWord + rank;
                *c = epoch;
            }
            __threadfence_system();
        }
    }
}

__device__ __forceinline__ unsigned long long xshfl64(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ unsigned long long xshfl64_xor(unsigned long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// k <= 32: one WARP per query. A row of a rank is loaded one pair per lane as a 64-bit key (order-preserving score key,
// ~id: score descending, id ascending), sorted in the warp (rows arrive sorted by their pre-shift scores; the sort makes
// the merge independent of that), and bitonic-merged into the running top-32 held one key per lane. The non-occurrence
// shift of BM25L / BM25+ is added AFTER the merge, so the order - and with it the rows - do not depend on the number of
// shards. Eight queries per CTA; the rows are there when the kernel starts (exchange_wait_kernel runs ahead of it).
__global__ void __launch_bounds__(256) exchange_merge_warp_kernel(void *local, int world, long long cap, int parity, unsigned int epoch,
                                                                  int n_queries, int k, const float *__restrict__ shift,
                                                                  int32_t *out_ids, float *out_scores, unsigned int *err,
                                                                  PeerPtrs peers, int rank, unsigned int *merge_counter) {
    const ExchangeView v = view_of(local, world, cap);
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= n_queries) { raise_consumed(peers, rank, world, cap, epoch, merge_counter); return; }
    const int32_t *ids = v.ids[parity];
    const float *scores = v.scores[parity];
    unsigned long long t = 0ull;
    for (int r = 0; r < world; ++r) {
        unsigned long long b = 0ull;
        if (lane < k) {
            const size_t src = (size_t)r * cap + (size_t)q * k + lane;
            const int32_t d = __ldcv(ids + src);  // written by a peer: never from a stale cache line
            if (d >= 0) b = ((unsigned long long)bm25cu::f32_key(__ldcv(scores + src)) << 32) | (unsigned int)(~d);
        }
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const unsigned long long o = xshfl64_xor(b, stride);
                const bool lower = (lane & stride) == 0, desc = (lane & size) == 0;
                b = (lower == desc) ? (b > o ? b : o) : (b < o ? b : o);
            }
        }
        const unsigned long long rb = xshfl64(b, 31 - lane);
        t = t > rb ? t : rb;  // bitonic: the 32 largest of the union
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) {
            const unsigned long long o = xshfl64_xor(t, stride);
            t = ((lane & stride) == 0) ? (t > o ? t : o) : (t < o ? t : o);
        }
    }
    if (lane < k) {
        int32_t d = -1;
        float sc = __int_as_float(0xff800000);
        if (t != 0ull) {
            const uint32_t kk = (uint32_t)(t >> 32);
            sc = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
            if (shift != nullptr) sc = __fadd_rn(sc, shift[q]);
            d = (int32_t)~(unsigned int)(t & 0xffffffffu);
        }
        out_ids[(size_t)q * k + lane] = d;
        out_scores[(size_t)q * k + lane] = sc;
    }
    raise_consumed(peers, rank, world, cap, epoch, merge_counter);
}

// k > 32: one CTA per query; the ranks' rows (sorted by their producers) are rank-merged one after the other in shared
// memory (select.cuh, block_merge_sorted_rows) - for k = 1000 and 8 ranks 16 k binary searches per query where a bitonic
// sort of all world * k pairs took 370 k compare-exchanges.
__global__ void __launch_bounds__(256) exchange_merge_kernel(void *local, int world, long long cap, int parity, unsigned int epoch,
                                                             int n_queries, int k, int kcap, const float *__restrict__ shift,
                                                             int32_t *out_ids, float *out_scores, unsigned int *err,
                                                             PeerPtrs peers, int rank, unsigned int *merge_counter) {
    extern __shared__ __align__(16) unsigned long long xm[];  // cur, nxt, row: kcap keys each
    const ExchangeView v = view_of(local, world, cap);
    const int q = blockIdx.x;
    const int32_t *ids = v.ids[parity];
    const float *scores = v.scores[parity];
    auto load = [&](int r, int j) -> unsigned long long {
        const size_t src = (size_t)r * cap + (size_t)q * k + j;
        const int32_t d = __ldcv(ids + src);  // written by a peer: never from a stale cache line
        return d >= 0 ? (((unsigned long long)bm25cu::f32_key(__ldcv(scores + src)) << 32) | (unsigned int)(~d)) : 0ull;
    };
    const unsigned long long *top = bm25cu::block_merge_sorted_rows(world, k, kcap, xm, xm + kcap, xm + 2 * kcap, load);
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
    raise_consumed(peers, rank, world, cap, epoch, merge_counter);
}

// Sliced form, k > 32. One CTA per query of THIS rank's slice: the ranks' rows are rank-merged (select.cuh) and the merged
// row is stored into the gathered buffer of every peer; the last CTA to finish raises gathered[rank] on every peer.
__global__ void __launch_bounds__(256) exchange_merge_slice_kernel(void *local, int world, long long cap, int parity, unsigned int epoch,
                                                                   int n_queries, int k, int kcap, const float *__restrict__ shift,
                                                                   unsigned int *err, PeerPtrs peers, int rank, unsigned int *slice_counter) {
    extern __shared__ __align__(16) unsigned long long xm[];  // cur, nxt, row: kcap keys each
    const ExchangeView v = view_of(local, world, cap);
    const int q = slice_lo(n_queries, world, rank) + (int)blockIdx.x;
    if (q < slice_lo(n_queries, world, rank + 1)) {
        const int32_t *ids = v.ids[parity];
        const float *scores = v.scores[parity];
        auto load = [&](int r, int j) -> unsigned long long {
            const size_t src = (size_t)r * cap + (size_t)q * k + j;
            const int32_t d = __ldcv(ids + src);  // written by a peer: never from a stale cache line
            return d >= 0 ? (((unsigned long long)bm25cu::f32_key(__ldcv(scores + src)) << 32) | (unsigned int)(~d)) : 0ull;
        };
        const unsigned long long *top = bm25cu::block_merge_sorted_rows(world, k, kcap, xm, xm + kcap, xm + 2 * kcap, load);
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            const unsigned long long t = top[j];
            int32_t d = -1;
            float sc = __int_as_float(0xff800000);
            if (t != 0ull) {
                const uint32_t kk = (uint32_t)(t >> 32);
                sc = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
                if (shift != nullptr) sc = __fadd_rn(sc, shift[q]);
                d = (int32_t)~(unsigned int)(t & 0xffffffffu);
            }
            for (int pr = 0; pr < world; ++pr) {
                const ExchangeView pv = view_of(peers.p[pr], world, cap);
                pv.g_ids[parity][(size_t)q * k + j] = d;
                pv.g_scores[parity][(size_t)q * k + j] = sc;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(slice_counter, 1u);
        if (prev == gridDim.x - 1) {
            *slice_counter = 0u;
            __threadfence_system();
            for (int pr = 0; pr < world; ++pr) {
                const ExchangeView pv = view_of(peers.p[pr], world, cap);
                volatile unsigned int *f = pv.flags + kGatheredWord + parity * kMaxWorld + rank;
                *f = epoch;
            }
            __threadfence_system();
        }
    }
}

// Sliced form: once every rank's merged slice has arrived, the gathered rows go to the caller's arrays.
__global__ void __launch_bounds__(256) exchange_copy_out_kernel(void *local, int world, long long cap, int parity, unsigned int epoch, long long n,
                                                                int32_t *out_ids, float *out_scores, unsigned int *err, PeerPtrs peers,
                                                                int rank, unsigned int *merge_counter) {
    const ExchangeView v = view_of(local, world, cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        out_ids[i] = __ldcv(v.g_ids[parity] + i);
        out_scores[i] = __ldcv(v.g_scores[parity] + i);
    }
    raise_consumed(peers, rank, world, cap, epoch, merge_counter);
}

}  // namespace

struct bm25cu_exchange {
    int device = 0, rank = 0, world = 1;
    long long cap = 0;  // elements per rank slot (>= n_queries * k of every call)
    void *local = nullptr;
    void *peer[kMaxWorld] = {nullptr};
    bool connected = false;
    unsigned int epoch = 0;
    unsigned int *d_counter = nullptr;  // [0] block counter of the push kernel, [1] error bits, [2] block counter of the merge kernel
    size_t bytes = 0;
    cudaStream_t side = nullptr;        // the merge runs here, behind ev_pushed; ev_merged marks its end
    cudaEvent_t ev_pushed = nullptr, ev_merged = nullptr;
    bool merged_pending = false;
};

using namespace bm25cu;

extern "C" {

int bm25cu_exchange_create(int32_t device, int32_t rank, int32_t world, int64_t max_elems, bm25cu_exchange **out) {
    BM25CU_REQUIRE(out != nullptr, "bm25cu_exchange_create: null out");
    BM25CU_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bm25cu_exchange_create: bad rank / world (at most 16 ranks)");
    BM25CU_REQUIRE(max_elems > 0, "bm25cu_exchange_create: max_elems must be positive");
    BM25CU_CHECK(cudaSetDevice(device));
    bm25cu_exchange *x = new bm25cu_exchange();
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->cap = (max_elems + 3) & ~3ll;
    x->bytes = kFlagBytes + (size_t)2 * 2 * world * x->cap * 4 + (size_t)2 * 2 * x->cap * 4;
    cudaError_t e = cudaMalloc(&x->local, x->bytes);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
    if (e == cudaSuccess) e = cudaMalloc(&x->d_counter, 4 * sizeof(unsigned int));  // push blocks, error bits, merge / copy-out blocks, slice-merge blocks
    if (e == cudaSuccess) e = cudaMemset(x->d_counter, 0, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x->ev_pushed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x->ev_merged, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error(std::string("bm25cu_exchange_create: ") + cudaGetErrorString(e));
        cudaFree(x->local);
        cudaFree(x->d_counter);
        if (x->side) cudaStreamDestroy(x->side);
        if (x->ev_pushed) cudaEventDestroy(x->ev_pushed);
        if (x->ev_merged) cudaEventDestroy(x->ev_merged);
        delete x;
        return e == cudaErrorMemoryAllocation ? BM25CU_ENOMEM : BM25CU_ECUDA;
    }
    x->peer[rank] = x->local;
    *out = x;
    return BM25CU_OK;
}

int bm25cu_exchange_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int bm25cu_exchange_handle(bm25cu_exchange *x, void *handle_out) {
    BM25CU_REQUIRE(x && handle_out, "bm25cu_exchange_handle: null pointer");
    BM25CU_CHECK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    BM25CU_CHECK(cudaIpcGetMemHandle(&h, x->local));
    std::memcpy(handle_out, &h, sizeof(h));
    return BM25CU_OK;
}

int bm25cu_exchange_connect(bm25cu_exchange *x, const void *handles) {
    BM25CU_REQUIRE(x && handles, "bm25cu_exchange_connect: null pointer");
    BM25CU_CHECK(cudaSetDevice(x->device));
    const char *hp = static_cast<const char *>(handles);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hp + (size_t)r * sizeof(h), sizeof(h));
        BM25CU_CHECK(cudaIpcOpenMemHandle(&x->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    x->connected = true;
    return BM25CU_OK;
}

int bm25cu_exchange_merge(bm25cu_exchange *x, const int32_t *ids_device, const float *scores_device, int32_t n_queries,
                          int32_t k, const float *shift_per_query_device, int32_t *out_ids_device, float *out_scores_device,
                          void *stream) {
    BM25CU_REQUIRE(x != nullptr, "bm25cu_exchange_merge: null handle");
    BM25CU_REQUIRE(x->connected || x->world == 1, "bm25cu_exchange_merge: bm25cu_exchange_connect has not been called");
    BM25CU_REQUIRE(n_queries >= 0 && k >= 0, "bm25cu_exchange_merge: negative size");
    if (n_queries == 0 || k == 0) return BM25CU_OK;
    const long long n = (long long)n_queries * k;
    BM25CU_REQUIRE(n <= x->cap, "bm25cu_exchange_merge: n_queries * k exceeds the capacity given to bm25cu_exchange_create");
    BM25CU_REQUIRE(ids_device && scores_device && out_ids_device && out_scores_device, "bm25cu_exchange_merge: null pointer");
    BM25CU_CHECK(cudaSetDevice(x->device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int epoch = ++x->epoch;
    const int parity = (int)(epoch & 1u);
    PeerPtrs pp;
    for (int r = 0; r < kMaxWorld; ++r) pp.p[r] = r < x->world ? x->peer[r] : nullptr;
    int grid = (int)((n + 255) / 256);
    if (grid > 592) grid = 592;
    const bool sliced = k > 32 && (size_t)3 * next_pow2(k) * 8 <= 200 * 1024;
    const ExchangeView lv = view_of(x->local, x->world, x->cap);
    unsigned int *err = x->d_counter + 1;
    if (epoch > 2u)  // every peer has merged call epoch - 2, the previous user of this parity's buffers (normally long true)
        exchange_wait_kernel<<<1, 32, 0, st>>>(lv.flags + kConsumedWord, x->world, epoch - 2u, err, 2u);
    exchange_push_kernel<<<grid, 256, 0, st>>>(ids_device, scores_device, n, pp, x->rank, x->world, x->cap, parity, epoch, x->d_counter,
                                               x->d_counter + 1, sliced ? n_queries : 0, k);
    BM25CU_CHECK(cudaGetLastError());
    // the merge (and its wait for the slowest shard) runs on the side stream: the caller's stream is free for the next batch
    BM25CU_CHECK(cudaEventRecord(x->ev_pushed, st));
    BM25CU_CHECK(cudaStreamWaitEvent(x->side, x->ev_pushed, 0));
    st = x->side;
    exchange_wait_kernel<<<1, 32, 0, st>>>(lv.flags + parity * kMaxWorld, x->world, epoch, err, 1u);  // every rank's rows are here
    if (k <= 32) {
        exchange_merge_warp_kernel<<<(n_queries + 7) / 8, 256, 0, st>>>(x->local, x->world, x->cap, parity, epoch, n_queries, k,
                                                                       shift_per_query_device, out_ids_device, out_scores_device,
                                                                       x->d_counter + 1, pp, x->rank, x->d_counter + 2);
    } else if (sliced) {
        const int kcap = next_pow2(k);
        const size_t smem = (size_t)3 * kcap * 8;
        const int mine = slice_lo(n_queries, x->world, x->rank + 1) - slice_lo(n_queries, x->world, x->rank);
        BM25CU_CHECK(cudaFuncSetAttribute(exchange_merge_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exchange_merge_slice_kernel<<<mine > 0 ? mine : 1, 256, smem, st>>>(x->local, x->world, x->cap, parity, epoch, n_queries, k, kcap,
                                                                          shift_per_query_device, x->d_counter + 1, pp, x->rank, x->d_counter + 3);
        BM25CU_CHECK(cudaGetLastError());
        exchange_wait_kernel<<<1, 32, 0, st>>>(lv.flags + kGatheredWord + parity * kMaxWorld, x->world, epoch, err, 1u);  // every merged slice
        exchange_copy_out_kernel<<<grid, 256, 0, st>>>(x->local, x->world, x->cap, parity, epoch, n, out_ids_device, out_scores_device,
                                                       x->d_counter + 1, pp, x->rank, x->d_counter + 2);
    } else {
        const int kcap = next_pow2(k);
        const size_t smem = (size_t)3 * kcap * 8;
        BM25CU_REQUIRE(smem <= 200 * 1024, "bm25cu_exchange_merge: k too large for one CTA (max 8192)");
        BM25CU_CHECK(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exchange_merge_kernel<<<n_queries, 256, smem, st>>>(x->local, x->world, x->cap, parity, epoch, n_queries, k, kcap,
                                                           shift_per_query_device, out_ids_device, out_scores_device, x->d_counter + 1,
                                                           pp, x->rank, x->d_counter + 2);
    }
    BM25CU_CHECK(cudaGetLastError());
    BM25CU_CHECK(cudaEventRecord(x->ev_merged, x->side));
    x->merged_pending = true;
    return BM25CU_OK;
}

/* error bits of the merge kernels so far (1: a peer's rows did not arrive within 10 s); synchronises the stream */
int bm25cu_exchange_check(bm25cu_exchange *x, void *stream) {
    BM25CU_REQUIRE(x != nullptr, "bm25cu_exchange_check: null handle");
    BM25CU_CHECK(cudaSetDevice(x->device));
    if (x->merged_pending) BM25CU_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, x->ev_merged, 0));
    unsigned int err = 0;
    BM25CU_CHECK(cudaMemcpyAsync(&err, x->d_counter + 1, sizeof(err), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    BM25CU_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    if (err) {
        set_error(err & 1u ? "bm25cu_exchange: a peer's rows did not arrive within 10 s"
                           : "bm25cu_exchange: a peer did not finish merging the call before last within 10 s");
        return BM25CU_ECUDA;
    }
    return BM25CU_OK;
}

/* make `stream` wait for the merge enqueued last (no host synchronisation): work enqueued on it afterwards sees the rows */
int bm25cu_exchange_join(bm25cu_exchange *x, void *stream) {
    BM25CU_REQUIRE(x != nullptr, "bm25cu_exchange_join: null handle");
    BM25CU_CHECK(cudaSetDevice(x->device));
    if (x->merged_pending) BM25CU_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, x->ev_merged, 0));
    return BM25CU_OK;
}

int bm25cu_exchange_free(bm25cu_exchange *x) {
    if (!x) return BM25CU_OK;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->local);
    cudaFree(x->d_counter);
    if (x->side) cudaStreamDestroy(x->side);
    if (x->ev_pushed) cudaEventDestroy(x->ev_pushed);
    if (x->ev_merged) cudaEventDestroy(x->ev_merged);
    delete x;
    return BM25CU_OK;
}

}  // extern "C"
