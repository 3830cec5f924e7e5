// Block-level selection primitives: order-preserving keys, MSB-first radix select of the k largest,
// deterministic collection (ties by lowest index) and a bitonic sort (key descending, id ascending).
// Used by the dense general kernel, the standalone top-k (selection.topk replacement) and the merge kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bm25cu {

// float32 -> uint32 such that a < b  <=>  key(a) < key(b)   (-0.0 is folded onto +0.0, as numpy compares them equal)
__device__ __forceinline__ uint32_t f32_key(float f) {
    uint32_t u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t f64_key(double f) {
    uint64_t u = (uint64_t)__double_as_longlong(f);
    if (u == 0x8000000000000000ull) u = 0ull;
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

template <typename Key>
struct SelectState {
    Key prefix;
    unsigned long long need;
};

// MSB-first radix select. `load(i)` returns the key of element i (0 <= i < n). On return every thread holds the
// key of the k-th largest element (thr) and how many elements equal to thr belong to the top-k (need >= 1).
// Requires 1 <= k <= n. `hist` is 256 unsigned ints of shared memory, `st` one SelectState in shared memory.
template <typename Key, typename Load>
__device__ void block_radix_select(Load load, long long n, long long k, unsigned int *hist, SelectState<Key> *st,
                                   Key common_key, Key &thr, long long &need_out) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    Key prefix = 0, mask = 0;
    unsigned long long need = (unsigned long long)k;
    for (int pass = (int)sizeof(Key) - 1; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += nt) hist[i] = 0;
        __syncthreads();
        const int shift = 8 * pass;
        unsigned int common_cnt = 0;  // warp-aggregated count of the (very frequent) common key, e.g. score 0
        const bool common_live = ((common_key & mask) == prefix);
        const long long n_round = ((n + nt - 1) / nt) * nt;
        for (long long i = tid; i < n_round; i += nt) {
            bool valid = i < n;
            Key key = valid ? load(i) : 0;
            bool live = valid && ((key & mask) == prefix);
            bool is_common = live && common_live && key == common_key;
            unsigned int b = __ballot_sync(0xffffffffu, is_common);
            if (lane == 0) common_cnt += __popc(b);
            if (live && !is_common) atomicAdd(&hist[(unsigned int)(key >> shift) & 0xFFu], 1u);
        }
        if (lane == 0 && common_cnt) atomicAdd(&hist[(unsigned int)(common_key >> shift) & 0xFFu], common_cnt);
        __syncthreads();
        if (tid == 0) {
            unsigned long long acc = 0;
            int sel = 0;
            for (int b = 255; b >= 0; --b) {
                unsigned long long c = hist[b];
                if (acc + c >= need) { sel = b; break; }
                acc += c;
            }
            st->prefix = prefix | ((Key)sel << shift);
            st->need = need - acc;
        }
        __syncthreads();
        prefix = st->prefix;
        need = st->need;
        mask |= ((Key)0xFF << shift);
        __syncthreads();
    }
    thr = prefix;
    need_out = (long long)need;
}

// Collect the k selected elements: keys > thr (any order) and the `need` lowest-index elements with key == thr.
// out_key/out_id have k slots (shared or global). `warp_cnt` is 32 (+1) unsigned ints of shared memory and
// `gt_counter` one unsigned int of shared memory. Deterministic result set.
template <typename Key, typename Id, typename Load>
__device__ void block_collect(Load load, long long n, long long k, Key thr, long long need, Key *out_key, Id *out_id,
                              unsigned int *warp_cnt, unsigned int *gt_counter) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const long long chunk = (((n + nwarp - 1) / nwarp + 31) / 32) * 32;  // per-warp contiguous chunk
    const long long lo = (long long)warp * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    unsigned int eq = 0;
    for (long long i = lo + lane; i < lo + chunk; i += 32) {
        bool e = (i < hi) && load(i) == thr;
        eq += __popc(__ballot_sync(0xffffffffu, e));
    }
    if (lane == 0) warp_cnt[warp] = eq;
    if (tid == 0) *gt_counter = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned int acc = 0;
        for (int w = 0; w < nwarp; ++w) { unsigned int c = warp_cnt[w]; warp_cnt[w] = acc; acc += c; }
    }
    __syncthreads();
    unsigned long long tie_base = warp_cnt[warp];
    const long long n_gt = k - need;
    for (long long i = lo + lane; i < lo + chunk; i += 32) {
        bool valid = i < hi;
        Key key = valid ? load(i) : 0;
        bool e = valid && key == thr;
        bool g = valid && key > thr;
        unsigned int be = __ballot_sync(0xffffffffu, e);
        if (g) {
            unsigned int slot = atomicAdd(gt_counter, 1u);
            out_key[slot] = key;
            out_id[slot] = (Id)i;
        }
        if (e) {
            unsigned long long rank = tie_base + __popc(be & ((1u << lane) - 1u));
            if ((long long)rank < need) {
                out_key[n_gt + rank] = key;
                out_id[n_gt + rank] = (Id)i;
            }
        }
        tie_base += __popc(be);
    }
    __syncthreads();
}

// Bitonic sort of P (power of two) pairs: key descending, id ascending among equal keys. Arrays may live in
// shared or global memory (one block works on them).
template <typename Key, typename Id>
__device__ void block_bitonic_sort_desc(Key *key, Id *id, int P) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (P >> 1); t += nt) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);  // first half of each bitonic block sorted "descending" overall
                Key ka = key[lo], kb = key[hi];
                Id ia = id[lo], ib = id[hi];
                // a_before_b: a should come first in descending-key / ascending-id order
                bool a_before_b = (ka > kb) || (ka == kb && ia < ib);
                bool swap = desc ? !a_before_b : a_before_b;
                if (swap) {
                    key[lo] = kb; key[hi] = ka;
                    id[lo] = ib; id[hi] = ia;
                }
            }
        }
    }
    __syncthreads();
}

// Block-wide: `add[0..cnt)` (descending, distinct non-zero keys) merged into `cur[0..cap)` (descending, zeros = empty);
// the cap largest go to `nxt` (rank of an element = its index + how many keys of the other list beat it).
__device__ __forceinline__ void rank_merge(const unsigned long long *cur, unsigned long long *nxt, int cap,
                                           const unsigned long long *add, int cnt, int tid, int nt) {
    for (int a = tid; a < cap; a += nt) {
        const unsigned long long key = cur[a];
        int lo = 0, hi = cnt;  // first add[i] <= key
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (add[mid] > key) lo = mid + 1; else hi = mid; }
        if (a + lo < cap) nxt[a + lo] = key;
    }
    for (int b = tid; b < cnt; b += nt) {
        const unsigned long long key = add[b];
        int lo = 0, hi = cap;  // first cur[i] < key
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cur[mid] > key) lo = mid + 1; else hi = mid; }
        if (b + lo < cap) nxt[b + lo] = key;
    }
}

// Block-wide bitonic sort, descending, n a power of two, keys in shared memory
__device__ __forceinline__ void block_sort_desc(unsigned long long *a, int n, int tid, int nt) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n >> 1); t += nt) {
                const int lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
    }
}

// Block-wide: the k largest 64-bit keys of n_rows rows of k keys each (0 = empty slot), e.g. the shards' top-k rows as
// (order-preserving score key << 32 | ~doc id). `load(r, j)` returns key j of row r. Rows are normally sorted descending
// (every producer in this library sorts by score descending, id ascending); a row that is not is sorted here first. Row by
// row: the row is rank-merged into the running top list - two binary searches per key instead of a bitonic sort over
// all n_rows * k keys (8 rows of 1 000: 16 k searches against 370 k compare-exchanges). Shared memory: cur, nxt, row of
// `cap` = next_pow2(k) keys each; the result is returned in the buffer the function returns (cur or nxt).
template <typename Load>
__device__ unsigned long long *block_merge_sorted_rows(int n_rows, int k, int cap, unsigned long long *cur, unsigned long long *nxt,
                                                        unsigned long long *row, Load load) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < cap; e += nt) cur[e] = 0ull;
    for (int r = 0; r < n_rows; ++r) {
        __syncthreads();
        for (int e = tid; e < cap; e += nt) row[e] = e < k ? load(r, e) : 0ull;
        __syncthreads();
        bool unsorted = false;
        for (int e = tid; e + 1 < k; e += nt) unsorted |= row[e] < row[e + 1];
        if (__syncthreads_or(unsorted)) block_sort_desc(row, cap, tid, nt);  // (ends with a barrier)
        int cnt = 0;  // sorted descending, zeros (empty) at the end: first zero by binary search
        { int lo = 0, hi = k; while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] != 0ull) lo = mid + 1; else hi = mid; } cnt = lo; }
        rank_merge(cur, nxt, cap, row, cnt, tid, nt);  // writes every slot of nxt: cap + cnt distinct ranks
        unsigned long long *t = cur; cur = nxt; nxt = t;
    }
    __syncthreads();
    return cur;
}

__host__ __device__ inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace bm25cu
