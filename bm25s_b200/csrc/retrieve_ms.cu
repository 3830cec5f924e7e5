// K1m: list-at-a-time scoring + top-k with exact upper-bound pruning and skipping (sm_100a).
//
// Replaces, per query, BM25._compute_relevance_from_scores (bm25s/__init__.py:312-318), the weight mask and the
// BM25L/BM25+ shift of get_scores_from_ids (:610-618) and selection.topk (bm25s/selection.py:14-64) - the same
// contract as retrieve_fast.cu, reached with far fewer postings touched:
//
//   * the query's distinct tokens are ordered by their upper bound ub = multiplicity * column maximum, largest first
//     (rare tokens first). List i is STREAMED - only its float32 scores, 16 bytes per load - and a posting (d, v) is a
//     survivor when  mult*v + sum of ub over the lists after i  can still reach theta, the k-th best key so far.
//     Documents that contain a token of an earlier list were scored when that list was streamed.
//   * a survivor is looked up in the other lists: one load of the column's presence bitmap, and for a "maybe" ONE 32-byte
//     sector of the column's slot table - the (doc id, score) slots of the half bucket the document falls in (index.cu
//     builds both at upload; columns without a slot table are searched in the slice their bucket table names). The bound
//     is re-checked after every lookup, and a document that passes is summed in QUERY-TOKEN ORDER, duplicates added again
//     at their position - the reference's float32 order, so scores are bit-identical to numpy's.
//   * when the upper bounds of the lists not yet streamed sum below theta, no unseen document can enter the top-k and
//     the query ends: the long lists of frequent tokens are usually never streamed, only probed.
//   * candidates are 64-bit keys (score bits, ~doc): order = score descending, doc ascending, as in retrieve_fast.cu;
//     pruning keeps everything whose bound is >= theta's score, acceptance is key > theta, so the result does not
//     depend on the order in which documents are met. For k <= 32 the top list lives in one warp's registers (CTAs of
//     128 threads, nine per SM), for k <= 1024 in shared memory (CTAs of 256 threads, four per SM).
//   * a 0/1 weight mask resident as a bitmap (bm25cu_index_set_mask) is tested as soon as a survivor's doc id is known:
//     a removed document costs no lookup; other masks in [0, 1] are applied where a candidate is emitted.
//   * a query is cut into doc-range PARTS of bounded cost (ms_plan_*_kernel); CTAs pull parts longest first, the parts of
//     a query publish their k-th best score to each other, ms_merge_kernel unites their top lists and writes the row.
//
// Queries this kernel does not take (empty, > 15 tokens, k > 1024, fewer than k documents with a positive score, masks
// outside [0, 1], a part over its survivor budget) are appended to `fb_list` and served by retrieve_fast_kernel in the
// same call.
#include "fast_common.cuh"

namespace bm25cu {

struct MsParams {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    const float *colmax;
    const float *colkth;        // start value table valid for this k (nullptr: none)
    const unsigned int *btab;   // bucket tables (index.cu) or nullptr: binary search
    const long long *btab_off;  // per column: offset into btab, -1 = no table
    const unsigned char *btab_sh;
    const unsigned int *pbm;    // presence bitmaps (index.cu) or nullptr
    const long long *pbm_off;
    const unsigned char *pbm_sh;
    long long nnz, btab_len, pbm_len, stab_len, partial_rows;  // array extents, read by the bounds checks of `make checked` only
    const uint4 *stab;          // slot tables (index.cu) or nullptr: a bucket is four uint4 = two sectors
    const long long *stab_off;
    const unsigned char *stab_sh;
    int32_t n_docs, n_vocab, doc_base;
    RetrieveArgs a;
    Ctrl *ctrl;
    unsigned int *fb_list;
    int route_all;
    unsigned int budget;  // survivors a work item may look up before its query is handed on (safety valve)
    // work items: a query is cut into P doc-range parts served by different CTAs (MsPlan below)
    const unsigned long long *items;  // q | part << 32 | P << 40, longest first
    const unsigned int *part_base;    // first row of a query's parts in `partial`
    unsigned int *gtheta;             // per query: score bits every part may prune with (max of the parts' k-th best)
    unsigned int *qstat;              // per query: 1 = handed on
    unsigned long long *partial;      // [row][k] sorted keys of a part's top list
    int kcap;                         // next_pow2(k), at least 32; > 32: the top list lives in shared memory
    unsigned int big_merge_at;        // k > 32: pending keys that start a merge in the middle of a probe call (<= 512)
    float *theta_io;                  // experiment (BM25CU_DEBUG_THETA): 1 = the merge kernel records every query's final
    int theta_mode;                   // threshold, 2 = every part starts from the recorded value (what a perfect start buys)
};

// `make checked` (-DBM25CU_CHECKED; compute-sanitizer is not available on the GPU pool): every index this kernel computes
// into the posting arrays, the lookup structures, the shared-memory queues and the partial-result rows is compared with
// the array's extent; a violation sets error bit 3 ("bounds check failed") and the access is skipped.
#ifdef BM25CU_CHECKED
#define MS_OK(cond) ((cond) ? true : (atomicOr(&p.ctrl->error, 8u), false))
#else
#define MS_OK(cond) true
#endif

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ unsigned long long shfl64_xor(unsigned long long v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// 32 keys, one per lane, sorted descending by lane (bitonic network on shuffles)
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long x, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long o = shfl64_xor(x, stride);
            const bool lower = (lane & stride) == 0;
            const bool desc = (lane & size) == 0;
            x = (lower == desc) ? umax64(x, o) : umin64(x, o);
        }
    }
    return x;
}

__device__ __forceinline__ bool reaches(float acc, float rest, float th) {
    return __fmul_ru(__fadd_ru(acc, rest), 1.00002f) >= th;
}

// Slot-table lookup of document d (index.cu, build_slot_tables): the 32-byte sector of d's half of its 64-byte bucket;
// flagged sectors lent postings to the other sector of the line. 1: found (val), 0: absent, -1: the bucket holds more
// than its slots and d is not among them - search the column. Two sector formats, one per index (STAB_COMPACT):
//   wide    four (doc id, score) pairs;
//   compact five 16-bit doc offsets inside the bucket + a meta field (words 0..2), then the five scores (words 3..7).
__device__ __forceinline__ bool sector_match(const uint4 a, const uint4 b, unsigned off, unsigned meta, float &val) {
    if ((meta & 1u) && (a.x & 0xffffu) == off) { val = __uint_as_float(a.w); return true; }
    if ((meta & 2u) && (a.x >> 16) == off) { val = __uint_as_float(b.x); return true; }
    if ((meta & 4u) && (a.y & 0xffffu) == off) { val = __uint_as_float(b.y); return true; }
    if ((meta & 8u) && (a.y >> 16) == off) { val = __uint_as_float(b.z); return true; }
    if ((meta & 16u) && (a.z & 0xffffu) == off) { val = __uint_as_float(b.w); return true; }
    return false;
}

template <bool CP>
__device__ __forceinline__ int slot_lookup(const uint4 *__restrict__ stab, long long first_bucket, int sh, int d, float &val) {
    const uint4 *bk = stab + ((first_bucket + (long long)((unsigned)d >> sh)) << 2);
    const unsigned h = ((unsigned)d >> (sh - 1)) & 1u;
    const uint4 a = __ldg(bk + 2 * h), b = __ldg(bk + 2 * h + 1);
    if (CP) {
        const unsigned off = (unsigned)d & ((1u << sh) - 1u);
        const unsigned meta = ~(a.z >> 16);  // stored inverted: an untouched sector (all ones) is empty
        if (sector_match(a, b, off, meta, val)) return 1;
        if (!(meta & 32u)) return 0;
        const uint4 c = __ldg(bk + 2 * (1u - h)), e = __ldg(bk + 2 * (1u - h) + 1);  // the other sector of the same 64-byte line
        if (sector_match(c, e, off, ~(c.z >> 16), val)) return 1;
        return (meta & 64u) ? -1 : 0;
    }
    if ((a.x & 0x7fffffffu) == (unsigned)d) { val = __uint_as_float(a.y); return 1; }
    if ((a.z & 0x7fffffffu) == (unsigned)d) { val = __uint_as_float(a.w); return 1; }
    if (b.x == (unsigned)d) { val = __uint_as_float(b.y); return 1; }
    if (b.z == (unsigned)d) { val = __uint_as_float(b.w); return 1; }
    if (!((a.x >> 31) && a.x != 0xffffffffu)) return 0;
    const uint4 c = __ldg(bk + 2 * (1u - h)), e = __ldg(bk + 2 * (1u - h) + 1);  // the other sector of the same 64-byte line
    if ((c.x & 0x7fffffffu) == (unsigned)d) { val = __uint_as_float(c.y); return 1; }
    if ((c.z & 0x7fffffffu) == (unsigned)d) { val = __uint_as_float(c.w); return 1; }
    if (e.x == (unsigned)d) { val = __uint_as_float(e.y); return 1; }
    if (e.z == (unsigned)d) { val = __uint_as_float(e.w); return 1; }
    return ((a.z >> 31) && a.z != 0xffffffffu) ? -1 : 0;
}

template <int NT, int UNROLL, int R, bool BIG, bool MB, bool CP>  // MB: a 0/1 weight mask as a bitmap; CP: compact slot tables (own instantiations: registers)
struct Ms {
    static constexpr int CHUNK = NT * 4 * UNROLL;
    static constexpr int QCAP = CHUNK + NT * R;
    struct Sh {
        long long beg[16];
        unsigned long long top[32];
        unsigned long long pend[BIG ? (NT * R + 512 <= 1024 ? 1024 : 2048) : NT * R + 32];  // BIG: sorted in place, padded to a power of two (a round adds at most NT * R keys to at most 512 waiting ones)
        unsigned long long theta_key;
        int df[16];
        float ub[16];
        float mult[16];
        float suf[17];
        int bt_sh[16];
        long long bt_off[16];
        long long pb_off[16];
        int pb_sh[16];
        long long st_off[16];
        int st_sh[16];
        unsigned char pos2u[16];
        unsigned int pend_cnt, q_cnt, q_slot, work, gth, abort, top_sel;
        int U, T, route;
        int lo[16], hi[16];   // the part's slice of every list
        unsigned long long tot;
        unsigned int queue[QCAP];
    };
    const MsParams &p;
    Sh &s;
    unsigned long long *big;  // k > 32: two top lists of kcap keys (dynamic shared memory), s.top_sel names the current one
    int kcap;
#ifdef BM25CU_PROFILE
    // 0 postings streamed, 1 survivors queued, 2 survivors probed, 3 lookups, 4 lookup steps, 5 candidates accepted,
    // 6 probe rounds, 7 merges, 8 lists streamed, 9 lists skipped, 10 queries served, 11 queries handed on
    unsigned long long cnt[16] = {0};  // 12 survivors of dense lists (df >= n_docs/64), 13 lookups that passed a presence bitmap, 14 steps in dense lists, 15 found
#define MS_CNT(i, v) (cnt[i] += (unsigned long long)(v))
#else
#define MS_CNT(i, v)
#endif
    __device__ Ms(const MsParams &p_, Sh &s_, unsigned long long *big_) : p(p_), s(s_), big(big_), kcap(p_.kcap) {}

    // is document d in list j? (val = its score)
    __device__ __forceinline__ bool lookup(int j, int d, float &val) {
        MS_CNT(3, 1);
        const long long po = s.pb_off[j];
        if (po >= 0) {
            const unsigned x = (unsigned)d >> s.pb_sh[j];
            if (!MS_OK(po + (x >> 5) < p.pbm_len)) return false;
            if (!((__ldg(p.pbm + po + (x >> 5)) >> (x & 31u)) & 1u)) return false;
            MS_CNT(13, 1);
        }
        return search(j, d, val);
    }

    // k > 32: all threads. The pending keys are sorted (bitonic, shared memory) and rank-merged into the top list.
    __device__ void merge_pending_big(int tid, int k, unsigned int *gq) {
        const int cnt = (int)s.pend_cnt;
        const int P2 = next_pow2(cnt > 1 ? cnt : 1);
        for (int i = cnt + tid; i < P2; i += NT) s.pend[i] = 0ull;
        __syncthreads();
        block_sort_desc(s.pend, P2, tid, NT);
        unsigned long long *cur = big + (s.top_sel ? kcap : 0), *nxt = big + (s.top_sel ? 0 : kcap);
        rank_merge(cur, nxt, kcap, s.pend, cnt, tid, NT);
        __syncthreads();
        if (tid == 0) {
            s.top_sel ^= 1u;
            const unsigned long long kth = nxt[k - 1];
            if (kth > s.theta_key) {
                s.theta_key = kth;
                if (gq) atomicMax(gq, (unsigned int)(kth >> 32));
            }
            s.pend_cnt = 0;
        }
    }

    __device__ void merge_pending(int lane, int k, unsigned int *gq) {
        const unsigned cnt = s.pend_cnt;
        unsigned long long t = s.top[lane];
        for (unsigned base = 0; base < cnt; base += 32) {
            unsigned long long b = (base + lane < cnt) ? s.pend[base + lane] : 0ull;
            if (cnt - base > 1) b = warp_sort_desc(b, lane);
            t = umax64(t, shfl64(b, 31 - lane));  // bitonic: the 32 largest of the union
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) {
                const unsigned long long o = shfl64_xor(t, stride);
                t = ((lane & stride) == 0) ? umax64(t, o) : umin64(t, o);
            }
        }
        s.top[lane] = t;
        const unsigned long long kth = shfl64(t, k - 1);
        if (lane == 0) {
            if (kth > s.theta_key) {
                s.theta_key = kth;
                if (gq) atomicMax(gq, (unsigned int)(kth >> 32));  // k documents of this part score at least this much
            }
            s.pend_cnt = 0;
        }
    }

    // position search of document d in list j (after its presence bitmap said "maybe")
    __device__ __forceinline__ bool search(int j, int d, float &val) {
        const long long so = s.st_off[j];
        if (so >= 0) {
            if (!MS_OK(so + (long long)((unsigned)d >> s.st_sh[j]) < p.stab_len)) return false;
            const int r = slot_lookup<CP>(p.stab, so, s.st_sh[j], d, val);
            if (r >= 0) { if (r) MS_CNT(15, 1); return r != 0; }
        }
        const long long beg = s.beg[j];
        int lo = 0, hi = s.df[j];
        const long long bo = s.bt_off[j];
        if (bo >= 0) {
            const unsigned b = (unsigned)d >> s.bt_sh[j];
            if (!MS_OK(bo + b + 1 < p.btab_len)) return false;
            lo = (int)__ldg(p.btab + bo + b);
            hi = (int)__ldg(p.btab + bo + b + 1);
            if (!MS_OK(lo >= 0 && hi <= s.df[j])) return false;
        }
        const int32_t *col = p.indices + beg;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int x = __ldg(col + mid);
            MS_CNT(4, 1);
            if (x < d) lo = mid + 1;
            else if (x > d) hi = mid;
            else { val = __ldg(p.data + beg + mid); MS_CNT(15, 1); return true; }
        }
        return false;
    }

    // A survivor that passed every bound: drop it if an earlier list holds it (scored there), else sum it in the
    // reference's order - token by token, duplicates again (bm25s/__init__.py:316-318) - and offer it to the top list.
    __device__ void finish(int i, int d, float v, unsigned found, const float *vals, unsigned long long tk, float th) {
        float dummy;
        for (int j = i - 1; j >= 0; --j)
            if (lookup(j, d, dummy)) return;
        float sc = 0.0f;
        const int T = s.T;
        for (int t = 0; t < T; ++t) {
            const unsigned u = s.pos2u[t];
            if (u < 16u && ((found >> u) & 1u)) sc = __fadd_rn(sc, (int)u == i ? v : vals[u]);
        }
        if (p.a.mask) sc = (float)((double)sc * p.a.mask[d]);  // (a bitmap mask dropped its documents at the start of the round)
        const unsigned long long key = pack_cand(sc, d);
        if (key > tk && sc >= th) {
            const unsigned at = atomicAdd(&s.pend_cnt, 1u);
            if (MS_OK(at < sizeof(s.pend) / sizeof(s.pend[0]))) s.pend[at] = key;
            MS_CNT(5, 1);
        }
    }

    // A round: every thread takes R survivors; their (doc, score) loads and the presence test of the next list are
    // issued together (the usual survivor is absent there and ends on that test).
    // Pruning threshold: this part's k-th best score or what the other parts of the query published (s.gth is refreshed
    // from global memory by thread 0 only, in front of a barrier, so every decision on it is uniform over the CTA).
    __device__ __forceinline__ float threshold() const {
        return fmaxf(cand_score(s.theta_key), __uint_as_float(s.gth));
    }
    __device__ __forceinline__ void refresh(unsigned int *gq) {
        if (gq) { const unsigned g = *(volatile unsigned int *)gq; if (g > s.gth) s.gth = g; }
    }

    __device__ void probe_rounds(int i, unsigned n, int tid, int k, unsigned int *gq) {
        const long long beg = s.beg[i];
        const float m = s.mult[i], rest = s.suf[i + 1];
        const int U = s.U;
        for (unsigned base = 0; base < n; base += NT * R) {
            const unsigned long long tk = s.theta_key;
            const float th = threshold();
            if (tid == 0) MS_CNT(6, 1);
            int d[R];
            float v[R], acc[R];
            unsigned found[R];
            bool alive[R];
            float vals[R * 16];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned idx = base + r * NT + tid;
                long long at = beg + ((idx < n) ? s.queue[idx] : 0u);
                if (!MS_OK(idx >= n || (idx < (unsigned)QCAP && at >= 0 && at < p.nnz))) at = beg;
                d[r] = __ldg(p.indices + at);
                v[r] = __ldg(p.data + at);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = __fmul_ru(m, v[r]);
                alive[r] = base + r * NT + tid < n && reaches(acc[r], rest, th);
                found[r] = 1u << i;
                if (alive[r]) MS_CNT(2, 1);
            }
            if (MB) {  // 0/1 weight mask: a document it removes scores 0 and ends here, before any lookup
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (alive[r]) alive[r] = (__ldg(p.a.mask_bits + ((unsigned)d[r] >> 5)) >> ((unsigned)d[r] & 31u)) & 1u;
            }
            // The other lists, largest upper bound first. Per list: the presence tests of the R survivors go out together,
            // the (few) position searches run with the lanes that need one side by side, then the bounds are re-checked.
            for (int j = i + 1; j < U; ++j) {
                const long long po = s.pb_off[j];
                const int psh = s.pb_sh[j];
                unsigned w[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned x = (unsigned)d[r] >> psh;
                    w[r] = (po >= 0 && alive[r] && MS_OK(po + (x >> 5) < p.pbm_len)) ? __ldg(p.pbm + po + (x >> 5)) : 0xffffffffu;
                }
                unsigned todo = 0u;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned x = (unsigned)d[r] >> psh;
                    if (alive[r]) { MS_CNT(3, 1); if ((w[r] >> (x & 31u)) & 1u) todo |= 1u << r; }
                }
                const float mj = s.mult[j];
                while (todo) {
                    const int r = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    int dd = d[0];
#pragma unroll
                    for (int rr = 1; rr < R; ++rr) dd = (rr == r) ? d[rr] : dd;
                    float val;
                    MS_CNT(13, 1);
                    if (search(j, dd, val)) {
                        const float add = __fmul_ru(mj, val);
                        vals[r * 16 + j] = val;
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            if (rr == r) { acc[rr] = __fadd_ru(acc[rr], add); found[rr] |= 1u << j; }
                    }
                }
                const float sufn = s.suf[j + 1];
                bool any = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    alive[r] = alive[r] && reaches(acc[r], sufn, th);
                    any |= alive[r];
                }
                if (!__any_sync(0xffffffffu, any)) break;
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (alive[r]) finish(i, d[r], v[r], found[r], vals + r * 16, tk, th);
            if (tid == 0) refresh(gq);
            // the decision to merge is taken by the barrier itself (a thread that pushed sees its own push): a plain read
            // behind the barrier could race with the reset at the end of the merge
            const unsigned pc = s.pend_cnt;
            const bool want = BIG ? (pc >= p.big_merge_at || (pc > 0u && base + NT * R >= n))  // the shared-memory merge is dearer: batch it
                                  : pc > 0u;
            if (__syncthreads_or(want)) {
                if (tid == 0) MS_CNT(7, 1);
                if (BIG) merge_pending_big(tid, k, gq);
                else if (tid < 32) merge_pending(tid, k, gq);
                __syncthreads();
            }
        }
        if (tid == 0) { s.q_cnt = 0; s.work += n; }
        __syncthreads();
    }

    // (one shared-memory atomic per surviving posting: taking one slot range per warp with a ballot + prefix count instead
    // was measured slower, 2.21 -> 2.32 ms per batch)
    __device__ __forceinline__ void check(float v, int pos, int df, int lo, float m, float rest, float th) {
        if (__fmul_ru(__fmaf_ru(m, v, rest), 1.00002f) >= th && (unsigned)pos < (unsigned)df)
        {
            const unsigned at = atomicAdd(&s.q_cnt, 1u);
            if (MS_OK(at < (unsigned)QCAP)) s.queue[at] = (unsigned)(pos + lo);
            MS_CNT(1, 1);
        }
    }

    __device__ bool stream(int i, int tid, int k, unsigned int *gq) {
        const int lo = s.lo[i];
        const long long beg = s.beg[i] + lo;  // positions are pushed relative to the column start
        const int df = s.hi[i] - lo;
        const float m = s.mult[i], rest = s.suf[i + 1];
        const int head = (int)(beg & 3);
        const int nvec = (df + head + 3) >> 2;
        const float4 *vp = reinterpret_cast<const float4 *>(p.data + (beg - head));
        if (tid == 0) { MS_CNT(0, df); MS_CNT(8, 1); }
        for (int v0 = 0; v0 < nvec; v0 += NT * UNROLL) {
            const float th = threshold();
            float4 x[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int v = v0 + u * NT + tid;
                x[u] = (v < nvec) ? __ldcs(vp + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int v = v0 + u * NT + tid;
                const int pos = 4 * v - head;
                if (v < nvec) {
                    check(x[u].x, pos, df, lo, m, rest, th);
                    check(x[u].y, pos + 1, df, lo, m, rest, th);
                    check(x[u].z, pos + 2, df, lo, m, rest, th);
                    check(x[u].w, pos + 3, df, lo, m, rest, th);
                }
            }
            if (tid == 0) refresh(gq);
            __syncthreads();
            const unsigned n = s.q_cnt;
            __syncthreads();  // nobody pushes the next chunk's survivors before everybody read the count
            const bool last = v0 + NT * UNROLL >= nvec;
            if (n >= (unsigned)(NT * R) || (last && n > 0)) {
                probe_rounds(i, n, tid, k, gq);
                if (s.work > p.budget) return false;  // too many survivors: the streaming kernel does this query
            }
        }
        return true;
    }

    // warp 0: tokens -> distinct lists ordered by upper bound, suffix sums, start threshold
    __device__ void prologue(int q, int part, int P, int lane) {
        const long long qs = p.a.q_ptr[q], qe = p.a.q_ptr[q + 1];
        const long long Tl = qe - qs;
        bool route = p.route_all || Tl > kMaxTok || Tl == 0 || (p.a.mask && p.ctrl->mask_bad);
        if (!route) {
            const int T = (int)Tl;
            const int tok = (lane < T) ? p.a.q_tokens[qs + lane] : -1 - lane;
            long long beg = 0;
            int df = 0;
            float mx = 0.0f, kth = 0.0f;
            if (lane < T) {
                if ((unsigned)tok >= (unsigned)p.n_vocab) atomicOr(&p.ctrl->error, 1u);
                else {
                    beg = p.indptr[tok];
                    df = (int)(p.indptr[tok + 1] - beg);
                    mx = p.colmax[tok];
                    if (p.colkth) kth = p.colkth[tok];
                }
            }
            int first = lane, mult = 0;
            for (int j = 0; j < T; ++j) {
                const int tj = __shfl_sync(0xffffffffu, tok, j);
                if (tj == tok) { ++mult; if (j < first) first = j; }
            }
            const int uniq = (lane < T && first == lane && df > 0) ? 1 : 0;
            const float ub = uniq ? __fmul_ru((float)mult, mx) : 0.0f;
            int rank = 0;
            for (int j = 0; j < T; ++j) {
                const float ubj = __shfl_sync(0xffffffffu, ub, j);
                const int uj = __shfl_sync(0xffffffffu, uniq, j);
                if (uj && j != lane && (ubj > ub || (ubj == ub && j < lane))) ++rank;
            }
            const int U = __popc(__ballot_sync(0xffffffffu, uniq));
            if (uniq) {
                s.beg[rank] = beg;
                s.df[rank] = df;
                s.ub[rank] = ub;
                s.mult[rank] = (float)mult;
                long long bo = -1;
                int bsh = 0;
                if (p.btab) { bo = p.btab_off[tok]; bsh = p.btab_sh[tok]; }
                s.bt_off[rank] = bo;
                s.bt_sh[rank] = bsh;
                long long po = -1;
                int psh = 0;
                if (p.pbm) { po = p.pbm_off[tok]; psh = p.pbm_sh[tok]; }
                s.pb_off[rank] = po;
                s.pb_sh[rank] = psh;
                long long so = -1;
                int ssh = 0;
                if (p.stab) { so = p.stab_off[tok]; ssh = p.stab_sh[tok]; }
                s.st_off[rank] = so;
                s.st_sh[rank] = ssh;
            }
            const int myrank = uniq ? rank : 255;
            const int rf = __shfl_sync(0xffffffffu, myrank, first & 31);
            if (lane < T) s.pos2u[lane] = (unsigned char)rf;
            unsigned long long tot = (lane < T) ? (unsigned long long)df : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const unsigned kb = __reduce_max_sync(0xffffffffu, __float_as_uint(kth));
            __syncwarp();
            if (lane == 0) {
                float acc = 0.0f;
                s.suf[U] = 0.0f;
                for (int j = U - 1; j >= 0; --j) { acc = __fadd_ru(acc, s.ub[j]); s.suf[j] = acc; }
                // the k-th largest posting of one token is reached by k distinct documents and a document scores at
                // least each of its postings: the float below it is a safe start value (as in retrieve_fast.cu)
                s.theta_key = 0ull;
                s.gth = kb > 0u ? kb - 1u : 0u;
                if (p.theta_mode == 2) { const unsigned tb = __float_as_uint(p.theta_io[q]); if (tb > s.gth) s.gth = tb; }
                s.U = U;
                s.T = T;
                s.tot = tot;
                s.pend_cnt = 0;
                s.q_cnt = 0;
                s.work = 0;
                s.top_sel = 0;
            }
            s.top[lane] = 0ull;
            if (BIG) for (int e = lane; e < kcap; e += 32) big[e] = 0ull;
            __syncwarp();
            if (tot == 0) route = true;
            // this part's slice of every list: documents [part * n_docs / P, (part + 1) * n_docs / P)
            if (!route && lane < U) {
                int lo = 0, hi = s.df[lane];
                if (P > 1) {
                    const int d0 = (int)((long long)p.n_docs * part / P), d1 = (int)((long long)p.n_docs * (part + 1) / P);
                    if (part > 0) lo = lower_bound(lane, d0);
                    if (part + 1 < P) hi = lower_bound(lane, d1);
                }
                s.lo[lane] = lo;
                s.hi[lane] = hi;
            }
        }
        if (lane == 0) s.route = route ? 1 : 0;
    }

    // first position of list j whose document is >= d
    __device__ int lower_bound(int j, int d) const {
        int lo = 0, hi = s.df[j];
        const long long bo = s.bt_off[j];
        if (bo >= 0) {
            const unsigned b = (unsigned)d >> s.bt_sh[j];
            lo = (int)__ldg(p.btab + bo + b);
            hi = (int)__ldg(p.btab + bo + b + 1);
        }
        const int32_t *col = p.indices + s.beg[j];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(col + mid) < d) lo = mid + 1; else hi = mid;
        }
        return lo;
    }

    __device__ void run() {
        const int tid = threadIdx.x;
        const int k = p.a.k;
        const unsigned n_items = p.ctrl->ms_items;
        for (;;) {
            if (tid == 0) s.q_slot = atomicAdd(&p.ctrl->ms_next, 1u);
            __syncthreads();
            const unsigned slot = s.q_slot;
            if (slot >= n_items) break;
            const unsigned long long item = p.items[slot];
            const int q = (int)(unsigned)(item & 0xffffffffull), part = (int)((item >> 32) & 0xffu), P = (int)((item >> 40) & 0xffu);
            unsigned int *gq = (P > 1) ? p.gtheta + q : nullptr;
            if (tid < 32) prologue(q, part, P, tid);
            __syncthreads();
            if (!s.route) {
                if (gq && tid == 0 && s.gth) atomicMax(gq, s.gth);
                const int U = s.U;
                for (int i = 0; i < U; ++i) {
                    if (P > 1) {
                        if (tid == 0) { refresh(gq); s.abort = *(volatile unsigned int *)(p.qstat + q); }
                        __syncthreads();
                        if (s.abort) break;  // another part gave the query up
                    }
                    const float th = threshold();
                    if (__fmul_ru(s.suf[i], 1.00002f) < th) { if (tid == 0) MS_CNT(9, U - i); break; }  // no unseen document can reach theta
                    if (s.hi[i] > s.lo[i] && !stream(i, tid, k, gq)) break;
                }
                if (s.work > p.budget) {
                    if (tid == 0) s.route = 1;
                } else if (BIG) {
                    const unsigned long long *cur = big + (s.top_sel ? kcap : 0);
                    if (MS_OK((long long)p.part_base[q] + part < p.partial_rows))
                        for (int t = tid; t < k; t += NT) p.partial[((size_t)p.part_base[q] + part) * k + t] = cur[t];
                } else if (tid < 32) {
                    // the part's top list (sorted, zeros = empty); merged, checked and written by ms_merge_kernel
                    if (tid < k && MS_OK((long long)p.part_base[q] + part < p.partial_rows)) p.partial[((size_t)p.part_base[q] + part) * k + tid] = s.top[tid];
                }
                __syncthreads();
            }
            if (tid == 0) MS_CNT(s.route ? 11 : 10, 1);
            if (s.route && tid == 0) p.qstat[q] = 1u;
            __syncthreads();
        }
#ifdef BM25CU_PROFILE
        for (int i = 0; i < 16; ++i)
            if (cnt[i]) atomicAdd(&p.ctrl->prof[32 + i], cnt[i]);
#endif
    }
};

// ---- plan: cut every query into doc-range parts of bounded cost and order the parts longest first ------------------
// cost of a query = its postings without the longest list (the list that is usually only probed). Two grid-wide passes:
// costs + histogram of log2(cost per part), then the scatter into the bucket-ordered item list.
constexpr int kMsMaxParts = 16;
__global__ void __launch_bounds__(256) ms_plan_cost_kernel(const int32_t *__restrict__ q_tokens, const int64_t *__restrict__ q_ptr,
                                                           int32_t n_queries, const int64_t *__restrict__ indptr, int32_t n_vocab,
                                                           long long part_cost, int max_parts, unsigned int *part_base,
                                                           unsigned long long *qtot, unsigned int *gtheta, unsigned int *qstat,
                                                           unsigned char *scratch /* 2 bytes per query */, Ctrl *ctrl) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    unsigned long long tot = 0, mx = 0;
    const long long T = q_ptr[q + 1] - q_ptr[q];
    for (int64_t i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
        const int tok = q_tokens[i];
        if ((unsigned)tok < (unsigned)n_vocab) {
            const unsigned long long df = (unsigned long long)(indptr[tok + 1] - indptr[tok]);
            tot += df;
            mx = df > mx ? df : mx;
        } else {
            atomicOr(&ctrl->error, 1u);  // every token of the batch is range-checked here, whatever route its query takes
        }
    }
    const unsigned long long cost = (T > 1 && tot > mx) ? tot - mx : tot;
    int P = 1;
    if (T >= 1 && T <= kMaxTok && cost > (unsigned long long)part_cost) {
        const unsigned long long want = (cost + part_cost - 1) / part_cost;
        P = want > (unsigned long long)max_parts ? max_parts : (int)want;
    }
    const unsigned long long per = cost / P;
    const int bk = per ? 64 - __clzll((long long)per) : 0;
    scratch[q] = (unsigned char)bk;
    scratch[n_queries + q] = (unsigned char)P;
    qtot[q] = tot;
    gtheta[q] = 0u;
    qstat[q] = 0u;
    atomicAdd(&ctrl->ms_hist[bk], (unsigned)P);
    part_base[q] = atomicAdd(&ctrl->ms_rows, (unsigned)P);  // rows of the query's parts in `partial` (any order)
}

__global__ void __launch_bounds__(256) ms_plan_scatter_kernel(int32_t n_queries, const unsigned char *__restrict__ scratch,
                                                              unsigned long long *items, Ctrl *ctrl) {
    __shared__ unsigned int base[64];
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int bk = 63; bk >= 0; --bk) { base[bk] = acc; acc += ctrl->ms_hist[bk]; }
        if (blockIdx.x == 0) ctrl->ms_items = acc;
    }
    __syncthreads();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const unsigned bk = scratch[q], P = scratch[n_queries + q];
    const unsigned at = base[bk] + atomicAdd(&ctrl->ms_cursor[bk], P);
    for (unsigned j = 0; j < P; ++j)
        items[at + j] = (unsigned long long)(unsigned)q | ((unsigned long long)j << 32) | ((unsigned long long)P << 40);
}

// ---- merge: one warp per query unites the parts' top lists, checks the result and writes the row -------------------
__global__ void __launch_bounds__(256) ms_merge_kernel(const unsigned char *__restrict__ nparts, const unsigned int *__restrict__ part_base,
                                                       const unsigned long long *__restrict__ partial, const unsigned int *__restrict__ qstat,
                                                       const unsigned long long *__restrict__ qtot, RetrieveArgs a, int32_t doc_base,
                                                       Ctrl *ctrl, unsigned int *fb_list, float *theta_out) {
    const int q = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= a.n_queries) return;
    const int k = a.k;
    bool route = qstat[q] != 0u;
    unsigned long long t = 0ull;
    if (!route) {
        const unsigned r0 = part_base[q], r1 = r0 + nparts[q];
        for (unsigned r = r0; r < r1; ++r) {
            const unsigned long long b = (lane < k) ? partial[(size_t)r * k + lane] : 0ull;  // sorted descending
            t = umax64(t, shfl64(b, 31 - lane));
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) {
                const unsigned long long o = shfl64_xor(t, stride);
                t = ((lane & stride) == 0) ? umax64(t, o) : umin64(t, o);
            }
        }
        const unsigned long long kth = shfl64(t, k - 1);
        route = !(kth != 0ull && cand_score(kth) > 0.0f);  // fewer than k positive matches: padding rules live in retrieve_fast.cu
        if (theta_out && lane == 0) { const unsigned kb = (unsigned)(kth >> 32); theta_out[q] = (!route && kb > 0u) ? __uint_as_float(kb - 1u) : 0.0f; }
    }
    if (route) {
        if (lane == 0) fb_list[atomicAdd(&ctrl->fb_count, 1u)] = (unsigned)q;
        return;
    }
    if (lane < k) {
        float sc = cand_score(t);
        if (a.shift != nullptr) sc = __fadd_rn(sc, a.shift[q]);
        a.out_ids[(size_t)q * k + lane] = cand_doc(t) + doc_base;
        a.out_scores[(size_t)q * k + lane] = sc;
    }
    if (lane == 0) atomicAdd(&ctrl->posting_count, qtot[q]);
}

template <int NT, int UNROLL, int R, int MINB, bool BIG, bool MB, bool CP>
__global__ void __launch_bounds__(NT, MINB) retrieve_ms_kernel(const __grid_constant__ MsParams p) {
    __shared__ typename Ms<NT, UNROLL, R, BIG, MB, CP>::Sh sh;
    extern __shared__ __align__(16) unsigned long long ms_dyn[];
    Ms<NT, UNROLL, R, BIG, MB, CP> m(p, sh, ms_dyn);
    m.run();
}

template <int NT, int MINB, bool BIG>
static cudaError_t launch_ms_variant(const MsParams &p, bool mb, bool cp, int grid, size_t dyn, cudaStream_t st) {
#define BM25CU_MS_LAUNCH(MBV, CPV)                                                                                          \
    do {                                                                                                                    \
        if (dyn > 0) {                                                                                                      \
            cudaError_t e_ = cudaFuncSetAttribute(retrieve_ms_kernel<NT, 4, 4, MINB, BIG, MBV, CPV>,                        \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                   \
            if (e_ != cudaSuccess) return e_;                                                                               \
        }                                                                                                                   \
        retrieve_ms_kernel<NT, 4, 4, MINB, BIG, MBV, CPV><<<grid, NT, dyn, st>>>(p);                                        \
    } while (0)
    if (mb && cp) BM25CU_MS_LAUNCH(true, true);
    else if (mb) BM25CU_MS_LAUNCH(true, false);
    else if (cp) BM25CU_MS_LAUNCH(false, true);
    else BM25CU_MS_LAUNCH(false, false);
#undef BM25CU_MS_LAUNCH
    return cudaGetLastError();
}

// k > 32: one CTA per query, the parts' rows are rank-merged one after the other in shared memory
__global__ void __launch_bounds__(128) ms_merge_big_kernel(const unsigned char *__restrict__ nparts, const unsigned int *__restrict__ part_base, const unsigned long long *__restrict__ partial,
                                                           const unsigned int *__restrict__ qstat, const unsigned long long *__restrict__ qtot,
                                                           RetrieveArgs a, int32_t doc_base, int kcap, Ctrl *ctrl, unsigned int *fb_list) {
    extern __shared__ __align__(16) unsigned long long mg[];  // [2][kcap] top lists + [kcap] one row
    const int q = blockIdx.x, tid = threadIdx.x, k = a.k;
    bool route = qstat[q] != 0u;
    unsigned long long *cur = mg, *nxt = mg + kcap, *row = mg + 2 * kcap;
    if (!route) {
        for (int e = tid; e < kcap; e += 128) cur[e] = 0ull;
        const unsigned r0 = part_base[q], r1 = r0 + nparts[q];
        for (unsigned r = r0; r < r1; ++r) {
            __syncthreads();
            for (int e = tid; e < k; e += 128) row[e] = partial[(size_t)r * k + e];
            __syncthreads();
            int cnt = 0;  // rows are sorted, zeros (empty) at the end: first zero by binary search
            { int lo = 0, hi = k; while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] != 0ull) lo = mid + 1; else hi = mid; } cnt = lo; }
            rank_merge(cur, nxt, kcap, row, cnt, tid, 128);
            unsigned long long *t = cur; cur = nxt; nxt = t;
        }
        __syncthreads();
        const unsigned long long kth = cur[k - 1];
        route = !(kth != 0ull && cand_score(kth) > 0.0f);
    }
    if (route) {
        if (tid == 0) fb_list[atomicAdd(&ctrl->fb_count, 1u)] = (unsigned)q;
        return;
    }
    for (int e = tid; e < k; e += 128) {
        const unsigned long long t = cur[e];
        float sc = cand_score(t);
        if (a.shift != nullptr) sc = __fadd_rn(sc, a.shift[q]);
        a.out_ids[(size_t)q * k + e] = cand_doc(t) + doc_base;
        a.out_scores[(size_t)q * k + e] = sc;
    }
    if (tid == 0) atomicAdd(&ctrl->posting_count, qtot[q]);
}

// The batch-wide conditions under which every query would be handed on: the caller skips this file altogether then.
bool retrieve_ms_takes(const bm25cu_index *ix, const RetrieveArgs &a) {
    const int kmax = ix->knobs.get("MS_KMAX", 1024);  // larger k: the streaming kernel
    return ix->nonneg && a.k <= kmax && a.k <= 1024 && !((a.mask != nullptr || a.mask_bits != nullptr) && ix->knobs.get("MASK_FAST", 1) == 0) &&
           !ix->knobs.get("FORCE_GENERAL", 0);
}

int launch_retrieve_ms(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_fb_list, int *launches) {
    MsParams p;
    p.data = ix->d_data;
    p.indices = ix->d_indices;
    p.indptr = ix->d_indptr;
    p.colmax = ix->d_colmax;
    p.colkth = nullptr;
    if (ix->knobs.get("SEED_KTH", 1) && ix->d_colkth && a.mask == nullptr && a.mask_bits == nullptr) {
        if (a.k == 1) p.colkth = ix->d_colmax;
        else if (a.k <= 10) p.colkth = ix->d_colkth;
        else if (a.k <= 100) p.colkth = ix->d_colkth + (size_t)ix->n_vocab;
        else if (a.k <= 1000) p.colkth = ix->d_colkth + 2 * (size_t)ix->n_vocab;
    }
    const bool use_bt = ix->knobs.get("MS_BTAB", 1) && ix->d_btab != nullptr;
    p.btab = use_bt ? ix->d_btab : nullptr;
    p.btab_off = ix->d_btab_off;
    p.btab_sh = ix->d_btab_sh;
    p.pbm = (ix->knobs.get("MS_PBM", 1) && ix->d_pbm) ? ix->d_pbm : nullptr;
    p.pbm_off = ix->d_pbm_off;
    p.pbm_sh = ix->d_pbm_sh;
    p.stab = (use_bt && ix->knobs.get("MS_STAB", 1) && ix->d_stab) ? reinterpret_cast<const uint4 *>(ix->d_stab) : nullptr;
    p.stab_off = ix->d_stab_off;
    p.stab_sh = ix->d_stab_sh;
    p.nnz = ix->nnz + kPadElems;
    p.btab_len = ix->btab_len;
    p.pbm_len = ix->pbm_len;
    p.stab_len = ix->stab_len;
    p.partial_rows = 0;  // set below, once the number of parts is known
    p.n_docs = ix->n_docs;
    p.n_vocab = ix->n_vocab;
    p.doc_base = ix->doc_base;
    p.a = a;
    p.ctrl = d_ctrl;
    p.fb_list = d_fb_list;
    p.kcap = next_pow2(a.k < 32 ? 32 : a.k);
    p.route_all = retrieve_ms_takes(ix, a) ? 0 : 1;  // (the caller does not come here then; kept as a safe default)
    p.budget = (unsigned int)ix->knobs.get("MS_BUDGET", 1 << 20);
    p.theta_mode = ix->knobs.get("DEBUG_THETA", 0);
    p.theta_io = nullptr;
    if (p.theta_mode) {
        int rc0 = ix->d_theta.reserve((size_t)(a.n_queries > 0 ? a.n_queries : 1) * sizeof(float));
        if (rc0) return rc0;
        p.theta_io = ix->d_theta.as<float>();
    }
    p.big_merge_at = (unsigned int)ix->knobs.get("MS_BIGMERGE", 512);
    if (p.big_merge_at > 512u) p.big_merge_at = 512u;
    if (p.big_merge_at < 1u) p.big_merge_at = 1u;
    // plan: parts of at most `part_cost` postings (BM25CU_MS_PARTCOST), at most BM25CU_MS_MAXPARTS per query
    int max_parts = ix->knobs.get("MS_MAXPARTS", kMsMaxParts);
    if (max_parts < 1) max_parts = 1;
    if (max_parts > 255) max_parts = 255;
    if (max_parts > 8192 / p.kcap && !ix->knobs.has("MS_MAXPARTS")) max_parts = 8192 / p.kcap > 1 ? 8192 / p.kcap : 1;  // rows of k keys per part
    {   // the partial-result buffer (n_queries * max_parts rows of k keys) stays under 1 GiB: large k * large batches get fewer parts
        const size_t row_bytes = (size_t)(a.n_queries > 0 ? a.n_queries : 1) * (size_t)a.k * 8;
        const size_t fit = ((size_t)1 << 30) / row_bytes;
        if ((size_t)max_parts > fit) max_parts = fit > 1 ? (int)fit : 1;
    }
    // smaller batches and smaller shards are cut finer so that every CTA still gets several parts (NQ shape, 3 452 queries:
    // 1.34 -> 1.20 ms; a 1/8 shard of the MS-MARCO shape: 0.81 -> 0.68 ms)
    const int per_sm_cfg = ix->knobs.get("MS_CTAS", 9);
    const bool small = a.n_queries < 3 * ix->sm_count * per_sm_cfg || ix->nnz < 300000000ll;
    // larger top lists: every part has to fill a top list of its own before it prunes, so parts are made coarser
    // (k = 1000 at the MS-MARCO shape: 13.2 ms with 65 536, 9.5 ms with 524 288 postings per part; with the later kernel
    // 8.57 ms with 524 288, 8.22 ms with 1 M and above; k = 100: 3.83 ms with 262 144, 3.60 ms with 524 288, 4.29 ms with 1 M)
    const int k_scale = p.kcap <= 32 ? 1 : (p.kcap <= 64 ? 2 : (p.kcap <= 128 ? 8 : 16));
    long long part_cost = (long long)(small ? 32768 : 65536) * k_scale;
    if (p.kcap <= 32) {
        // top lists in registers: parts proportional to the batch's work, so that the number of parts per CTA stays put
        // (queries x postings of the shard as the measure; 32 768 .. 98 304 postings per part for one batch at a time:
        // 1/8 .. 1/1 of the MS-MARCO shape 0.50 / 0.79 / 1.29 / 2.15 ms). With L batches in flight on views of the shard
        // the other batches fill the GPU while one batch's last parts run, and coarser parts mean fewer thresholds
        // re-discovered and fewer partly filled probe rounds: (L + 1) / 2 times coarser, no upper limit
        // (profiles/round2_lanes_partcost_sweep.log; 3 lanes: 0.42 -> 0.39 / 0.70 -> 0.60 / 1.25 -> 0.97 / 2.04 -> 1.70 ms)
        const bm25cu_index *root = ix->root ? ix->root : ix;
        int in_flight = ix->knobs.get("IN_FLIGHT", 1 + root->n_views.load());
        if (in_flight < 1) in_flight = 1;
        double c = (double)a.n_queries * (double)ix->nnz / 1.75e7;
        if (c < 32768.0) c = 32768.0;
        if (in_flight == 1 && c > 98304.0) c = 98304.0;
        c *= 0.5 * (in_flight + 1);
        if (c > 1048576.0) c = 1048576.0;
        part_cost = (long long)c;
    }
    part_cost = (long long)ix->knobs.get("MS_PARTCOST", (int)part_cost);
    const size_t nq = (size_t)a.n_queries;
    int rc;
    if ((rc = ix->d_ms_items.reserve(nq * (size_t)max_parts * 8))) return rc;
    if ((rc = ix->d_ms_rows.reserve((nq + 1) * 4 + nq * 2 + 16))) return rc;
    if ((rc = ix->d_ms_q.reserve(nq * 16))) return rc;
    if ((rc = ix->d_ms_partial.reserve(nq * (size_t)max_parts * (size_t)a.k * 8))) return rc;
    unsigned long long *items = ix->d_ms_items.as<unsigned long long>();
    unsigned int *part_base = ix->d_ms_rows.as<unsigned int>();
    unsigned char *scratch = reinterpret_cast<unsigned char *>(part_base + nq + 1);
    unsigned long long *qtot = ix->d_ms_q.as<unsigned long long>();
    unsigned int *gtheta = reinterpret_cast<unsigned int *>(qtot + nq);
    unsigned int *qstat = gtheta + nq;
    const unsigned qblocks = (unsigned)((nq + 255) / 256);
    ms_plan_cost_kernel<<<qblocks, 256, 0, ix->stream>>>(a.q_tokens, a.q_ptr, a.n_queries, ix->d_indptr, ix->n_vocab, part_cost, max_parts,
                                                         part_base, qtot, gtheta, qstat, scratch, d_ctrl);
    BM25CU_CHECK(cudaGetLastError());
    ms_plan_scatter_kernel<<<qblocks, 256, 0, ix->stream>>>(a.n_queries, scratch, items, d_ctrl);
    BM25CU_CHECK(cudaGetLastError());
    p.items = items;
    p.part_base = part_base;
    p.gtheta = gtheta;
    p.qstat = qstat;
    p.partial = ix->d_ms_partial.as<unsigned long long>();
    p.partial_rows = (long long)nq * max_parts;
    const int per_sm = ix->knobs.get("MS_CTAS", 9);  // 55 registers: nine 128-thread CTAs are resident per SM
    const int grid = ix->sm_count * (per_sm > 0 ? per_sm : 9);
    const size_t dyn = p.kcap > 32 ? (size_t)2 * p.kcap * 8 : 0;
    const bool mb = a.mask_bits != nullptr, cp = ix->stab_compact && p.stab != nullptr;
    if (p.kcap <= 32) {
        BM25CU_CHECK((launch_ms_variant<128, 8, false>(p, mb, cp, grid, 0, ix->stream)));
    } else if (ix->knobs.get("MS_BIG_NT", 256) == 256) {
        // shared-memory top lists: CTAs of 256 threads, four per SM - twice the survivors per probe round and per merge of
        // the pending keys (k = 1000 at the MS-MARCO shape: 9.5 -> 8.7 ms)
        const int per_sm_b = ix->knobs.get("MS_CTAS", 4);
        BM25CU_CHECK((launch_ms_variant<256, 4, true>(p, mb, cp, ix->sm_count * (per_sm_b > 0 ? per_sm_b : 4), dyn, ix->stream)));
    } else {
        BM25CU_CHECK((launch_ms_variant<128, 8, true>(p, mb, cp, grid, dyn, ix->stream)));
    }
    BM25CU_CHECK(cudaGetLastError());
    if (p.kcap <= 32)
        ms_merge_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, ix->stream>>>(scratch + nq, part_base, p.partial, qstat, qtot, a, ix->doc_base,
                                                                                  d_ctrl, d_fb_list, p.theta_mode == 1 ? p.theta_io : nullptr);
    else
        ms_merge_big_kernel<<<(unsigned)nq, 128, (size_t)3 * p.kcap * 8, ix->stream>>>(scratch + nq, part_base, p.partial, qstat, qtot, a, ix->doc_base,
                                                                                      p.kcap, d_ctrl, d_fb_list);
    BM25CU_CHECK(cudaGetLastError());
    if (launches) *launches += 4;
    return BM25CU_OK;
}

}  // namespace bm25cu
