// K1: fused streaming scoring + top-k for a batch of queries (sm_100a).
//
// Replaces, per query, BM25._compute_relevance_from_scores (bm25s/__init__.py:312-318), the BM25L/BM25+ shift
// (bm25s/__init__.py:616-618) and selection.topk (bm25s/selection.py:14-37) without ever materialising a
// num_docs vector. Persistent CTAs (one per SM by default) pull queries from a device-side queue. For a query the
// document range is walked in *windows*; a window's postings of every query token are streamed from HBM into a
// double-buffered shared-memory stage by 1-D bulk async copies (cp.async.bulk + mbarrier: the TMA engine).
//
// Warp specialisation: warp 0 is the PRODUCER - it waits for a stage to land, derives the window (first staged
// document, coverage end of truncated tokens, postings per token inside the window), publishes that metadata,
// advances the per-token cursors and issues the bulk copies of the following window as soon as the consumers have
// released the other stage buffer. All other warps are CONSUMERS and only see ready windows.
//
// Exactness (bit-identical float32 scores): a document hit by ONE query token scores exactly that posting's value
// (0 + s = s), so no accumulator is needed for it. Ownership of a document inside a window is resolved through a
// one-byte tag per document in shared memory (P1: write tag = token slot; P2: losers mark 0x80|slot; P3: owners of
// unmarked tags score directly, the surviving marker takes charge of the contested document). Documents hit by >= 2
// tokens ("contested", a few percent) are re-summed in the reference's order - token by token in query order - from
// binary searches of the staged window postings (P4). Candidates above the running k-th best score go to a
// shared-memory candidate buffer of (score, ~doc) keys which is cut back to k by a bitonic sort; that key order also
// makes ties deterministic (lower doc id first), independent of the tiling.
//
// The kernel requires non-negative finite scores (true for the five BM25 variants; checked at upload), at most
// 64 query tokens and no weight mask; anything else is routed to the dense general kernel (general.cu).
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace bm25cu {

#ifdef BM25CU_PROFILE
#define PROF_DECL long long _pt = clock64(); unsigned long long _pc[32] = {0}
#define PROF(i) do { long long _n = clock64(); _pc[i] += (unsigned long long)(_n - _pt); _pt = _n; } while (0)
#define PROF_CNT(i, v) do { _pc[i] += (unsigned long long)(v); } while (0)
#define PROF_FLUSH(cond) do { if (cond) for (int _i = 0; _i < 32; ++_i) if (_pc[_i]) atomicAdd(&p.ctrl->prof[_i], _pc[_i]); } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_CNT(i, v)
#define PROF_FLUSH(cond)
#endif

constexpr int kMaxTok = 32;    // token slots per query on this path (tag byte: slot, 0x80 | slot, or 0xFF = clean)
constexpr int kMinCap = 16;    // minimum staged postings per token per window (multiple of 4)
constexpr unsigned char kClean = 0xFF;

struct FastParams {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    const float *colmax;
    int32_t n_docs, n_vocab, doc_base;
    RetrieveArgs a;
    Ctrl *ctrl;
    unsigned int *general_list;
    int tagcap;    // bytes of tag array == max doc span of a window
    int stagecap;  // postings per stage buffer
    int cbcap;     // candidate entries (power of two, >= 2 * next_pow2(k))
    int rcap;      // contested-doc list entries
    int partcap;   // floats of scratch for the contested-document partial scores
    unsigned long long *spill;  // per-CTA global overflow of the candidate buffer
    int spillcap;
    int route_all_general;
    int prune;     // 0: every token is "essential" (no threshold pruning), 1: essential / non-essential split
};

// ---------------------------------------------------------------- PTX wrappers (mbarrier + bulk async copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- shared-memory layout
struct SmemHdr {
    uint64_t bar_data[2];   // stage b landed            (producer waits)
    uint64_t bar_meta[2];   // window metadata of stage b published (consumers wait)
    uint64_t bar_free[2];   // consumers are done with stage b      (producer waits)
    long long t_cur[kMaxTok];
    long long t_end[kMaxTok];
    long long w_goff[2][kMaxTok];  // global posting index of the token's first posting in the window
    int t_cap[kMaxTok];
    int t_off[kMaxTok];       // element offset of the token's region inside a stage buffer
    float t_max[kMaxTok];     // largest score of the token's column
    float t_rest[kMaxTok];    // upper bound of what the OTHER query tokens can add to a document
    float t_nebound[kMaxTok]; // upper bound of a document hit only by this token and tokens with smaller maxima
    int t_skip[2][kMaxTok];   // leading elements to skip (16-byte alignment of the global source)
    int t_n[2][kMaxTok];      // valid staged postings
    int w_ne[2][kMaxTok];     // 1: token is non-essential in this window (scores not staged)
    int w_cnt[2][kMaxTok];    // postings of the token inside the window of stage b
    int w_off[2][kMaxTok];    // t_off + t_skip of stage b
    int w_preE[2][kMaxTok], w_endE[2][kMaxTok];    // flattened index space of the essential postings
    int w_preN[2][kMaxTok], w_endN[2][kMaxTok];    // flattened index space of the non-essential postings
    int w_base[2];            // first document of the window
    int w_last[2];            // 1 if this is the query's last window
    int w_ME[2], w_MN[2];     // essential / non-essential postings in the window
    int w_nE[2];              // essential tokens with postings in the window
    unsigned int w_kept[2];   // essential postings that survived the upper-bound filter (consumers count)
    unsigned int q_slot;
    unsigned int cb_cnt;
    unsigned int r_cnt;
    float theta;
    unsigned long long total_df;
};

__device__ __forceinline__ unsigned long long pack_cand(float s, int d) {
    return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned int)(~d);
}
__device__ __forceinline__ float cand_score(unsigned long long c) { return __uint_as_float((unsigned int)(c >> 32)); }
__device__ __forceinline__ int cand_doc(unsigned long long c) { return (int)~(unsigned int)(c & 0xffffffffu); }

template <int NT>
struct Fast {
    static constexpr int CNT = NT - 32;  // consumer threads
    static constexpr int NW = CNT / 32;  // consumer warps
    SmemHdr *h;
    unsigned long long *cb;   // [cbcap]
    int *rlist;               // [rcap]
    float *part;              // [partcap]
    int *sdocs;               // [2][stagecap]
    float *sscores;           // [2][stagecap]
    unsigned char *tag;       // [tagcap]
    const FastParams &p;
    unsigned long long *spill;

    __device__ Fast(const FastParams &pp, unsigned char *smem) : p(pp) {
        h = reinterpret_cast<SmemHdr *>(smem);
        size_t o = (sizeof(SmemHdr) + 15) & ~size_t(15);
        cb = reinterpret_cast<unsigned long long *>(smem + o);
        o += (size_t)p.cbcap * 8;
        rlist = reinterpret_cast<int *>(smem + o);
        o += (size_t)p.rcap * 4;
        part = reinterpret_cast<float *>(smem + o);
        o += (size_t)p.partcap * 4;
        o = (o + 15) & ~size_t(15);
        sdocs = reinterpret_cast<int *>(smem + o);
        o += (size_t)2 * p.stagecap * 4;
        sscores = reinterpret_cast<float *>(smem + o);
        o += (size_t)2 * p.stagecap * 4;
        tag = smem + o;
        spill = p.spill + (size_t)blockIdx.x * p.spillcap;
    }
    static size_t fixed_smem(const FastParams &p) {
        size_t o = (sizeof(SmemHdr) + 15) & ~size_t(15);
        o += (size_t)p.cbcap * 8 + (size_t)p.rcap * 4 + (size_t)p.partcap * 4;
        o = (o + 15) & ~size_t(15);
        o += (size_t)4 * p.stagecap * 4;
        return o;
    }

    // ---- synchronisation among the consumer warps only (named barrier 1) or the whole CTA
    template <bool CONS>
    __device__ __forceinline__ void bsync() {
        if (CONS) asm volatile("bar.sync 1, %0;" ::"n"(CNT) : "memory");
        else __syncthreads();
    }

    __device__ __forceinline__ void emit(float s, int d) {
        const unsigned int slot = atomicAdd(&h->cb_cnt, 1u);
        const unsigned long long c = pack_cand(s, d);
        if (slot < (unsigned)p.cbcap) cb[slot] = c;
        else if (slot - p.cbcap < (unsigned)p.spillcap) spill[slot - p.cbcap] = c;
    }

    // bitonic sort of key[0..P) descending (P power of two, <= cbcap), shared memory
    template <bool CONS>
    __device__ void sort_desc(unsigned long long *key, int P, int id, int nthr) {
        for (int size = 2; size <= P; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                bsync<CONS>();
                for (int t = id; t < (P >> 1); t += nthr) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const unsigned long long a = key[lo], b2 = key[hi];
                    if (desc ? (a < b2) : (a > b2)) { key[lo] = b2; key[hi] = a; }
                }
            }
        bsync<CONS>();
    }

    // Cut the candidate buffer back to the k best (key order: score desc, doc asc); theta = k-th best score.
    // Leaves cb[0..cb_cnt) sorted descending.
    template <bool CONS>
    __device__ void compact(int k, int id, int nthr) {
        bsync<CONS>();
        const unsigned int total = min(h->cb_cnt, (unsigned)(p.cbcap + p.spillcap));
        unsigned int n = min(total, (unsigned)p.cbcap);
        unsigned int spilled = total - n, sp_done = 0;
        for (;;) {
            const int P = next_pow2((int)(n > 0 ? n : 1));
            for (int i = (int)n + id; i < P; i += nthr) cb[i] = 0ull;
            sort_desc<CONS>(cb, P, id, nthr);
            unsigned int keep = min(n, (unsigned)k);
            if (sp_done >= spilled) {
                if (id == 0) {
                    h->cb_cnt = keep;
                    if (keep == (unsigned)k) h->theta = fmaxf(h->theta, cand_score(cb[k - 1]));
                }
                bsync<CONS>();
                return;
            }
            // rare: merge the spilled candidates in chunks of (cbcap - keep)
            const float th = (keep == (unsigned)k) ? cand_score(cb[k - 1]) : 0.0f;
            const unsigned int room = (unsigned)p.cbcap - keep;
            const unsigned int chunk = min(room, spilled - sp_done);
            if (id == 0) h->cb_cnt = keep;
            bsync<CONS>();
            for (unsigned int i = id; i < chunk; i += nthr) {
                const unsigned long long c = spill[sp_done + i];
                if (cand_score(c) > th || keep < (unsigned)k) cb[atomicAdd(&h->cb_cnt, 1u)] = c;
            }
            bsync<CONS>();
            n = h->cb_cnt;
            sp_done += chunk;
        }
    }

    // producer: stage the next window of every live token into buffer b. A token is non-essential when even a
    // document containing it and every token with a smaller column maximum cannot reach the current threshold:
    // only its document ids are staged then.
    __device__ void issue_loads(int b, int T) {
        const int lane = threadIdx.x & 31;
        const float theta_p = *reinterpret_cast<volatile float *>(&h->theta);
        if (lane < T) {
            const int t = lane;
            const long long cur = h->t_cur[t], rem = h->t_end[t] - cur;
            const int ne = (p.prune && h->t_nebound[t] < theta_p) ? 1 : 0;
            int n = 0, a = 0;
            if (rem > 0) {
                a = (int)(cur & 3);
                const int cap = h->t_cap[t];
                n = (int)min((long long)(cap - a), rem);
                const uint32_t bytes = (uint32_t)(((a + n + 3) & ~3) * 4);
                mbar_expect_tx(&h->bar_data[b], ne ? bytes : 2 * bytes);
                const size_t so = (size_t)b * p.stagecap + h->t_off[t];
                bulk_g2s(sdocs + so, p.indices + (cur - a), bytes, &h->bar_data[b]);
                if (!ne) bulk_g2s(sscores + so, p.data + (cur - a), bytes, &h->bar_data[b]);
            }
            h->t_skip[b][t] = a;
            h->t_n[b][t] = n;
            h->w_ne[b][t] = ne;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&h->bar_data[b]);
    }

    // number of entries < bound in sorted docs[0..n) (whole warp cooperates, result in every lane)
    __device__ __forceinline__ int warp_count_below(const int *docs, int n, int bound) {
        const int lane = threadIdx.x & 31;
        int lo = 0, hi = n;
        while (hi - lo > 32) {
            const int step = (hi - lo + 31) >> 5;
            const int idx = lo + lane * step;
            const bool pr = (idx < hi) && (docs[idx] < bound);
            const int c = __popc(__ballot_sync(0xffffffffu, pr));
            const int nlo = (c > 0) ? lo + (c - 1) * step + 1 : lo;
            const int cand_hi = lo + c * step;
            hi = (c < 32 && cand_hi < hi) ? cand_hi : hi;
            lo = nlo;
        }
        const int idx = lo + lane;
        const bool pr = (idx < hi) && (docs[idx] < bound);
        return lo + __popc(__ballot_sync(0xffffffffu, pr));
    }

    __device__ __forceinline__ int warp_excl_scan(int v, int &total) {
        const int lane = threadIdx.x & 31;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }

    // ------------------------------------------------------------------------------------------ producer warp
    __device__ void producer(int T, uint32_t &ph_data0, uint32_t &ph_data1, uint32_t &ph_free0, uint32_t &ph_free1) {
        const int lane = threadIdx.x & 31;
        int b = 0;
        int nwin = 0;
        for (;;) {
            if (b == 0) { mbar_wait(&h->bar_data[0], ph_data0); ph_data0 ^= 1; }
            else { mbar_wait(&h->bar_data[1], ph_data1); ph_data1 ^= 1; }
            const int *bd = sdocs + (size_t)b * p.stagecap;
            int n0 = 0, off0 = 0, first = 0x7fffffff, e = 0x7fffffff, last0 = -1, ne = 0;
            long long cur = 0, tend = 0;
            if (lane < T) {
                n0 = h->t_n[b][lane];
                off0 = h->t_off[lane] + h->t_skip[b][lane];
                ne = h->w_ne[b][lane];
                cur = h->t_cur[lane];
                tend = h->t_end[lane];
                if (n0 > 0) {
                    first = bd[off0];
                    last0 = bd[off0 + n0 - 1];
                    if (cur + n0 < tend) e = last0 + 1;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                e = min(e, __shfl_xor_sync(0xffffffffu, e, o));
            }
            const int base = first;
            const int cap_end = (base > 0x7fffffff - p.tagcap) ? 0x7fffffff : base + p.tagcap;
            const int wend = min(e, cap_end);
            int c0 = (n0 > 0 && last0 < wend) ? n0 : 0;
            // tokens whose staged postings reach beyond the window end need a search (rare)
            unsigned int need0 = __ballot_sync(0xffffffffu, n0 > 0 && last0 >= wend);
            while (need0) {
                const int src = __ffs(need0) - 1;
                need0 &= need0 - 1;
                const int so = __shfl_sync(0xffffffffu, off0, src), sn = __shfl_sync(0xffffffffu, n0, src);
                const int c = warp_count_below(bd + so, sn, wend);
                if (lane == src) c0 = c;
            }
            int ME, MN;
            const int wE = ne ? 0 : c0, wN = ne ? c0 : 0;
            const int preE = warp_excl_scan(wE, ME), preN = warp_excl_scan(wN, MN);
            const int nE = __popc(__ballot_sync(0xffffffffu, wE > 0));
            bool more = false;
            if (lane < T) {
                h->w_cnt[b][lane] = c0;
                h->w_off[b][lane] = off0;
                h->w_goff[b][lane] = cur;
                h->w_preE[b][lane] = preE;
                h->w_endE[b][lane] = preE + wE;
                h->w_preN[b][lane] = preN;
                h->w_endN[b][lane] = preN + wN;
                h->t_cur[lane] = cur + c0;
                more = (cur + c0) < tend;
            }
            const bool any_more = __any_sync(0xffffffffu, more);
            if (lane == 0) {
                h->w_base[b] = base;
                h->w_last[b] = any_more ? 0 : 1;
                h->w_ME[b] = ME;
                h->w_MN[b] = MN;
                h->w_nE[b] = nE;
                h->w_kept[b] = 0;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h->bar_meta[b]);  // release: metadata visible to the consumers
            ++nwin;
            if (!any_more) break;
            if (nwin >= 2) {  // the other stage was used by window nwin-2: wait until the consumers released it
                if (b == 0) { mbar_wait(&h->bar_free[1], ph_free1); ph_free1 ^= 1; }
                else { mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1; }
            }
            issue_loads(b ^ 1, T);
            b ^= 1;
        }
        // consume the releases not waited for above (keeps the barrier parities in step across queries)
        if (nwin >= 2) {
            const int bb = (nwin - 2) & 1;
            if (bb == 0) { mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1; }
            else { mbar_wait(&h->bar_free[1], ph_free1); ph_free1 ^= 1; }
        }
        {
            const int bb = (nwin - 1) & 1;
            if (bb == 0) { mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1; }
            else { mbar_wait(&h->bar_free[1], ph_free1); ph_free1 ^= 1; }
        }
    }

    // Visit the flattened positions [lo, hi) of one index space (pre/end per token slot) with the whole warp:
    // body(t, f, j) is called by the lane that owns position f; `off[t] - pre[t] + f` is the staged element.
    // j counts the warp's iterations (one bit of the per-thread masks), identical on every replay.
    template <typename Body>
    __device__ __forceinline__ void for_range(const int *pre, const int *end, int lo, int hi, Body body) {
        const int lane = threadIdx.x & 31;
        int t = 0, j = 0, f = lo;
        while (f < hi) {
            while (f >= end[t]) ++t;  // warp-uniform
            const int seg_hi = min(end[t], hi);
            const int pre_t = pre[t];
            for (; f < seg_hi; f += 32, ++j) {
                const int ff = f + lane;
                if (ff < seg_hi) body(t, ff - pre_t, j);
            }
            f = seg_hi;
        }
    }

    // ------------------------------------------------------------------------------------------ consumer warps
    __device__ void consumer(int T, int k, uint32_t &ph_meta0, uint32_t &ph_meta1) {
        const int ctid = threadIdx.x - 32;
        const int cw = ctid >> 5, lane = ctid & 31;
        const int trigger = p.cbcap / 2;
        float theta = 0.0f;
        int b = 0;
        bool first_window = true;
        PROF_DECL;
        for (;;) {
            PROF(0);
            if (b == 0) { mbar_wait(&h->bar_meta[0], ph_meta0); ph_meta0 ^= 1; }
            else { mbar_wait(&h->bar_meta[1], ph_meta1); ph_meta1 ^= 1; }
            PROF(1);
            PROF_CNT(16, 1);
            const int *bd = sdocs + (size_t)b * p.stagecap;
            const float *bs = sscores + (size_t)b * p.stagecap;
            const int base = h->w_base[b];
            const int last = h->w_last[b];
            const int *wcnt = h->w_cnt[b];
            const int *woff = h->w_off[b];
            const int ME = h->w_ME[b], MN = h->w_MN[b];

            if (first_window) {
                // Seed theta: the k-th largest score among the staged postings of the token with the most postings in
                // this window is reached by >= k distinct documents (scores are non-negative), so anything strictly
                // below it cannot be in the top-k. Avoids flooding the candidate buffer while theta is still 0.
                first_window = false;
                int best_t = 0, best_c = 0;
                for (int t = 0; t < T; ++t) { const int c = wcnt[t]; if (c > best_c) { best_c = c; best_t = t; } }
                const int m = min(best_c, p.cbcap);
                if (m >= k) {
                    const int P = next_pow2(m);
                    const float *sc = bs + woff[best_t];
                    for (int i = ctid; i < P; i += CNT) cb[i] = (i < m) ? ((unsigned long long)__float_as_uint(sc[i]) << 32) : 0ull;
                    sort_desc<true>(cb, P, ctid, CNT);
                    const unsigned int bits = (unsigned int)(cb[k - 1] >> 32);
                    theta = bits > 0u ? __uint_as_float(bits - 1u) : 0.0f;  // next float below the k-th sampled score
                    bsync<true>();
                    if (ctid == 0) h->theta = theta;
                }
            }

            PROF(2);
            PROF_CNT(17, ME); PROF_CNT(18, MN);
            // this warp's slice of the two flattened index spaces (32-bit: postings per window * warps < 2^31)
            const int eLo = (int)((unsigned)ME * (unsigned)cw / (unsigned)NW), eHi = (int)((unsigned)ME * (unsigned)(cw + 1) / (unsigned)NW);
            const int nLo = (int)((unsigned)MN * (unsigned)cw / (unsigned)NW), nHi = (int)((unsigned)MN * (unsigned)(cw + 1) / (unsigned)NW);
            const int *preE = h->w_preE[b], *endE = h->w_endE[b], *preN = h->w_preN[b], *endN = h->w_endN[b];
            unsigned long long keptE = 0ull, hitN = 0ull;

            // ---- P1: essential postings that can still reach the threshold claim their document's tag byte.
            //      A posting (d, s) of token t is dropped when s + (sum of the other tokens' column maxima) <= theta:
            //      no document containing it can enter the top-k.
            for_range(preE, endE, eLo, eHi, [&](int t, int i, int j) {
                const int o = woff[t] + i;
                const float s = bs[o];
                const float ub = __fmul_ru(__fadd_ru(s, h->t_rest[t]), 1.00002f);
                if (ub > theta) {
                    tag[bd[o] - base] = (unsigned char)t;
                    keptE |= 1ull << j;
                }
            });
            {
                int kc = __popcll(keptE);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) kc += __shfl_xor_sync(0xffffffffu, kc, o);
                if (lane == 0 && kc > 0) atomicAdd(&h->w_kept[b], (unsigned)kc);
            }
            PROF(3);
            bsync<true>();
            PROF(4);
            const bool active = h->w_kept[b] > 0;
            PROF_CNT(19, active ? 1 : 0); PROF_CNT(20, h->w_kept[b]);  // no surviving essential posting: nothing in this window can score
            if (active) {
                // ---- P2: essential postings that lost the claim mark the document as contested; non-essential
                //      postings probe the tag and mark documents that some essential posting claimed
                if (h->w_nE[b] >= 2 && keptE) {
                    for_range(preE, endE, eLo, eHi, [&](int t, int i, int j) {
                        if ((keptE >> j) & 1ull) {
                            const int o = bd[woff[t] + i] - base;
                            if (tag[o] != (unsigned char)t) tag[o] = (unsigned char)(0x80 | t);
                        }
                    });
                }
                for_range(preN, endN, nLo, nHi, [&](int t, int i, int j) {
                    const int o = bd[woff[t] + i] - base;
                    if (tag[o] != kClean) {
                        tag[o] = (unsigned char)(0x80 | t);
                        hitN |= 1ull << j;
                    }
                });
                PROF(5);
                bsync<true>();
                PROF(6);
                // ---- P3: uncontested owners score directly; the surviving marker owns the contested document
                if (keptE) {
                    for_range(preE, endE, eLo, eHi, [&](int t, int i, int j) {
                        if ((keptE >> j) & 1ull) {
                            const int o = woff[t] + i;
                            const int d = bd[o];
                            const unsigned char x = tag[d - base];
                            if (x == (unsigned char)t) {
                                const float s = bs[o];
                                if (s > theta) emit(s, d);
                            } else if (x == (unsigned char)(0x80 | t)) {
                                const unsigned int r = atomicAdd(&h->r_cnt, 1u);
                                if (r < (unsigned)p.rcap) rlist[r] = d;
                            }
                        }
                    });
                }
                if (hitN) {
                    for_range(preN, endN, nLo, nHi, [&](int t, int i, int j) {
                        if ((hitN >> j) & 1ull) {
                            const int d = bd[woff[t] + i];
                            if (tag[d - base] == (unsigned char)(0x80 | t)) {
                                const unsigned int r = atomicAdd(&h->r_cnt, 1u);
                                if (r < (unsigned)p.rcap) rlist[r] = d;
                            }
                        }
                    });
                }
                PROF(7);
                bsync<true>();
                PROF(8);
                // ---- P5: the essential postings clean the tags they claimed (tags are all-clean between windows)
                if (keptE) {
                    for_range(preE, endE, eLo, eHi, [&](int t, int i, int j) {
                        if ((keptE >> j) & 1ull) tag[bd[woff[t] + i] - base] = kClean;
                    });
                }
                // ---- P4: contested documents are summed in query-token order (the reference's float32 order).
                //      a) one thread per (document, token): binary search of the token's window postings
                //      b) one thread per document: ordered sum of the partial scores found
                const int R = min((int)h->r_cnt, p.rcap);
                PROF(9);
                PROF_CNT(21, R);
                if (R > 0) {
                    const int chunk_docs = max(1, p.partcap / T);
                    for (int r0 = 0; r0 < R; r0 += chunk_docs) {
                        const int nd = min(chunk_docs, R - r0);
                        const int pairs = nd * T;
                        for (int jj = ctid; jj < pairs; jj += CNT) {
                            const int r = jj / T, t = jj - r * T;
                            const int d = rlist[r0 + r];
                            const int c = wcnt[t];
                            float v = -1.0f;
                            if (c > 0) {
                                const int *dt = bd + woff[t];
                                int lo = 0, hi = c;
                                while (lo < hi) { const int m = (lo + hi) >> 1; if (dt[m] < d) lo = m + 1; else hi = m; }
                                if (lo < c && dt[lo] == d)
                                    v = h->w_ne[b][t] ? p.data[h->w_goff[b][t] + lo] : bs[woff[t] + lo];
                            }
                            part[jj] = v;
                        }
                        bsync<true>();
                        for (int r = ctid; r < nd; r += CNT) {
                            float sum = 0.0f;
                            const float *pv = part + r * T;
                            for (int t = 0; t < T; ++t) { const float v = pv[t]; if (v >= 0.0f) sum = __fadd_rn(sum, v); }
                            if (sum > theta) emit(sum, rlist[r0 + r]);
                        }
                        bsync<true>();
                    }
                } else {
                    bsync<true>();
                }
            }
            PROF(10);
            // all consumers are done with stage b and the tags of this window
            if (ctid == 0) { h->r_cnt = 0; mbar_arrive(&h->bar_free[b]); }
            if (h->cb_cnt >= (unsigned)trigger) { compact<true>(k, ctid, CNT); theta = fmaxf(theta, h->theta); PROF_CNT(22, 1); }
            PROF(11);
            if (last) break;
            b ^= 1;
        }
        PROF_FLUSH(ctid == 0);
    }

    __device__ void run() {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int k = p.a.k;
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(&h->bar_data[i], 1); mbar_init(&h->bar_meta[i], 1); mbar_init(&h->bar_free[i], 1); }
            fence_mbar_init();
        }
        for (int i = tid * 16; i < p.tagcap; i += NT * 16) *reinterpret_cast<uint4 *>(tag + i) = make_uint4(~0u, ~0u, ~0u, ~0u);
        __syncthreads();
        uint32_t ph_data0 = 0, ph_data1 = 0, ph_free0 = 0, ph_free1 = 0, ph_meta0 = 0, ph_meta1 = 0;

        PROF_DECL;
        for (;;) {
            PROF(12);
            if (tid == 0) h->q_slot = atomicAdd(&p.ctrl->fast_next, 1u);
            __syncthreads();
            const unsigned int slot = h->q_slot;
            if (slot >= (unsigned)p.a.n_queries) break;
            const int q = (int)slot;
            const long long qs = p.a.q_ptr[q], qe = p.a.q_ptr[q + 1];
            const long long Tl = qe - qs;
            if (p.route_all_general || Tl > kMaxTok) {
                if (tid == 0) p.general_list[atomicAdd(&p.ctrl->general_count, 1u)] = (unsigned)q;
                __syncthreads();
                continue;
            }
            const int T = (int)Tl;
            if (tid < T) {
                const int tok = p.a.q_tokens[qs + tid];
                long long beg = 0, end = 0;
                float mx = 0.0f;
                if ((unsigned)tok >= (unsigned)p.n_vocab) atomicOr(&p.ctrl->error, 1u);
                else { beg = p.indptr[tok]; end = p.indptr[tok + 1]; mx = p.colmax[tok]; }
                h->t_cur[tid] = beg;
                h->t_end[tid] = end;
                h->t_max[tid] = mx;
            }
            if (tid == 0) { h->cb_cnt = 0; h->theta = 0.0f; h->r_cnt = 0; }
            __syncthreads();
            if (warp == 0) {
                // stage capacities proportional to document frequency; pruning bounds from the column maxima
                const long long df0 = (lane < T) ? h->t_end[lane] - h->t_cur[lane] : 0;
                long long tot = df0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                const long long freecap = (long long)p.stagecap - (long long)T * kMinCap;
                int c0 = 0;
                if (lane < T) c0 = kMinCap + (int)(tot > 0 ? ((freecap * df0) / tot) & ~3ll : 0);
                int csum;
                const int off = warp_excl_scan(c0, csum);
                const float mx = (lane < T) ? h->t_max[lane] : 0.0f;
                double msum = (double)mx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
                // documents hit only by this token and tokens ranked below it (by column maximum, ties by slot)
                double below = (double)mx;
                for (int j = 0; j < T; ++j) {
                    const float mj = __shfl_sync(0xffffffffu, mx, j);
                    if (j != lane && (mj < mx || (mj == mx && j < lane))) below += (double)mj;
                }
                if (lane < T) {
                    h->t_cap[lane] = c0;
                    h->t_off[lane] = off;
                    // float32 sums of <= 32 non-negative terms exceed the exact sum by < 2^-18 relative: pad by 2^-15
                    h->t_rest[lane] = __double2float_ru((msum - (double)mx) * 1.00003);
                    h->t_nebound[lane] = __double2float_ru(below * 1.00003);
                }
                if (lane == 0) {
                    h->total_df = (unsigned long long)tot;
                    if (tot > 0) atomicAdd(&p.ctrl->posting_count, (unsigned long long)tot);
                }
                __syncwarp();
                if (tot > 0) issue_loads(0, T);
            }
            __syncthreads();
            PROF(13);
            if (h->total_df > 0) {
                if (warp == 0) producer(T, ph_data0, ph_data1, ph_free0, ph_free1);
                else consumer(T, k, ph_meta0, ph_meta1);
            }
            __syncthreads();
            PROF(14);
            finalize(q, T);
            PROF(15);
        }
        PROF_FLUSH(tid == 32);
    }

    // Sort the surviving candidates, pad with untouched documents if fewer than k matched, write the result row.
    __device__ void finalize(int q, int T) {
        const int tid = threadIdx.x;
        const int k = p.a.k;
        compact<false>(k, tid, NT);  // leaves cb[0..n) sorted descending, n <= k
        const int n = (int)h->cb_cnt;
        const bool has_shift = (p.a.shift != nullptr) && T > 0;  // empty query: zeros, no shift (__init__.py:653-657)
        const float shift = has_shift ? p.a.shift[q] : 0.0f;
        int32_t *oi = p.a.out_ids + (size_t)q * k;
        float *os = p.a.out_scores + (size_t)q * k;
        for (int j = tid; j < n; j += NT) {
            const unsigned long long c = cb[j];
            float s = cand_score(c);
            if (has_shift) s = __fadd_rn(s, shift);
            oi[j] = cand_doc(c) + p.doc_base;
            os[j] = s;
        }
        if (n < k) {
            // padding: the (k - n) lowest document ids that are not candidates score 0 (+ shift)
            __syncthreads();
            const int P = next_pow2(n > 0 ? n : 1);
            // sort the candidate doc ids ascending: keys ~doc descending
            for (int i = tid; i < P; i += NT) cb[i] = (i < n) ? (unsigned long long)(unsigned int)(~cand_doc(cb[i])) : 0ull;
            sort_desc<false>(cb, P, tid, NT);
            const float pad_score = has_shift ? __fadd_rn(0.0f, shift) : 0.0f;
            for (int j = n + tid; j < k; j += NT) {
                const long long m = j - n;  // m-th (0-based) non-candidate document id
                int lo = 0, hi = n;         // first i with doc_i - i > m   (doc_i ascending = ~cb[i])
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const long long dm = (long long)(unsigned int)(~(unsigned int)cb[mid]);
                    if (dm - mid > m) hi = mid; else lo = mid + 1;
                }
                const long long x = m + lo;
                if (x < p.n_docs) { oi[j] = (int32_t)x + p.doc_base; os[j] = pad_score; }
                else { oi[j] = -1; os[j] = __int_as_float(0xff800000); }
            }
        }
        __syncthreads();
    }
};

template <int NT>
__global__ void __launch_bounds__(NT) retrieve_fast_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Fast<NT> f(p, smem);
    f.run();
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int launch_retrieve_fast(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_general_list,
                         int *launches) {
    FastParams p;
    p.data = ix->d_data;
    p.indices = ix->d_indices;
    p.indptr = ix->d_indptr;
    p.colmax = ix->d_colmax;
    p.prune = env_int("BM25CU_PRUNE", 1);
    p.n_docs = ix->n_docs;
    p.n_vocab = ix->n_vocab;
    p.doc_base = ix->doc_base;
    p.a = a;
    p.ctrl = d_ctrl;
    p.general_list = d_general_list;
    const int threads = env_int("BM25CU_THREADS", 512);
    const int ctas_per_sm = env_int("BM25CU_CTAS_PER_SM", 1);
    const int kcap = next_pow2(a.k < 1 ? 1 : a.k);
    const int cb_min = env_int("BM25CU_CBMIN", 512);
    p.cbcap = 2 * kcap < cb_min ? next_pow2(cb_min) : 2 * kcap;
    p.route_all_general = (!ix->nonneg) || (a.mask != nullptr) || (a.k > 2048) || env_int("BM25CU_FORCE_GENERAL", 0);
    const int budget = ix->max_smem_optin / (ctas_per_sm > 0 ? ctas_per_sm : 1) - (ctas_per_sm > 1 ? 1024 : 0);
    int stage = env_int("BM25CU_STAGE", ctas_per_sm > 1 ? 2048 : 4096);
    stage = (stage + 3) & ~3;
    if (stage < kMaxTok * kMinCap + 64) stage = kMaxTok * kMinCap + 64;
    {   // the per-thread posting masks hold 64 warp iterations: (postings per consumer warp) / 32 + kMaxTok <= 64
        const int nw = ((threads == 64 || threads == 128 || threads == 256 || threads == 1024) ? threads : 512) / 32 - 1;
        const int max_stage = nw * 32 * (64 - kMaxTok - 2);
        if (stage > max_stage) stage = max_stage & ~3;
    }
    p.stagecap = stage;
    p.rcap = stage / 2 + 64;
    p.partcap = stage < 256 ? 256 : (stage > 4096 ? 4096 : stage);
    size_t fixed = Fast<512>::fixed_smem(p);
    long long tagcap = (long long)budget - (long long)fixed - 64;
    if (tagcap < 4096) {
        p.route_all_general = 1;  // k too large for this shared-memory budget
        p.cbcap = 512;
        fixed = Fast<512>::fixed_smem(p);
        tagcap = (long long)budget - (long long)fixed - 64;
    }
    const int tag_env = env_int("BM25CU_TAG", 0);
    if (tag_env > 0 && tag_env < tagcap) tagcap = tag_env;
    p.tagcap = (int)(tagcap & ~15ll);
    p.spillcap = p.stagecap + p.rcap;
    const int grid = ix->sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    int rc = ix->d_spill.reserve((size_t)grid * p.spillcap * 8);
    if (rc) return rc;
    p.spill = ix->d_spill.as<unsigned long long>();
    const size_t smem = fixed + p.tagcap;
    cudaError_t e;
#define BM25CU_LAUNCH(NTV)                                                                                          \
    e = cudaFuncSetAttribute(retrieve_fast_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e == cudaSuccess) retrieve_fast_kernel<NTV><<<grid, NTV, smem, ix->stream>>>(p);
    switch (threads) {
        case 64: BM25CU_LAUNCH(64); break;
        case 128: BM25CU_LAUNCH(128); break;
        case 256: BM25CU_LAUNCH(256); break;
        case 1024: BM25CU_LAUNCH(1024); break;
        default: BM25CU_LAUNCH(512); break;
    }
#undef BM25CU_LAUNCH
    BM25CU_CHECK(e);
    BM25CU_CHECK(cudaGetLastError());
    if (launches) ++*launches;
    return BM25CU_OK;
}

}  // namespace bm25cu
