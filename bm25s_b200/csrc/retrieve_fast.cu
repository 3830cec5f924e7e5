// K1: fused streaming scoring + top-k for a batch of queries (sm_100a).
//
// Replaces, per query, BM25._compute_relevance_from_scores (bm25s/__init__.py:312-318), the BM25L/BM25+ shift
// (bm25s/__init__.py:616-618) and selection.topk (bm25s/selection.py:14-37) without ever materialising a
// num_docs vector. Persistent CTAs (one per SM) pull queries from a device-side queue. For a query the document
// range is walked in *windows*; a window's postings of every query token are streamed from HBM into a
// double-buffered shared-memory pool by 1-D bulk async copies (cp.async.bulk + mbarrier: the TMA engine).
//
// Warp specialisation: warp 0 is the PRODUCER - it waits for a stage to land, derives the window (first staged
// document, coverage end of truncated tokens, postings per token inside the window), cuts the window's document range
// into one sub-range per consumer warp (per-token offsets by binary search), publishes that metadata, advances the
// per-token cursors and issues the bulk copies of the following window as soon as the consumers have released the other
// stage. All other warps are CONSUMERS: each owns a document sub-range of the window - its postings of every token,
// its tag bytes - and runs all passes on it with warp-level synchronisation only, so warps drift apart and hide each
// other's latencies; the consumers meet once per window (stage release, candidate-buffer compaction).
//
// Exactness (bit-identical float32 scores): a document hit by ONE query token scores exactly that posting's value
// (0 + s = s), so no accumulator is needed for it. Ownership inside a window is resolved through one tag byte per
// GROUP of G = 1, 2 or 4 consecutive documents (G grows when the claiming postings are sparse, which makes windows
// G times larger): tag = [contested:1][token slot:4][valid:1][doc low bits:2].
//   P1  every (surviving) essential posting claims its group:  tag = mine
//   P2  essential postings that read back something else, and non-essential postings that find their own document
//       claimed, set the contested bit
//   P3  a posting that reads back exactly `mine` is the only hit of its document: score = s. A contested winner, or
//       a posting whose document has no winner (another document of the group won), goes to the slow list
//   P4  slow list: the document is re-summed in query-token order - the reference's float32 order - from binary
//       searches of the staged window postings of every token (one thread per document)
//   P5  the claiming postings clean their tags (the tag array is all-clean between windows)
//
// Threshold pruning (exact): theta = running k-th best score. A token is NON-ESSENTIAL in a window when a document
// containing it and every token with a smaller column maximum cannot exceed theta: only its document ids are staged
// and its postings merely probe the tags (P2). An essential posting (d, s) is dropped in P1 when
// s + (sum of the other tokens' column maxima) <= theta. Bounds carry a 2^-15 relative pad against float32 rounding;
// a dropped posting can only belong to a document whose exact score is <= theta, so results equal exhaustive scoring.
//
// Candidates above theta go to a shared-memory buffer of (score, ~doc) keys which is cut back to k by a bitonic
// sort; that key order also makes ties deterministic (lower doc id first), independent of tiling and sharding.
//
// The kernel requires non-negative finite scores (true for the five BM25 variants; checked at upload), at most
// 15 query tokens, k <= 2048 and no weight mask; anything else is routed to the dense general kernel (general.cu).
#include <cstdlib>

#include "fast_common.cuh"

namespace bm25cu {


// ---------------------------------------------------------------- shared-memory layout
struct SmemHdr {
    uint64_t bar_data[2];   // stage b landed                        (producer waits)
    uint64_t bar_meta[2];   // window metadata of stage b published  (consumers wait)
    uint64_t bar_free[2];   // consumers are done with stage b       (producer waits)
    long long t_cur[kTokArr];
    long long t_end[kTokArr];
    long long t_df[kTokArr];
    long long w_goff[2][kTokArr];  // global posting index of the token's first posting in the window
    float t_max[kTokArr];     // largest score of the token's column
    float t_kth[kTokArr];     // k-th largest score of the token's column (0: unknown / fewer than k postings)
    float t_rest[kTokArr];    // upper bound of what the OTHER query tokens can add to a document
    float t_nebound[kTokArr]; // upper bound of a document hit only by this token and tokens with smaller maxima
    int t_n[2][kTokArr];      // staged postings of stage b (valid elements)
    int t_ne[2][kTokArr];     // 1: token is non-essential in stage b (scores not staged)
    int w_cnt[2][kTokArr];    // postings of the token inside the window
    int w_doc[2][kTokArr];    // pool slot of the first window document id
    int w_sc[2][kTokArr];     // pool slot of the first window score (essential tokens)
    int s_lo[kMaxWarps][kTokArr];          // per consumer warp: first / one-past-last window posting of token t in the warp's
    int s_hi[kMaxWarps][kTokArr];          // document sub-range (computed by the warp itself for the current window)
    int w_end[2];                          // one past the last document of the window
    int2 w_slow[kMaxWarps][32];            // per-warp slow list (doc, slot | rule << 8)
    float w_slow_s[kMaxWarps][32];         //                    score of the listing posting
    int w_base[2];            // first document of the window, rounded down to a multiple of G
    int w_shift[2];           // log2(G) of the window
    int w_last[2];            // 1 if this is the query's last window
    int w_ME[2], w_MN[2];     // essential / non-essential postings in the window
    unsigned int w_kept[2];   // warps with essential postings that survived the upper-bound filter
    unsigned int sel_hist[256];  // radix-select histogram of the candidate cut (large k)
    unsigned int sel_bin, sel_need, sel_cnt;
    unsigned int q_slot;
    unsigned int cb_cnt;
    unsigned int r_cnt;
    float theta;
    unsigned long long total_df;
};


// one token's window postings as seen by a consumer warp (shared-space byte addresses, kept in registers)
struct Seg {
    uint32_t docs;  // address of the first document id of the segment
    uint32_t sc;    // address of the first score (essential tokens)
    float rest;
    int t;
    int n;          // postings of this segment
};

// ---------------------------------------------------------------- per-slice passes
// `tagv` is the shared address of the (virtual) tag of group 0: the tag of document d is at tagv + (d >> sh).
// Four warp iterations are processed together: all loads of a group are issued before any tag store, and the rare
// store paths sit behind a warp vote so that they cost nothing when no lane needs them.
// A posting (d, s) of an essential token is KEPT when s + rest (the other tokens' column maxima) can exceed theta;
// theta is constant within a window, so every pass re-derives the same predicate (`allkept`: rest alone exceeds theta).
__device__ __forceinline__ unsigned tag_of(unsigned tbits, unsigned d, unsigned gmask) { return tbits | (d & gmask); }

// P1: kept postings claim their document's group. Returns (per lane) whether anything was kept.
__device__ __forceinline__ bool pass_claim(const Seg &g, float theta, bool allkept, int sh, unsigned gmask, uint32_t tagv,
                                           int lane) {
    const unsigned tbits = (unsigned)(g.t << 3) | kTagValid;
    const int iters = (g.n + 31) >> 5;
    uint32_t ad = g.docs + (lane << 2), as = g.sc + (lane << 2);
    int rem = g.n - lane;  // > 32 * u  <=>  element (it + u) * 32 + lane exists
    bool any = false;
    for (int it = 0; it < iters; it += 4, ad += 512, as += 512, rem -= 128) {
        float sv[4];
        unsigned d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool v = rem > 32 * u;
            sv[u] = (v && !allkept) ? ldsf(as + 128 * u) : 0.0f;
            d[u] = v ? lds32(ad + 128 * u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (rem > 32 * u && (allkept || kept_posting(sv[u], g.rest, theta))) {
                sts8(tagv + (d[u] >> sh), tag_of(tbits, d[u], gmask));
                any = true;
            }
        }
    }
    return any;
}

// P2 (essential): kept postings that do not read back their own claim set the contested bit
__device__ __forceinline__ void pass_mark_e(const Seg &g, float theta, bool allkept, int sh, unsigned gmask, uint32_t tagv,
                                            int lane) {
    const unsigned tbits = (unsigned)(g.t << 3) | kTagValid;
    const int iters = (g.n + 31) >> 5;
    uint32_t ad = g.docs + (lane << 2), as = g.sc + (lane << 2);
    int rem = g.n - lane;
    for (int it = 0; it < iters; it += 4, ad += 512, as += 512, rem -= 128) {
        unsigned d[4], x[4];
        bool v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = rem > 32 * u;
            if (v[u] && !allkept) v[u] = kept_posting(ldsf(as + 128 * u), g.rest, theta);
            d[u] = v[u] ? lds32(ad + 128 * u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = v[u] ? lds8(tagv + (d[u] >> sh)) : 0u;
        unsigned bad = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u] && x[u] != tag_of(tbits, d[u], gmask)) bad |= 1u << u;
        if (__ballot_sync(0xffffffffu, bad != 0u)) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((bad >> u) & 1u) sts8(tagv + (d[u] >> sh), x[u] | kTagContested);
        }
    }
}

// P2 (non-essential): postings that find their own document claimed set the contested bit
__device__ __forceinline__ void pass_probe_n(const Seg &g, int sh, unsigned gmask, uint32_t tagv, int lane) {
    const int iters = (g.n + 31) >> 5;
    uint32_t ad = g.docs + (lane << 2);
    int rem = g.n - lane;
    for (int it = 0; it < iters; it += 4, ad += 512, rem -= 128) {
        unsigned d[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] = rem > 32 * u ? lds32(ad + 128 * u) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = rem > 32 * u ? lds8(tagv + (d[u] >> sh)) : 0u;
        unsigned hit = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if ((x[u] & 7u) == (kTagValid | (d[u] & gmask))) hit |= 1u << u;  // invalid lanes read x = 0: never valid
        if (__ballot_sync(0xffffffffu, hit != 0u)) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((hit >> u) & 1u) sts8(tagv + (d[u] >> sh), x[u] | kTagContested);
        }
    }
}

// P5: kept postings clean the tags they (or a posting of the same group) claimed
__device__ __forceinline__ void pass_clean(const Seg &g, float theta, bool allkept, int sh, uint32_t tagv, int lane) {
    const int iters = (g.n + 31) >> 5;
    uint32_t ad = g.docs + (lane << 2), as = g.sc + (lane << 2);
    int rem = g.n - lane;
    for (int it = 0; it < iters; it += 4, ad += 512, as += 512, rem -= 128) {
        unsigned d[4];
        bool v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = rem > 32 * u;
            if (v[u] && !allkept) v[u] = kept_posting(ldsf(as + 128 * u), g.rest, theta);
            d[u] = v[u] ? lds32(ad + 128 * u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u]) sts8(tagv + (d[u] >> sh), 0u);
    }
}

// ---------------------------------------------------------------- chunk passes (essential postings)
// A consumer warp's essential postings of a window are cut into chunks of <= 32 consecutive postings of one token
// (descriptor: docs address, scores address, count | slot << 8 | allkept << 16, rest). Four chunks are processed
// together, one posting per lane and chunk, so short per-token slices still fill the unrolled group.
struct Chunk4 {
    uint32_t docs[4], sc[4];
    float rest[4];
    unsigned tbits[4];
    int cnt[4];
    bool allkept[4];
};
__device__ __forceinline__ void load_chunks(uint32_t chunk_s, int c0, int nch, Chunk4 &c) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        c.cnt[u] = 0;
        c.docs[u] = 0; c.sc[u] = 0; c.rest[u] = 0.0f; c.tbits[u] = 0; c.allkept[u] = true;
        if (c0 + u < nch) {
            const uint4 d = lds128(chunk_s + 16u * (unsigned)(c0 + u));
            c.docs[u] = d.x;
            c.sc[u] = d.y;
            c.cnt[u] = (int)(d.z & 0xFFu);
            c.tbits[u] = (((d.z >> 8) & 0xFFu) << 3) | kTagValid;
            c.allkept[u] = ((d.z >> 16) & 1u) != 0u;
            c.rest[u] = __uint_as_float(d.w);
        }
    }
}

// P1 over chunks: kept postings claim their document's group. Returns (per lane) whether anything was kept.
__device__ __forceinline__ bool chunks_claim(uint32_t chunk_s, int nch, float theta, int sh, unsigned gmask, uint32_t tagv, int lane) {
    bool any = false;
    for (int c0 = 0; c0 < nch; c0 += 4) {
        Chunk4 c;
        load_chunks(chunk_s, c0, nch, c);
        unsigned d[4];
        float sv[4];
        bool v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = lane < c.cnt[u];
            sv[u] = (v[u] && !c.allkept[u]) ? ldsf(c.sc[u] + 4u * lane) : 0.0f;
            d[u] = v[u] ? lds32(c.docs[u] + 4u * lane) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v[u] && (c.allkept[u] || kept_posting(sv[u], c.rest[u], theta))) {
                sts8(tagv + (d[u] >> sh), tag_of(c.tbits[u], d[u], gmask));
                any = true;
            }
        }
    }
    return any;
}

// P2 (essential) over chunks: kept postings that do not read back their own claim set the contested bit
__device__ __forceinline__ void chunks_mark(uint32_t chunk_s, int nch, float theta, int sh, unsigned gmask, uint32_t tagv, int lane) {
    for (int c0 = 0; c0 < nch; c0 += 4) {
        Chunk4 c;
        load_chunks(chunk_s, c0, nch, c);
        unsigned d[4], x[4];
        bool v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = lane < c.cnt[u];
            if (v[u] && !c.allkept[u]) v[u] = kept_posting(ldsf(c.sc[u] + 4u * lane), c.rest[u], theta);
            d[u] = v[u] ? lds32(c.docs[u] + 4u * lane) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = v[u] ? lds8(tagv + (d[u] >> sh)) : 0u;
        unsigned bad = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u] && x[u] != tag_of(c.tbits[u], d[u], gmask)) bad |= 1u << u;
        if (__ballot_sync(0xffffffffu, bad != 0u)) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((bad >> u) & 1u) sts8(tagv + (d[u] >> sh), x[u] | kTagContested);
        }
    }
}

// P5 over chunks: kept postings clean the tags
__device__ __forceinline__ void chunks_clean(uint32_t chunk_s, int nch, float theta, int sh, uint32_t tagv, int lane) {
    for (int c0 = 0; c0 < nch; c0 += 4) {
        Chunk4 c;
        load_chunks(chunk_s, c0, nch, c);
        unsigned d[4];
        bool v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = lane < c.cnt[u];
            if (v[u] && !c.allkept[u]) v[u] = kept_posting(ldsf(c.sc[u] + 4u * lane), c.rest[u], theta);
            d[u] = v[u] ? lds32(c.docs[u] + 4u * lane) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u]) sts8(tagv + (d[u] >> sh), 0u);
    }
}

// P2 (non-essential), vectorised: each lane probes four consecutive document ids per 16-byte load. `gs` points to the
// global score of the first of the four postings: a hit prefetches its score into L2, the slow path reads it later.
__device__ __forceinline__ void prefetch_l2(const float *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void probe4(const uint4 dv, int sh, unsigned gmask, uint32_t tagv, const float *gs) {
    const unsigned d[4] = {dv.x, dv.y, dv.z, dv.w};
    unsigned x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = lds8(tagv + (d[u] >> sh));
    unsigned hit = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if ((x[u] & 7u) == (kTagValid | (d[u] & gmask))) hit |= 1u << u;
    if (hit) {  // rare
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if ((hit >> u) & 1u) { sts8(tagv + (d[u] >> sh), x[u] | kTagContested); prefetch_l2(gs + u); }
    }
}
__device__ __forceinline__ void pass_probe_n_vec(const Seg &g, int sh, unsigned gmask, uint32_t tagv, int lane, const float *gs) {
    // head: scalar elements up to the first 16-byte boundary
    const int head = min(g.n, (int)((4u - ((g.docs >> 2) & 3u)) & 3u));
    if (lane < head) {
        const unsigned d = lds32(g.docs + 4u * lane);
        const unsigned x = lds8(tagv + (d >> sh));
        if ((x & 7u) == (kTagValid | (d & gmask))) { sts8(tagv + (d >> sh), x | kTagContested); prefetch_l2(gs + lane); }
    }
    const int nbody = g.n - head;
    const int nv = nbody >> 2;  // full vectors
    const uint32_t vb = g.docs + 4u * head;
    const float *gv = gs + head;
    int i = lane;
    for (; i + 32 < nv; i += 64) {  // two vectors per lane in flight
        const uint4 a = lds128(vb + 16u * i), b2 = lds128(vb + 16u * (i + 32));
        probe4(a, sh, gmask, tagv, gv + 4 * i);
        probe4(b2, sh, gmask, tagv, gv + 4 * (i + 32));
    }
    if (i < nv) probe4(lds128(vb + 16u * i), sh, gmask, tagv, gv + 4 * i);
    // tail
    const int tail0 = head + (nv << 2);
    if (lane < g.n - tail0) {
        const unsigned d = lds32(g.docs + 4u * (tail0 + lane));
        const unsigned x = lds8(tagv + (d >> sh));
        if ((x & 7u) == (kTagValid | (d & gmask))) { sts8(tagv + (d >> sh), x | kTagContested); prefetch_l2(gs + tail0 + lane); }
    }
}

template <int NT>
struct Fast {
    static constexpr int CNT = NT - 32;  // consumer threads
    static constexpr int NW = CNT / 32;  // consumer warps
    SmemHdr *h;
    unsigned long long *cb;   // [cbcap]
    uint4 *chunks;            // [nwarps][kMaxChunk] essential-posting chunk descriptors
    int *pool;                // [2][poolcap] 4-byte slots: document ids and (essential tokens) float scores
    unsigned char *tag;       // [tagcap]
    const FastParams &p;
    unsigned long long *spill;

    __device__ Fast(const FastParams &pp, unsigned char *smem) : p(pp) {
        h = reinterpret_cast<SmemHdr *>(smem);
        size_t o = (sizeof(SmemHdr) + 15) & ~size_t(15);
        cb = reinterpret_cast<unsigned long long *>(smem + o);
        o += (size_t)p.cbcap * 8;
        o = (o + 15) & ~size_t(15);
        chunks = reinterpret_cast<uint4 *>(smem + o);
        o += (size_t)p.nwarps * kMaxChunk * 16;
        pool = reinterpret_cast<int *>(smem + o);
        o += (size_t)p.stages * p.poolcap * 4;
        tag = smem + o;
        spill = p.spill + (size_t)blockIdx.x * p.spillcap;
    }
    static size_t fixed_smem(const FastParams &p) {
        size_t o = (sizeof(SmemHdr) + 15) & ~size_t(15);
        o += (size_t)p.cbcap * 8;
        o = (o + 15) & ~size_t(15);
        o += (size_t)p.nwarps * kMaxChunk * 16;
        o += (size_t)p.stages * p.poolcap * 4;
        return o;
    }

    // ---- synchronisation among the consumer warps only (named barrier 1) or the whole CTA
    template <bool CONS>
    __device__ __forceinline__ void bsync() {
        if (CONS) asm volatile("bar.sync 1, %0;" ::"n"(CNT) : "memory");
        else __syncthreads();
    }

    __device__ __forceinline__ void emit(float s, int d) {
        if (p.a.mask) {  // weight mask in [0, 1]: numpy multiplies in the promoted type and rounds back to float32
            s = (float)((double)s * p.a.mask[d]);
            if (!(s > *reinterpret_cast<volatile float *>(&h->theta))) return;
        } else if (p.a.mask_bits) {  // 0/1 mask, one bit per document: a removed document scores 0 and never beats theta >= 0
            if (!((__ldg(p.a.mask_bits + ((unsigned)d >> 5)) >> ((unsigned)d & 31u)) & 1u)) return;
        }
        const unsigned int slot = atomicAdd(&h->cb_cnt, 1u);
        const unsigned long long c = pack_cand(s, d);
        if (slot < (unsigned)p.cbcap) cb[slot] = c;
        else if (slot - p.cbcap < (unsigned)p.spillcap) spill[slot - p.cbcap] = c;
        else atomicOr(&p.ctrl->error, 4u);  // cannot happen: a window emits at most poolcap / 2 candidates between two cuts
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

    // Keeps the min(n, k) largest of cb[0..n) in cb[0..min(n, k)) and returns the k-th largest key (0 if n < k).
    // Small buffers (or want_sorted) are sorted; large ones are cut by an MSB-first radix select on the 64-bit keys
    // (keys are unique, so exactly k survive) followed by an in-place partition - no sort until the query's final cut.
    template <bool CONS>
    __device__ unsigned long long cut_to_k(unsigned int n, int k, int id, int nthr, bool want_sorted) {
        const int P = next_pow2((int)(n > 0 ? n : 1));
        if (want_sorted || P < 512 || n <= (unsigned)k) {
            for (int i = (int)n + id; i < P; i += nthr) cb[i] = 0ull;
            sort_desc<CONS>(cb, P, id, nthr);
            return (n >= (unsigned)k) ? cb[k - 1] : 0ull;
        }
        unsigned long long prefix = 0ull, mask = 0ull;
        unsigned int need = (unsigned int)k;
        for (int pass = 7; pass >= 0; --pass) {
            const int shift = 8 * pass;
            for (int i = id; i < 256; i += nthr) h->sel_hist[i] = 0u;
            bsync<CONS>();
            for (unsigned int i = id; i < n; i += nthr) {
                const unsigned long long key = cb[i];
                if ((key & mask) == prefix) atomicAdd(&h->sel_hist[(unsigned int)(key >> shift) & 0xFFu], 1u);
            }
            bsync<CONS>();
            if (id < 32) {  // lane L owns bins 8 L .. 8 L + 7; the bin holding the need-th largest live key, from the top
                unsigned int cnt[8], lane_sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { cnt[j] = h->sel_hist[id * 8 + j]; lane_sum += cnt[j]; }
                unsigned int above = lane_sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_down_sync(0xffffffffu, above, o); if (id + o < 32) above += y; }
                above -= lane_sum;
                if (above < need && need <= above + lane_sum) {
                    unsigned int acc = above;
                    for (int j = 7; j >= 0; --j) {
                        if (acc + cnt[j] >= need) { h->sel_bin = (unsigned int)(id * 8 + j); h->sel_need = need - acc; break; }
                        acc += cnt[j];
                    }
                }
            }
            bsync<CONS>();
            prefix |= (unsigned long long)h->sel_bin << shift;
            mask |= 0xFFull << shift;
            need = h->sel_need;
        }
        const unsigned long long thr = prefix;  // the k-th largest key
        if (id == 0) h->sel_cnt = 0u;
        bsync<CONS>();
        // in-place partition, one chunk of nthr elements at a time: a survivor is written below the number of elements
        // already read, so nothing unread is overwritten
        for (unsigned int base = 0; base < n; base += (unsigned int)nthr) {
            const unsigned int i = base + (unsigned int)id;
            const unsigned long long key = (i < n) ? cb[i] : 0ull;
            bsync<CONS>();
            if (i < n && key >= thr) cb[atomicAdd(&h->sel_cnt, 1u)] = key;
        }
        bsync<CONS>();
        return thr;
    }

    // Cut the candidate buffer back to the k best (key order: score desc, doc asc); theta = k-th best score.
    // want_sorted: leaves cb[0..cb_cnt) sorted descending (the query's final cut).
    template <bool CONS>
    __device__ void compact(int k, int id, int nthr, bool want_sorted = false) {
        bsync<CONS>();
        const unsigned int total = min(h->cb_cnt, (unsigned)(p.cbcap + p.spillcap));
        unsigned int n = min(total, (unsigned)p.cbcap);
        const unsigned int spilled = total - n;
        unsigned int sp_done = 0;
        for (;;) {
            const bool final_cut = sp_done >= spilled;
            const unsigned long long kth = cut_to_k<CONS>(n, k, id, nthr, want_sorted && final_cut);
            const unsigned int keep = min(n, (unsigned)k);
            if (final_cut) {
                if (id == 0) {
                    h->cb_cnt = keep;
                    if (keep == (unsigned)k) h->theta = fmaxf(h->theta, cand_score(kth));
                }
                bsync<CONS>();
                return;
            }
            // Spilled candidates: those that still beat the k-th key are poured into the free room tile by tile; the
            // buffer is cut again only when the room is (nearly) used up - most tiles are filtered out entirely.
            const unsigned long long th_key = (keep == (unsigned)k) ? kth : 0ull;  // full key: score, then lower doc id
            if (id == 0) h->cb_cnt = keep;
            unsigned int cur = keep;
            for (;;) {
                bsync<CONS>();  // cb_cnt settled and read by everyone before the next tile adds to it
                const unsigned int chunk = min((unsigned)p.cbcap - cur, spilled - sp_done);
                for (unsigned int i = id; i < chunk; i += nthr) {
                    const unsigned long long c = spill[sp_done + i];
                    if (c > th_key) cb[atomicAdd(&h->cb_cnt, 1u)] = c;
                }
                sp_done += chunk;
                bsync<CONS>();
                cur = h->cb_cnt;
                if (sp_done >= spilled || (unsigned)p.cbcap - cur < 64u) break;
            }
            bsync<CONS>();
            n = cur;
        }
    }

    __device__ __forceinline__ int warp_excl_scan(int v, int &total) {
        const int lane = threadIdx.x & 31;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }

    // producer: stage the next window of every live token into stage b. Pool slots are shared out in proportion to
    // the column lengths (an essential token needs two slots per posting, a non-essential one a single slot). The
    // share-out (cap / offset per token, tag granularity) only changes when a token runs dry or turns non-essential,
    // so it is cached in the producer lanes' registers (lane == token slot) and recomputed on change.
    struct Layout {
        int cap, off, ne, live, sh;
    };
    __device__ __forceinline__ void issue_loads(int b, int T, Layout &L) {
        const int lane = threadIdx.x & 31;
        const float theta_p = *reinterpret_cast<volatile float *>(&h->theta);
        long long cur = 0, rem = 0;
        int ne = 0;
        if (lane < T) {
            cur = h->t_cur[lane];
            rem = h->t_end[lane] - cur;
            ne = (p.prune && h->t_nebound[lane] < theta_p) ? 1 : 0;
        }
        const int live = rem > 0 ? 1 : 0;
        if (__any_sync(0xffffffffu, ne != L.ne || live != L.live)) {
            long long w = 0, dfE = 0;
            if (live) {
                w = h->t_df[lane] * (ne ? 1 : 2);
                if (!ne) dfE = h->t_df[lane];
            }
            long long wsum = w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                dfE += __shfl_xor_sync(0xffffffffu, dfE, o);
            }
            const int nlive = __reduce_add_sync(0xffffffffu, live);
            int cap = 0, slots = 0;
            if (live) {
                const long long freeslots = (long long)p.poolcap - (long long)nlive * (2 * kMinCap);
                const long long share = (freeslots * w) / wsum;        // slots
                cap = kMinCap + (int)((share / (ne ? 1 : 2)) & ~3ll);  // postings, multiple of 4
                slots = cap * (ne ? 1 : 2);
            }
            int tot;
            L.off = warp_excl_scan(slots, tot);
            L.cap = cap;
            L.ne = ne;
            L.live = live;
            // tag granularity G = 2^sh: a claiming posting shares its group with G - 1 neighbours, each claimed with
            // the density of the essential postings; keep (G - 1) * density <= 1/32
            int sh = 0;
            const long long nd = p.n_docs > 0 ? p.n_docs : 1;
            while (sh < p.max_shift && ((long long)((2 << sh) - 1) * dfE * 32 <= nd)) ++sh;
            L.sh = sh;
        }
        if (lane < T) {
            int n = 0, a = 0;
            if (live) {
                a = (int)(cur & 3);
                n = (int)min((long long)(L.cap - a), rem);
                const uint32_t bytes = (uint32_t)(((a + n + 3) & ~3) * 4);
                mbar_expect_tx(&h->bar_data[b], ne ? bytes : 2 * bytes);
                int *dst = pool + (size_t)b * p.poolcap + L.off;
                bulk_g2s(dst, p.indices + (cur - a), bytes, &h->bar_data[b]);
                if (!ne) bulk_g2s(dst + L.cap, p.data + (cur - a), bytes, &h->bar_data[b]);
            }
            h->t_n[b][lane] = n;
            h->t_ne[b][lane] = ne;
            h->w_doc[b][lane] = L.off + a;
            h->w_sc[b][lane] = L.off + L.cap + a;
        }
        if (lane == 0) h->w_shift[b] = L.sh;
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

    // ------------------------------------------------------------------------------------------ producer warp
    __device__ void producer(int T, Layout &L, uint32_t &ph_data0, uint32_t &ph_data1, uint32_t &ph_free0, uint32_t &ph_free1) {
        const int lane = threadIdx.x & 31;
        int b = 0;
        int nwin = 0;
        PROF_DECL;
        for (;;) {
            PROF(24);
            if (b == 0) { mbar_wait(&h->bar_data[0], ph_data0); ph_data0 ^= 1; }
            else { mbar_wait(&h->bar_data[1], ph_data1); ph_data1 ^= 1; }
            PROF(25);
            const int *bp = pool + (size_t)b * p.poolcap;
            const uint32_t pool_s = smem_u32(bp);
            int n0 = 0, off0 = 0, first = 0x7fffffff, e = 0x7fffffff, last0 = -1, ne = 0;
            long long cur = 0, tend = 0;
            if (lane < T) {
                n0 = h->t_n[b][lane];
                off0 = h->w_doc[b][lane];
                ne = h->t_ne[b][lane];
                cur = h->t_cur[lane];
                tend = h->t_end[lane];
                if (n0 > 0) {
                    first = bp[off0];
                    last0 = bp[off0 + n0 - 1];
                    if (cur + n0 < tend) e = last0 + 1;
                }
            }
            const int lastmax = __reduce_max_sync(0xffffffffu, last0);
            first = __reduce_min_sync(0xffffffffu, first);
            e = __reduce_min_sync(0xffffffffu, e);
            const int sh = h->w_shift[b];
            const int base = first & ~((1 << sh) - 1);
            const long long span = (long long)p.tagcap << sh;
            const int cap_end = ((long long)base + span > 0x7fffffffll) ? 0x7fffffff : (int)(base + span);
            int wend = min(e, cap_end);
            if (lastmax < wend - 1) wend = lastmax + 1;  // tight end: the sub-ranges below split real documents only
            int c0 = (n0 > 0 && last0 < wend) ? n0 : 0;
            // tokens whose staged postings reach beyond the window end need a search (rare)
            unsigned int need0 = __ballot_sync(0xffffffffu, n0 > 0 && last0 >= wend);
            while (need0) {
                const int src = __ffs(need0) - 1;
                need0 &= need0 - 1;
                const int so = __shfl_sync(0xffffffffu, off0, src), sn = __shfl_sync(0xffffffffu, n0, src);
                const int c = warp_count_below(bp + so, sn, wend);
                if (lane == src) c0 = c;
            }
            const int ME = __reduce_add_sync(0xffffffffu, ne ? 0 : c0), MN = __reduce_add_sync(0xffffffffu, ne ? c0 : 0);
            bool more = false;
            if (lane < T) {
                h->w_cnt[b][lane] = c0;
                h->w_goff[b][lane] = cur;
                h->t_cur[lane] = cur + c0;
                more = (cur + c0) < tend;
            }
            const bool any_more = __any_sync(0xffffffffu, more);
            if (lane == 0) {
                h->w_base[b] = base;
                h->w_end[b] = wend;
                h->w_last[b] = any_more ? 0 : 1;
                h->w_ME[b] = ME;
                h->w_MN[b] = MN;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h->bar_meta[b]);  // release: metadata visible to the consumers
            PROF(26);
            ++nwin;
            if (!any_more) break;
            if (p.stages == 1) {
                // single stage: the next window is loaded into the same buffer once the consumers are done with this one
                mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1;
                PROF(27);
                issue_loads(0, T, L);
                PROF(28);
                continue;
            }
            if (nwin >= 2) {  // the other stage was used by window nwin-2: wait until the consumers released it
                if (b == 0) { mbar_wait(&h->bar_free[1], ph_free1); ph_free1 ^= 1; }
                else { mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1; }
            }
            PROF(27);
            issue_loads(b ^ 1, T, L);
            PROF(28);
            b ^= 1;
        }
        PROF_FLUSH(lane == 0);
        // consume the releases not waited for above (keeps the barrier parities in step across queries)
        if (p.stages == 1) {
            mbar_wait(&h->bar_free[0], ph_free0); ph_free0 ^= 1;
            return;
        }
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

    // This warp's slice of token t in stage b (its document sub-range), as shared-space addresses
    __device__ __forceinline__ Seg slice_of(int b, int t, int cw, uint32_t pool_s) {
        const int o0 = h->s_lo[cw][t], o1 = h->s_hi[cw][t];
        Seg g;
        g.docs = pool_s + 4u * (unsigned)(h->w_doc[b][t] + o0);
        g.sc = pool_s + 4u * (unsigned)(h->w_sc[b][t] + o0);
        g.rest = h->t_rest[t];
        g.t = t;
        g.n = o1 - o0;
        return g;
    }

    // P3 on one slice: sole hits score directly; contested winners and documents without a winner go to the warp's
    // slow list, which is resolved whenever it fills up (and once more at the end of the window).
    __device__ __forceinline__ void pass_resolve(const Seg &g, float theta, bool allkept, int sh, unsigned gmask, uint32_t tagv,
                                                 int lane, int b, int T, int cw, uint32_t pool_s, int &slow_n) {
        const unsigned tbits = (unsigned)(g.t << 3) | kTagValid;
        const int iters = (g.n + 31) >> 5;
        uint32_t ad = g.docs + (lane << 2), as = g.sc + (lane << 2);
        int rem = g.n - lane;
        for (int it = 0; it < iters; it += 4, ad += 512, as += 512, rem -= 128) {
            unsigned d[4], x[4];
            float sv[4];
            bool v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = rem > 32 * u;
                sv[u] = v[u] ? ldsf(as + 128 * u) : 0.0f;
                if (v[u] && !allkept) v[u] = kept_posting(sv[u], g.rest, theta);
                d[u] = v[u] ? lds32(ad + 128 * u) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = v[u] ? lds8(tagv + (d[u] >> sh)) : 0u;
            unsigned act = 0, slow = 0;  // act: a sole hit above theta; slow: needs the slow path
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!v[u]) continue;
                const unsigned mine = tag_of(tbits, d[u], gmask);
                if (x[u] == mine) { if (sv[u] > theta) act |= 1u << u; }
                else if ((x[u] & 0x7Fu) == mine || ((x[u] ^ mine) & gmask) != 0u) slow |= 1u << u;
            }
            if (__ballot_sync(0xffffffffu, (act | slow) != 0u)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if ((act >> u) & 1u) emit(sv[u], (int)d[u]);
                    const unsigned m = __ballot_sync(0xffffffffu, (slow >> u) & 1u);
                    if (m) {
                        const int cnt = __popc(m);
                        if (slow_n + cnt > 32) { resolve_slow(b, T, cw, pool_s, theta, slow_n); slow_n = 0; }
                        if ((slow >> u) & 1u) {
                            const unsigned mine = tag_of(tbits, d[u], gmask);
                            const int rule = ((x[u] & 0x7Fu) == mine) ? 0 : 1;  // 0: contested winner, 1: another document won
                            { const int at = slow_n + __popc(m & ((1u << lane) - 1u)); h->w_slow[cw][at] = make_int2((int)d[u], g.t | (rule << 8)); h->w_slow_s[cw][at] = sv[u]; }
                        }
                        slow_n += cnt;
                        __syncwarp();
                    }
                }
            }
        }
    }

    // P3 over chunks (same logic as pass_resolve, one posting per lane and chunk)
    __device__ __forceinline__ void chunks_resolve(uint32_t chunk_s, int nch, float theta, int sh, unsigned gmask, uint32_t tagv,
                                                   int lane, int b, int T, int cw, uint32_t pool_s, int &slow_n) {
        for (int c0 = 0; c0 < nch; c0 += 4) {
            Chunk4 c;
            load_chunks(chunk_s, c0, nch, c);
            unsigned d[4], x[4];
            float sv[4];
            bool v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = lane < c.cnt[u];
                sv[u] = v[u] ? ldsf(c.sc[u] + 4u * lane) : 0.0f;
                if (v[u] && !c.allkept[u]) v[u] = kept_posting(sv[u], c.rest[u], theta);
                d[u] = v[u] ? lds32(c.docs[u] + 4u * lane) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = v[u] ? lds8(tagv + (d[u] >> sh)) : 0u;
            unsigned act = 0, slow = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!v[u]) continue;
                const unsigned mine = tag_of(c.tbits[u], d[u], gmask);
                if (x[u] == mine) { if (sv[u] > theta) act |= 1u << u; }
                else if ((x[u] & 0x7Fu) == mine || ((x[u] ^ mine) & gmask) != 0u) slow |= 1u << u;
            }
            if (__ballot_sync(0xffffffffu, (act | slow) != 0u)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if ((act >> u) & 1u) emit(sv[u], (int)d[u]);
                    const unsigned m = __ballot_sync(0xffffffffu, (slow >> u) & 1u);
                    if (m) {
                        const int cnt = __popc(m);
                        if (slow_n + cnt > 32) { resolve_slow(b, T, cw, pool_s, theta, slow_n); slow_n = 0; }
                        if ((slow >> u) & 1u) {
                            const unsigned mine = tag_of(c.tbits[u], d[u], gmask);
                            const int rule = ((x[u] & 0x7Fu) == mine) ? 0 : 1;
                            const int at = slow_n + __popc(m & ((1u << lane) - 1u));
                            h->w_slow[cw][at] = make_int2((int)d[u], (int)((c.tbits[u] >> 3) & 15u) | (rule << 8));
                            h->w_slow_s[cw][at] = sv[u];
                        }
                        slow_n += cnt;
                        __syncwarp();
                    }
                }
            }
        }
    }

    // P4: one lane per listed document re-sums it in query-token order (the reference's float32 order) from binary
    // searches of this warp's slice of every token. rule 1: only the essential hit with the smallest token slot emits
    // (several postings of the same document may be listed).
    __device__ void resolve_slow(int b, int T, int cw, uint32_t pool_s, float theta, int n) {
        const int lane = threadIdx.x & 31;
        if (p.debug & 64) return;  // timing experiment: what the slow path costs
        __syncwarp();
        if (lane < n) {
            const int2 e = h->w_slow[cw][lane];
            const int d = e.x, myslot = e.y & 0xFF, rule = e.y >> 8;
            float sum = 0.0f, pend = 0.0f;  // the newest contribution is added one hit later: a score that comes from
            bool has = false;               // global memory (non-essential token) is in flight during the next search
            int min_e = 99;
            const float own = h->w_slow_s[cw][lane];
            for (int t = 0; t < T; ++t) {
                if (t == myslot) {  // the listing posting itself (essential): no search
                    if (has) sum = __fadd_rn(sum, pend);
                    pend = own;
                    has = true;
                    if (t < min_e) min_e = t;
                    continue;
                }
                const int o0 = h->s_lo[cw][t], o1 = h->s_hi[cw][t];
                if (o1 == o0) continue;
                const uint32_t dt = pool_s + 4u * (unsigned)h->w_doc[b][t];
                int lo = o0, hi = o1;
                while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)lds32(dt + 4u * m) < d) lo = m + 1; else hi = m; }
                if (lo < o1 && (int)lds32(dt + 4u * lo) == d) {
                    const bool ne = h->t_ne[b][t] != 0;
                    if (has) sum = __fadd_rn(sum, pend);
                    pend = ne ? p.data[h->w_goff[b][t] + lo] : ldsf(pool_s + 4u * (unsigned)(h->w_sc[b][t] + lo));
                    has = true;
                    if (!ne && t < min_e) min_e = t;
                }
            }
            if (has) sum = __fadd_rn(sum, pend);
            if ((rule == 0 || min_e == myslot) && sum > theta) emit(sum, d);
        }
        __syncwarp();
    }

    // All passes of one warp and window when its essential postings fit 4 * G chunks (the common case): descriptors,
    // document ids, scores and the kept predicate are loaded once and stay in registers from P1 to P5.
    template <int G>
    __device__ __forceinline__ void window_small(uint32_t chunk_s, int nch, float theta, int sh, unsigned gmask, uint32_t tagv,
                                                 int lane, int b, int T, int cw, uint32_t pool_s, unsigned nmask) {
        constexpr int N = 4 * G;
        unsigned d[N], mine[N], slot[N];
        float sv[N];
        bool v[N];
        bool kept_any = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            Chunk4 c;
            load_chunks(chunk_s, 4 * g, nch, c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 4 * g + u;
                v[i] = lane < c.cnt[u];
                sv[i] = v[i] ? ldsf(c.sc[u] + 4u * lane) : 0.0f;
                if (v[i] && !c.allkept[u]) v[i] = kept_posting(sv[i], c.rest[u], theta);
                d[i] = v[i] ? lds32(c.docs[u] + 4u * lane) : 0u;
                mine[i] = tag_of(c.tbits[u], d[i], gmask);
                slot[i] = (c.tbits[u] >> 3) & 15u;
                kept_any |= v[i];
            }
        }
        // P1
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (v[i]) sts8(tagv + (d[i] >> sh), mine[i]);
        const bool warp_kept = __ballot_sync(0xffffffffu, kept_any) != 0u;
        __syncwarp();
        if (!warp_kept) return;
        // P2
        {
            unsigned x[N], bad = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = v[i] ? lds8(tagv + (d[i] >> sh)) : 0u;
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (v[i] && x[i] != mine[i]) bad |= 1u << i;
            if (__ballot_sync(0xffffffffu, bad != 0u)) {
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if ((bad >> i) & 1u) sts8(tagv + (d[i] >> sh), x[i] | kTagContested);
            }
        }
        for (unsigned m = nmask; m; m &= m - 1) {
            const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
            pass_probe_n_vec(g, sh, gmask, tagv, lane, p.data + h->w_goff[b][g.t] + h->s_lo[cw][g.t]);
        }
        __syncwarp();
        // P3
        int slow_n = 0;
        {
            unsigned x[N];
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = v[i] ? lds8(tagv + (d[i] >> sh)) : 0u;
            unsigned act = 0, slow = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (!v[i]) continue;
                if (x[i] == mine[i]) { if (sv[i] > theta) act |= 1u << i; }
                else if ((x[i] & 0x7Fu) == mine[i] || ((x[i] ^ mine[i]) & gmask) != 0u) slow |= 1u << i;
            }
            if (__ballot_sync(0xffffffffu, (act | slow) != 0u)) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    if ((act >> i) & 1u) emit(sv[i], (int)d[i]);
                    const unsigned m = __ballot_sync(0xffffffffu, (slow >> i) & 1u);
                    if (m) {
                        const int cnt = __popc(m);
                        if (slow_n + cnt > 32) { resolve_slow(b, T, cw, pool_s, theta, slow_n); slow_n = 0; }
                        if ((slow >> i) & 1u) {
                            const int rule = ((x[i] & 0x7Fu) == mine[i]) ? 0 : 1;
                            { const int at = slow_n + __popc(m & ((1u << lane) - 1u)); h->w_slow[cw][at] = make_int2((int)d[i], (int)slot[i] | (rule << 8)); h->w_slow_s[cw][at] = sv[i]; }
                        }
                        slow_n += cnt;
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        // P5
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (v[i]) sts8(tagv + (d[i] >> sh), 0u);
        // P4
        if (slow_n > 0) resolve_slow(b, T, cw, pool_s, theta, slow_n);
    }

    // ------------------------------------------------------------------------------------------ consumer warps
    __device__ void consumer(int T, int k, uint32_t &ph_meta0, uint32_t &ph_meta1) {
        const int ctid = threadIdx.x;  // consumers are warps 0..NW-1; the producer is the last warp (the scheduler favours it)
        const int cw = ctid >> 5, lane = ctid & 31;
        const int trigger = p.trigger;
        float theta = h->theta;
        int b = 0;
        bool first_window = true;
        PROF_DECL;
        for (;;) {
            PROF(0);
            if (b == 0) { mbar_wait(&h->bar_meta[0], ph_meta0); ph_meta0 ^= 1; }
            else { mbar_wait(&h->bar_meta[1], ph_meta1); ph_meta1 ^= 1; }
            PROF(1);
            PROF_CNT(16, 1);
            const int *bp = pool + (size_t)b * p.poolcap;
            const float *bpf = reinterpret_cast<const float *>(bp);
            const uint32_t pool_s = smem_u32(bp);
            const int sh = h->w_shift[b];
            const unsigned gmask = (1u << sh) - 1u;
            const uint32_t tagv = smem_u32(tag) - (uint32_t)(h->w_base[b] >> sh);  // tag of document d: tagv + (d >> sh)
            const int last = h->w_last[b];

            if (first_window) {
                // Seed theta: the k-th largest score among the staged postings of the token with the most postings in
                // this window is reached by >= k distinct documents (scores are non-negative), so anything strictly
                // below it cannot be in the top-k. Avoids flooding the candidate buffer while theta is still 0.
                first_window = false;
                int best_t = 0, best_c = 0;
                for (int t = 0; t < T; ++t) {
                    const int c = h->t_ne[b][t] ? 0 : h->w_cnt[b][t];
                    if (c > best_c) { best_c = c; best_t = t; }
                }
                const int m = min(best_c, p.cbcap);
                if (m >= k && p.a.mask == nullptr && p.a.mask_bits == nullptr && theta == 0.0f) {  // only when the k-th-largest table gave no start value
                    const int P = next_pow2(m);
                    const float *sc = bpf + h->w_sc[b][best_t];
                    for (int i = ctid; i < P; i += CNT) cb[i] = (i < m) ? ((unsigned long long)__float_as_uint(sc[i]) << 32) : 0ull;
                    sort_desc<true>(cb, P, ctid, CNT);
                    const unsigned int bits = (unsigned int)(cb[k - 1] >> 32);
                    theta = fmaxf(theta, bits > 0u ? __uint_as_float(bits - 1u) : 0.0f);  // next float below the k-th sampled score
                    bsync<true>();
                    if (ctid == 0) h->theta = theta;
                }
            }
            PROF(2);
            PROF_CNT(17, h->w_ME[b]); PROF_CNT(18, h->w_MN[b]);

            // ---- this warp's document sub-range [bnd(cw), bnd(cw + 1)), boundaries on multiples of G: lane t finds the
            //      first posting of token t at or after the lower bound, lane 16 + t does the same for the upper bound
            unsigned emask, nmask;  // tokens with a non-empty essential / non-essential slice (warp-uniform)
            {
                const int wbase = h->w_base[b], wendd = h->w_end[b];
                const unsigned wspan = (unsigned)(wendd - wbase);
                const unsigned step = ((wspan / (unsigned)NW) + gmask + 1u) & ~gmask;  // multiple of G; NW steps cover the window
                const int t = lane & 15, side = lane >> 4;
                const int wq = cw + side;
                const long long bl = (long long)wbase + (long long)step * wq;
                const int bnd = (wq == NW || bl > wendd) ? wendd : (int)bl;
                int pos = 0, ne = 0;
                if (t < T) {
                    const int c = h->w_cnt[b][t];
                    const uint32_t dt = pool_s + 4u * (unsigned)h->w_doc[b][t];
                    ne = h->t_ne[b][t];
                    int lo = 0, hi = c;
                    if (wq == 0) hi = 0;
                    else if (wq == NW) lo = c;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)lds32(dt + 4u * m) < bnd) lo = m + 1; else hi = m; }
                    pos = lo;
                    if (side == 0) h->s_lo[cw][t] = lo; else h->s_hi[cw][t] = lo;
                }
                const int upper = __shfl_down_sync(0xffffffffu, pos, 16);
                const bool nonempty = (side == 0) && (t < T) && (upper > pos);
                emask = __ballot_sync(0xffffffffu, nonempty && !ne);
                nmask = __ballot_sync(0xffffffffu, nonempty && ne);
                __syncwarp();
            }
            // ---- chunk list of this warp's essential postings (<= 32 consecutive postings of one token per chunk)
            const uint32_t chunk_s = smem_u32(chunks + (size_t)cw * kMaxChunk);
            int nch = 0;
            bool chunk_ok = true;
            for (unsigned m = emask; m; m &= m - 1) {
                const int t = __ffs(m) - 1;
                const int o0 = h->s_lo[cw][t], n = h->s_hi[cw][t] - o0;
                const int nc = (n + 31) >> 5;
                if (nch + nc > kMaxChunk) { chunk_ok = false; break; }
                const float rest = h->t_rest[t];
                const unsigned ak = kept_posting(0.0f, rest, theta) ? 1u : 0u;
                const uint32_t dbase = pool_s + 4u * (unsigned)(h->w_doc[b][t] + o0), sbase = pool_s + 4u * (unsigned)(h->w_sc[b][t] + o0);
                for (int c = lane; c < nc; c += 32) {
                    uint4 dsc;
                    dsc.x = dbase + 128u * c;
                    dsc.y = sbase + 128u * c;
                    dsc.z = (unsigned)min(32, n - 32 * c) | ((unsigned)t << 8) | (ak << 16);
                    dsc.w = __float_as_uint(rest);
                    chunks[(size_t)cw * kMaxChunk + nch + c] = dsc;
                }
                nch += nc;
            }
            __syncwarp();
            if (chunk_ok && nch <= 4 && !(p.debug & 63)) {
                window_small<1>(chunk_s, nch, theta, sh, gmask, tagv, lane, b, T, cw, pool_s, nmask);
            } else if (chunk_ok && nch <= 8 && !(p.debug & 63)) {
                window_small<2>(chunk_s, nch, theta, sh, gmask, tagv, lane, b, T, cw, pool_s, nmask);
            } else {
                // ---- P1: essential postings that can still reach the threshold claim their document's group
                bool kept_any = false;
                if (p.debug & 16) {
                } else if (chunk_ok) {
                    kept_any = chunks_claim(chunk_s, nch, theta, sh, gmask, tagv, lane);
                } else {
                    for (unsigned m = emask; m; m &= m - 1) {
                        const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
                        kept_any |= pass_claim(g, theta, kept_posting(0.0f, g.rest, theta), sh, gmask, tagv, lane);
                    }
                }
                const bool warp_kept = __ballot_sync(0xffffffffu, kept_any) != 0u;
                PROF(3);
                __syncwarp();
                PROF_CNT(19, warp_kept ? 1 : 0);
                if (warp_kept) {  // no surviving essential posting: nothing in this sub-range can score
                    // ---- P2: postings that do not read back their own claim set the contested bit; non-essential postings
                    //      that find their own document claimed do the same
                    if (p.debug & 2) {
                    } else if (chunk_ok) {
                        chunks_mark(chunk_s, nch, theta, sh, gmask, tagv, lane);
                    } else {
                        for (unsigned m = emask; m; m &= m - 1) {
                            const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
                            pass_mark_e(g, theta, kept_posting(0.0f, g.rest, theta), sh, gmask, tagv, lane);
                        }
                    }
                    if (!(p.debug & 1))
                    for (unsigned m = nmask; m; m &= m - 1) {
                        const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
                        pass_probe_n_vec(g, sh, gmask, tagv, lane, p.data + h->w_goff[b][g.t] + h->s_lo[cw][g.t]);
                    }
                    PROF(5);
                    __syncwarp();
                    // ---- P3: sole hits score directly; contested winners and documents without a winner go to the slow list
                    int slow_n = 0;
                    if (p.debug & 4) {
                    } else if (chunk_ok) {
                        chunks_resolve(chunk_s, nch, theta, sh, gmask, tagv, lane, b, T, cw, pool_s, slow_n);
                    } else {
                        for (unsigned m = emask; m; m &= m - 1) {
                            const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
                            pass_resolve(g, theta, kept_posting(0.0f, g.rest, theta), sh, gmask, tagv, lane, b, T, cw, pool_s, slow_n);
                        }
                    }
                    PROF(7);
                    __syncwarp();
                    // ---- P5: the claiming postings clean their tags (tags are all-clean between windows)
                    if (p.debug & 8) {
                    } else if (chunk_ok) {
                        chunks_clean(chunk_s, nch, theta, sh, tagv, lane);
                    } else {
                        for (unsigned m = emask; m; m &= m - 1) {
                            const Seg g = slice_of(b, __ffs(m) - 1, cw, pool_s);
                            pass_clean(g, theta, kept_posting(0.0f, g.rest, theta), sh, tagv, lane);
                        }
                    }
                    PROF(9);
                    // ---- P4: the rest of the slow list
                    PROF_CNT(21, slow_n);
                    if (slow_n > 0) resolve_slow(b, T, cw, pool_s, theta, slow_n);
                }
            }
            PROF(10);
            // the consumers meet once per window: stage b and its tags are released, the candidate buffer is checked
            bsync<true>();
            PROF(4);
            if (ctid == 0) mbar_arrive(&h->bar_free[b]);
            if (h->cb_cnt >= (unsigned)trigger) { compact<true>(k, ctid, CNT); theta = fmaxf(theta, h->theta); PROF_CNT(22, 1); }
            PROF(11);
            if (last) break;
            if (p.stages == 2) b ^= 1;
        }
        PROF_FLUSH(ctid == 32 * (p.debug >> 8));
    }

    __device__ void run() {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int k = p.a.k;
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(&h->bar_data[i], 1); mbar_init(&h->bar_meta[i], 1); mbar_init(&h->bar_free[i], 1); }
            fence_mbar_init();
        }
        for (int i = tid * 16; i < p.tagcap; i += NT * 16) *reinterpret_cast<uint4 *>(tag + i) = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        uint32_t ph_data0 = 0, ph_data1 = 0, ph_free0 = 0, ph_free1 = 0, ph_meta0 = 0, ph_meta1 = 0;
        PROF_DECL;
        for (;;) {
            PROF(12);
            if (tid == 0) h->q_slot = atomicAdd(&p.ctrl->fast_next, 1u);
            __syncthreads();
            const unsigned int slot = h->q_slot;
            if (slot >= (p.qcount ? *p.qcount : (unsigned)p.a.n_queries)) break;
            const int q = p.order ? (int)p.order[slot] : (int)slot;
            const long long qs = p.a.q_ptr[q], qe = p.a.q_ptr[q + 1];
            const long long Tl = qe - qs;
            if (p.route_all_general || Tl > kMaxTok || (p.a.mask && p.ctrl->mask_bad)) {
                if (tid == 0) p.general_list[atomicAdd(&p.ctrl->general_count, 1u)] = (unsigned)q;
                __syncthreads();
                continue;
            }
            const int T = (int)Tl;
            if (tid < T) {
                const int tok = p.a.q_tokens[qs + tid];
                long long beg = 0, end = 0;
                float mx = 0.0f, kth = 0.0f;
                if ((unsigned)tok >= (unsigned)p.n_vocab) atomicOr(&p.ctrl->error, 1u);
                else {
                    beg = p.indptr[tok]; end = p.indptr[tok + 1]; mx = p.colmax[tok];
                    if (p.colkth) kth = p.colkth[tok];
                }
                h->t_cur[tid] = beg;
                h->t_end[tid] = end;
                h->t_df[tid] = end - beg;
                h->t_max[tid] = mx;
                h->t_kth[tid] = kth;
            }
            if (tid == 0) { h->cb_cnt = 0; h->theta = (p.theta_mode == 2) ? p.theta_io[q] : 0.0f; }
            __syncthreads();
            Layout lay;
            lay.cap = 0; lay.off = 0; lay.ne = -1; lay.live = -1; lay.sh = 0;  // forces the first share-out
            if (warp == NW) {
                // pruning bounds from the column maxima
                const long long df0 = (lane < T) ? h->t_df[lane] : 0;
                long long tot = df0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
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
                    // float32 sums of <= 15 non-negative terms exceed the exact sum by < 2^-19 relative: pad by 2^-15
                    h->t_rest[lane] = __double2float_ru((msum - (double)mx) * 1.00003);
                    h->t_nebound[lane] = __double2float_ru(below * 1.00003);
                }
                // Start threshold: the k-th largest posting of one token is reached by k distinct documents, and a
                // document scores at least each of its postings (non-negative terms) - the next float below it is safe.
                {
                    const unsigned kb = __reduce_max_sync(0xffffffffu, (lane < T) ? __float_as_uint(h->t_kth[lane]) : 0u);
                    if (lane == 0 && kb > 0u) h->theta = fmaxf(h->theta, __uint_as_float(kb - 1u));
                }
                if (lane == 0) {
                    h->total_df = (unsigned long long)tot;
                    if (tot > 0) atomicAdd(&p.ctrl->posting_count, (unsigned long long)tot);
                }
                __syncwarp();
                if (tot > 0) issue_loads(0, T, lay);
            }
            __syncthreads();
            PROF(13);
            if (h->total_df > 0) {
                if (warp == NW) producer(T, lay, ph_data0, ph_data1, ph_free0, ph_free1);
                else consumer(T, k, ph_meta0, ph_meta1);
            }
            __syncthreads();
            PROF(14);
            finalize(q, T);
            PROF(15);
        }
        PROF_FLUSH(tid == 0);
    }

    // Sort the surviving candidates, pad with untouched documents if fewer than k matched, write the result row.
    __device__ void finalize(int q, int T) {
        const int tid = threadIdx.x;
        const int k = p.a.k;
        compact<false>(k, tid, NT, true);  // leaves cb[0..n) sorted descending, n <= k
        const int n = (int)h->cb_cnt;
        if (p.theta_mode == 1 && tid == 0) {
            const unsigned int bits = (n == k) ? (unsigned int)(cb[k - 1] >> 32) : 0u;
            p.theta_io[q] = bits > 0u ? __uint_as_float(bits - 1u) : 0.0f;
        }
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
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) retrieve_fast_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Fast<NT> f(p, smem);
    f.run();
}

int launch_retrieve_fast(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_general_list,
                         int *launches, const unsigned int *qlist, const unsigned int *qcount) {
    FastParams p;
    p.data = ix->d_data;
    p.indices = ix->d_indices;
    p.indptr = ix->d_indptr;
    p.colmax = ix->d_colmax;
    p.order = qlist ? qlist : (ix->order_valid ? ix->d_order.as<unsigned int>() : nullptr);
    p.qcount = qlist ? qcount : nullptr;
    // the k-th-largest table holds ranks 10 / 100 / 1000: rank r bounds every k <= r (k = 1 uses the column maxima)
    p.colkth = nullptr;
    if (ix->knobs.get("SEED_KTH", 1) && ix->d_colkth && a.mask == nullptr && a.mask_bits == nullptr) {  // a mask may remove the documents behind the bound
        if (a.k == 1) p.colkth = ix->d_colmax;
        else if (a.k <= 10) p.colkth = ix->d_colkth;
        else if (a.k <= 100) p.colkth = ix->d_colkth + (size_t)ix->n_vocab;
        else if (a.k <= 1000) p.colkth = ix->d_colkth + 2 * (size_t)ix->n_vocab;
    }
    p.prune = ix->knobs.get("PRUNE", 1);
    p.max_shift = ix->knobs.get("MAXSHIFT", 2);
    if (p.max_shift < 0) p.max_shift = 0;
    if (p.max_shift > 2) p.max_shift = 2;
    p.n_docs = ix->n_docs;
    p.n_vocab = ix->n_vocab;
    p.doc_base = ix->doc_base;
    p.a = a;
    p.ctrl = d_ctrl;
    p.general_list = d_general_list;
    int threads = ix->knobs.get("THREADS", 512);
    if (threads != 128 && threads != 256 && threads != 1024) threads = 512;
    const int ctas_per_sm = ix->knobs.get("CTAS_PER_SM", 1);
    const int kcap = next_pow2(a.k < 1 ? 1 : a.k);
    const int cb_min = ix->knobs.get("CBMIN", 32);  // candidate buffer: 2 * next_pow2(k) entries unless raised (small buffers compact often and keep theta fresh)
    // 2 * next_pow2(k) keys for small k; four times for larger k (fewer cuts), never more than 4096 unless k needs it
    int cb_auto = kcap <= 64 ? 2 * kcap : 4 * kcap;
    if (cb_auto > 4096) cb_auto = 2 * kcap > 4096 ? 2 * kcap : 4096;
    p.cbcap = cb_auto < cb_min ? next_pow2(cb_min) : cb_auto;
    p.route_all_general = (!ix->nonneg) || (a.k > 2048) || ix->knobs.get("FORCE_GENERAL", 0) ||
                          ((a.mask != nullptr || a.mask_bits != nullptr) && ix->knobs.get("MASK_FAST", 1) == 0);
    const int budget = ix->max_smem_optin / (ctas_per_sm > 0 ? ctas_per_sm : 1) - (ctas_per_sm > 1 ? 1024 : 0);
    const int nw = threads / 32 - 1;
    p.nwarps = nw;
    p.debug = ix->knobs.get("DEBUG_SKIP", 0);
    p.theta_mode = ix->knobs.get("DEBUG_THETA", 0);
    p.theta_io = nullptr;
    if (p.theta_mode) {
        int rc0 = ix->d_theta.reserve((size_t)(a.n_queries > 0 ? a.n_queries : 1) * sizeof(float));
        if (rc0) return rc0;
        p.theta_io = ix->d_theta.as<float>();
    }
    // pool: 4-byte slots per stage
    p.stages = ix->knobs.get("STAGES", 2) == 1 ? 1 : 2;
    p.trigger = 0;
    int pool = ix->knobs.get("POOL", ctas_per_sm > 1 ? 6144 : (p.stages == 1 ? 24576 : 14336));
    pool = (pool + 3) & ~3;
    const int min_pool = kMaxTok * 2 * kMinCap + 256;
    if (pool < min_pool) pool = min_pool;
    (void)nw;
    if (pool > 40960) pool = 40960;
    p.poolcap = pool;
    size_t fixed = Fast<512>::fixed_smem(p);
    long long tagcap = (long long)budget - (long long)fixed - 64;
    if (tagcap < 4096) {
        p.route_all_general = 1;  // k too large for this shared-memory budget
        p.cbcap = 512;
        fixed = Fast<512>::fixed_smem(p);
        tagcap = (long long)budget - (long long)fixed - 64;
    }
    const int tag_env = ix->knobs.get("TAG", 0);
    if (tag_env > 0 && tag_env < tagcap) tagcap = tag_env;
    p.tagcap = (int)(tagcap & ~15ll);
    p.trigger = ix->knobs.get("TRIGGER", p.cbcap / 2);  // candidates that start a cut of the buffer at a window end
    if (p.trigger <= a.k) p.trigger = a.k + 1;
    if (p.trigger > p.cbcap) p.trigger = p.cbcap;
    p.spillcap = (p.stages == 1 ? 1 : 2) * p.poolcap;
    const int grid = ix->sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    int rc = ix->d_spill.reserve((size_t)grid * (size_t)p.spillcap * 8);
    if (rc) return rc;
    p.spill = ix->d_spill.as<unsigned long long>();
    const size_t smem = fixed + p.tagcap;
    cudaError_t e;
#define BM25CU_LAUNCH(NTV)                                                                                          \
    e = cudaFuncSetAttribute(retrieve_fast_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e == cudaSuccess) retrieve_fast_kernel<NTV><<<grid, NTV, smem, ix->stream>>>(p);
    switch (threads) {
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
