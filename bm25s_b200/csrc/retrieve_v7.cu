// K1, alternative structure (BM25CU_KERNEL=7, experimental): bit tags and posting-index splits on top of the
// posting pool of retrieve_fast.cu.
//
// The producer warp stages, per window and double-buffered, the document ids of EVERY live token and the scores of
// the essential ones (bulk async copies); a window is what the pool covers for all tokens, capped by the tag span
// (2 bits per document: PRESENT and MULTI). The consumer warps split the window's essential postings and its
// non-essential postings by POSTING INDEX (one or two segments per warp whatever the number of query tokens - no
// per-warp document sub-ranges, no boundary searches) and meet at named barriers between the phases:
//   A   surviving essential postings set PRESENT with an atomic OR; a second hit of the same document sets MULTI
//   C   non-essential postings (read from the pool, 16 bytes per lane) whose document is PRESENT set MULTI and file
//       (doc, token, posting index) in a small hash table
//   D1  (the hash entries' scores arrive by asynchronous 4-byte copies issued at hit time); essential
//       postings read their document's bits: not MULTI -> the posting's score is the document's score (0 + s = s
//       exactly); MULTI -> the CTA's slow list
//   D2  slow list, one thread per entry: the document is re-summed in query-token order (the reference's float32
//       order) from binary searches of the staged essential postings and hash look-ups of the non-essential ones;
//       among a document's surviving essential postings the one with the smallest token slot emits. The tags of the
//       window are cleared.
// Exactness, pruning bounds and candidate handling are those of retrieve_fast.cu.
#include <cstdlib>

#include "fast_common.cuh"

namespace bm25cu {
namespace v7 {

constexpr int kHash = 1024;     // entries of the per-window hash of non-essential contributions (power of two)
constexpr int kHashMax = 768;   // inserts beyond this switch the window to binary searches in global memory
constexpr int kSlowCap = 1024;  // entries of the CTA's slow list (more: resolved on the spot)

struct Hdr {
    uint64_t bar_data[2];   // stage b landed                        (producer waits)
    uint64_t bar_meta[2];   // window metadata of stage b published  (consumers wait)
    uint64_t bar_free[2];   // consumers are done with stage b       (producer waits)
    long long t_beg[kTokArr];
    long long t_cur[kTokArr];
    long long t_end[kTokArr];
    long long t_df[kTokArr];
    long long w_goff[2][kTokArr];  // global posting index of the token's first posting in the window
    float t_max[kTokArr];     // largest score of the token's column
    float t_kth[kTokArr];     // k-th largest score of the token's column (0: unknown / fewer than k postings)
    float t_rest[kTokArr];    // upper bound of what the OTHER query tokens can add to a document
    float t_nebound[kTokArr]; // upper bound of a document hit only by this token and tokens with smaller maxima
    int t_n[2][kTokArr];      // staged postings of stage b (essential tokens)
    int t_ne[2][kTokArr];     // 1: token is non-essential in stage b
    int w_cnt[2][kTokArr];    // postings of the token inside the window
    int w_doc[2][kTokArr];    // pool slot of the first window document id (essential tokens)
    int w_sc[2][kTokArr];     // pool slot of the first window score
    int2 slow[kSlowCap];      // slow list (doc, token slot)
    int hk[kHash];            // hash: document (-1 empty)
    int ht[kHash];            //       token slot
    int hi[kHash];            //       posting index relative to the column start
    float hv[kHash];          //       score (fetched in D1)
    int w_base[2];            // first document of the window, a multiple of 4
    int w_end[2];             // one past the last document of the window
    int w_last[2];            // 1 if this is the query's last window
    int w_ME[2];              // essential postings in the window
    int w_MN[2];              // non-essential postings in the window
    unsigned int w_emask[2];  // tokens with essential postings in the window
    unsigned int w_nmask[2];  // tokens with non-essential postings in the window
    unsigned int h_cnt_w[kMaxWarps];  // hash inserts per consumer warp (a warp may file hash_max / NW entries per window)
    unsigned int h_any, h_over, slow_cnt;
    unsigned int sel_hist[256];  // radix-select histogram of the candidate cut (large k)
    unsigned int sel_bin, sel_need, sel_cnt;
    unsigned int q_slot;
    unsigned int cb_cnt;
    float theta;
    unsigned long long total_df;
};

__device__ __forceinline__ unsigned atoms_or(uint32_t a, unsigned v) {
    unsigned o;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ void reds_or(uint32_t a, unsigned v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128z(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }

// Tag of document d: byte at tagv + (d >> 2) (tagv already has the window base folded in); bit (d & 3) is PRESENT,
// bit 4 + (d & 3) is MULTI. Atomics work on the enclosing 32-bit word.
__device__ __forceinline__ uint32_t tag_addr(uint32_t tagv, unsigned d) { return tagv + (d >> 2); }
__device__ __forceinline__ unsigned tag_bit(uint32_t a, unsigned d) { return ((a & 3u) << 3) + (d & 3u); }

struct Seg {
    uint32_t docs;  // shared address of the first document id of the segment
    uint32_t sc;    // shared address of the first score
    float rest;
    int t;
    int n;
};

// A: surviving essential postings mark their document PRESENT; a second hit of the same document marks it MULTI
__device__ __forceinline__ void pass_claim(const Seg &g, float theta, bool allkept, uint32_t tagv, int lane) {
    const int iters = (g.n + 31) >> 5;
    uint32_t ad = g.docs + (lane << 2), as = g.sc + (lane << 2);
    int rem = g.n - lane;  // > 32 * u  <=>  element (it + u) * 32 + lane exists
    for (int it = 0; it < iters; it += 4, ad += 512, as += 512, rem -= 128) {
        float sv[4];
        unsigned d[4], old[4], bit[4];
        uint32_t aw[4];
        bool kp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool v = rem > 32 * u;
            sv[u] = (v && !allkept) ? ldsf(as + 128 * u) : 0.0f;
            d[u] = v ? lds32(ad + 128 * u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // all atomics are issued before any result is looked at
            kp[u] = rem > 32 * u && (allkept || kept_posting(sv[u], g.rest, theta));
            const uint32_t a = tag_addr(tagv, d[u]);
            bit[u] = tag_bit(a, d[u]);
            aw[u] = a & ~3u;
            old[u] = kp[u] ? atoms_or(aw[u], 1u << bit[u]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (kp[u] && ((old[u] >> bit[u]) & 1u)) reds_or(aw[u], 16u << bit[u]);
    }
}

// per-posting tag clean-up (used when the window holds few essential postings; otherwise the tag range is zeroed)
__device__ __forceinline__ void pass_clean(const Seg &g, float theta, bool allkept, uint32_t tagv, int lane) {
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
            if (v[u]) sts8(tag_addr(tagv, d[u]), 0u);
    }
}

template <int NT>
struct Fast {
    static constexpr int CNT = NT - 32;  // consumer threads
    static constexpr int NW = CNT / 32;  // consumer warps
    Hdr *h;
    unsigned long long *cb;   // [cbcap]
    int *pool;                // [2][poolcap] 4-byte slots: document ids and float scores of the essential tokens
    unsigned char *tag;       // [tagcap]
    const FastParams &p;
    unsigned long long *spill;
    bool hash_dirty = false;  // the hash holds entries of the previous window (cleared in the next claim phase)

    __device__ Fast(const FastParams &pp, unsigned char *smem) : p(pp) {
        h = reinterpret_cast<Hdr *>(smem);
        size_t o = (sizeof(Hdr) + 15) & ~size_t(15);
        cb = reinterpret_cast<unsigned long long *>(smem + o);
        o += (size_t)p.cbcap * 8;
        o = (o + 15) & ~size_t(15);
        pool = reinterpret_cast<int *>(smem + o);
        o += (size_t)2 * p.poolcap * 4;
        tag = smem + o;
        spill = p.spill + (size_t)blockIdx.x * p.spillcap;
    }
    static size_t fixed_smem(const FastParams &p) {
        size_t o = (sizeof(Hdr) + 15) & ~size_t(15);
        o += (size_t)p.cbcap * 8;
        o = (o + 15) & ~size_t(15);
        o += (size_t)2 * p.poolcap * 4;
        return o;
    }

    template <bool CONS>
    __device__ __forceinline__ void bsync() {
        if (CONS) asm volatile("bar.sync 1, %0;" ::"n"(CNT) : "memory");
        else __syncthreads();
    }

    __device__ __forceinline__ void emit(float s, int d) {
        if (p.a.mask) {  // weight mask in [0, 1]: numpy multiplies in the promoted type and rounds back to float32
            s = (float)((double)s * p.a.mask[d]);
            if (!(s > *reinterpret_cast<volatile float *>(&h->theta))) return;
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
    // share-out only changes when a token runs dry or turns non-essential, so it is cached in the producer lanes'
    // registers (lane == token slot).
    struct Layout {
        int cap, off, ne, live;
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
            long long w = live ? h->t_df[lane] * (ne ? 1 : 2) : 0;
            long long wsum = w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
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
        const long long span = (long long)p.tagcap << 2;
        PROF_DECL;
        for (;;) {
            PROF(24);
            if (b == 0) { mbar_wait(&h->bar_data[0], ph_data0); ph_data0 ^= 1; }
            else { mbar_wait(&h->bar_data[1], ph_data1); ph_data1 ^= 1; }
            PROF(25);
            const int *bp = pool + (size_t)b * p.poolcap;
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
            const int base = first & ~3;
            const int cap_end = ((long long)base + span > 0x7fffffffll) ? 0x7fffffff : (int)(base + span);
            int wend = min(e, cap_end);
            if (lastmax < wend - 1) wend = lastmax + 1;
            int c0 = (n0 > 0 && last0 < wend) ? n0 : 0;
            // tokens whose staged postings reach beyond the window end need a search
            unsigned int need0 = __ballot_sync(0xffffffffu, n0 > 0 && last0 >= wend);
            while (need0) {
                const int src = __ffs(need0) - 1;
                need0 &= need0 - 1;
                const int so = __shfl_sync(0xffffffffu, off0, src), sn = __shfl_sync(0xffffffffu, n0, src);
                const int c = warp_count_below(bp + so, sn, wend);
                if (lane == src) c0 = c;
            }
            const int ME = __reduce_add_sync(0xffffffffu, ne ? 0 : c0);
            const int MN = __reduce_add_sync(0xffffffffu, ne ? c0 : 0);
            const unsigned emask = __ballot_sync(0xffffffffu, !ne && c0 > 0);
            const unsigned nmask = __ballot_sync(0xffffffffu, ne && c0 > 0);
            PROF(26);
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
                h->w_emask[b] = emask;
                h->w_nmask[b] = nmask;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h->bar_meta[b]);  // release: metadata visible to the consumers
            ++nwin;
            if (!any_more) break;
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

    // Calls f(Seg) for every token segment of the flat essential-posting range [lo, hi) of stage b
    template <class F>
    __device__ __forceinline__ void for_segs(int b, unsigned emask, int lo, int hi, uint32_t pool_s, F f) {
        int pre = 0;
        for (unsigned m = emask; m; m &= m - 1) {
            const int t = __ffs(m) - 1;
            const int c = h->w_cnt[b][t];
            const int s0 = max(lo, pre), s1 = min(hi, pre + c);
            if (s1 > s0) {
                Seg g;
                g.docs = pool_s + 4u * (unsigned)(h->w_doc[b][t] + s0 - pre);
                g.sc = pool_s + 4u * (unsigned)(h->w_sc[b][t] + s0 - pre);
                g.rest = h->t_rest[t];
                g.t = t;
                g.n = s1 - s0;
                f(g);
            }
            pre += c;
            if (pre >= hi) break;
        }
    }
    // Calls f(token, offset inside the token's window slice, count) for every column segment of the flat non-essential
    // range [lo, hi)
    template <class F>
    __device__ __forceinline__ void for_nsegs(int b, unsigned nmask, int lo, int hi, F f) {
        int pre = 0;
        for (unsigned m = nmask; m; m &= m - 1) {
            const int t = __ffs(m) - 1;
            const int c = h->w_cnt[b][t];
            const int s0 = max(lo, pre), s1 = min(hi, pre + c);
            if (s1 > s0) f(t, s0 - pre, s1 - s0);
            pre += c;
            if (pre >= hi) break;
        }
    }

    // C, one hit: the document is PRESENT and a non-essential posting contributes to it. The entry's score is copied
    // from global memory straight into the table by an asynchronous 4-byte copy (waited for at the end of D1).
    __device__ __forceinline__ void ne_hit(unsigned d, uint32_t a, int t, long long gidx, long long beg) {
        reds_or(a & ~3u, 16u << tag_bit(a, d));
        const int cw = threadIdx.x >> 5;
        const unsigned slot = atomicAdd(&h->h_cnt_w[cw], 1u);
        if (slot == 0u) h->h_any = 1u;
        if (slot >= (unsigned)p.hash_max) { h->h_over = 1u; return; }
        unsigned i = (d * 2654435761u) >> 22;
        while (atomicCAS(&h->hk[i], -1, (int)d) != -1) i = (i + 1u) & (unsigned)(kHash - 1);
        h->ht[i] = t;
        h->hi[i] = (int)(gidx - beg);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&h->hv[i])), "l"(p.data + gidx) : "memory");
    }

    // C: probes four document ids (one 16-byte shared load); gi = global posting index of the first
    __device__ __forceinline__ void ne_probe4(const uint4 dv, int t, long long beg, long long gi, uint32_t tagv) {
        const unsigned d[4] = {dv.x, dv.y, dv.z, dv.w};
        unsigned x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = lds8(tag_addr(tagv, d[u]));
        unsigned hit = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) hit |= ((x[u] >> (d[u] & 3u)) & 1u) << u;
        if (hit) {  // rare
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((hit >> u) & 1u) ne_hit(d[u], tag_addr(tagv, d[u]), t, gi + u, beg);
        }
    }
    // C: one column segment of the pool: n document ids at shared address `docs`, the first one is posting gi0
    __device__ __forceinline__ void ne_probe_seg(int t, long long beg, uint32_t docs, int n, long long gi0, uint32_t tagv, int lane) {
        const int head = min(n, (int)((4u - ((docs >> 2) & 3u)) & 3u));  // scalar elements up to the first 16-byte boundary
        if (lane < head) {
            const unsigned d = lds32(docs + 4u * lane);
            const unsigned x = lds8(tag_addr(tagv, d));
            if ((x >> (d & 3u)) & 1u) ne_hit(d, tag_addr(tagv, d), t, gi0 + lane, beg);
        }
        const int nv = (n - head) >> 2;  // full vectors
        const uint32_t vb = docs + 4u * head;
        const long long gv = gi0 + head;
        int i = lane;
        for (; i + 32 < nv; i += 64) {  // two vectors per lane in flight
            const uint4 a = lds128(vb + 16u * i), b2 = lds128(vb + 16u * (i + 32));
            ne_probe4(a, t, beg, gv + 4 * i, tagv);
            ne_probe4(b2, t, beg, gv + 4 * (i + 32), tagv);
        }
        if (i < nv) ne_probe4(lds128(vb + 16u * i), t, beg, gv + 4 * i, tagv);
        const int tail0 = head + (nv << 2);
        if (lane < n - tail0) {
            const unsigned d = lds32(docs + 4u * (tail0 + lane));
            const unsigned x = lds8(tag_addr(tagv, d));
            if ((x >> (d & 3u)) & 1u) ne_hit(d, tag_addr(tagv, d), t, gi0 + tail0 + lane, beg);
        }
    }

    // Re-sums one document in query-token order. use_hv: the hash entries' scores are in shared memory (D2),
    // otherwise they are read from global memory.
    __device__ __forceinline__ void resolve_one(int b, int T, uint32_t pool_s, float theta, int d, int myslot, bool use_hv) {
        const bool over = h->h_over != 0u;
        float sum = 0.0f;
        int min_e = 99;
        for (int t = 0; t < T; ++t) {
            const int c = h->w_cnt[b][t];
            if (c == 0) continue;
            if (h->t_ne[b][t]) {
                if (!over) {
                    unsigned i = ((unsigned)d * 2654435761u) >> 22;
                    for (;;) {
                        const int kk = h->hk[i];
                        if (kk == -1) break;
                        if (kk == d && h->ht[i] == t) {
                            sum = __fadd_rn(sum, use_hv ? h->hv[i] : p.data[h->t_beg[t] + h->hi[i]]);
                            break;
                        }
                        i = (i + 1u) & (unsigned)(kHash - 1);
                    }
                } else {  // hash overflow in this window: search the column's window slice in global memory
                    const long long g0 = h->w_goff[b][t];
                    long long lo = g0, hi = g0 + c;
                    while (lo < hi) { const long long m = (lo + hi) >> 1; if (p.indices[m] < d) lo = m + 1; else hi = m; }
                    if (lo < g0 + c && p.indices[lo] == d) sum = __fadd_rn(sum, p.data[lo]);
                }
            } else {
                const uint32_t dt = pool_s + 4u * (unsigned)h->w_doc[b][t];
                int lo = 0, hi = c;
                while (lo < hi) { const int m = (lo + hi) >> 1; if ((int)lds32(dt + 4u * m) < d) lo = m + 1; else hi = m; }
                if (lo < c && (int)lds32(dt + 4u * lo) == d) {
                    const float v = ldsf(pool_s + 4u * (unsigned)(h->w_sc[b][t] + lo));
                    sum = __fadd_rn(sum, v);
                    if (t < min_e && kept_posting(v, h->t_rest[t], theta)) min_e = t;
                }
            }
        }
        if (min_e == myslot && sum > theta) emit(sum, d);
    }

    // D1 on one segment: documents that are not MULTI score directly; the others go to the CTA's slow list
    __device__ __forceinline__ void pass_resolve(const Seg &g, float theta, bool allkept, uint32_t tagv, int lane, int b, int T,
                                                 uint32_t pool_s) {
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
            for (int u = 0; u < 4; ++u) x[u] = v[u] ? lds8(tag_addr(tagv, d[u])) : 0u;
            unsigned act = 0, slow = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!v[u]) continue;
                if ((x[u] >> (4u + (d[u] & 3u))) & 1u) slow |= 1u << u;
                else if (sv[u] > theta) act |= 1u << u;
            }
            if (__ballot_sync(0xffffffffu, (act | slow) != 0u)) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if ((act >> u) & 1u) emit(sv[u], (int)d[u]);
                // slow entries of the whole group: one atomic per warp
                const int mine = __popc(slow);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total) {
                    unsigned base0 = 0;
                    if (lane == 31) base0 = atomicAdd(&h->slow_cnt, (unsigned)total);
                    base0 = __shfl_sync(0xffffffffu, base0, 31);
                    unsigned at = base0 + (unsigned)(incl - mine);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if ((slow >> u) & 1u) {
                            if (at < (unsigned)kSlowCap) h->slow[at] = make_int2((int)d[u], g.t);
                            else resolve_one(b, T, pool_s, theta, (int)d[u], g.t, false);  // list full: on the spot
                            ++at;
                        }
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------------------------------ consumer warps
    __device__ void consumer(int T, int k, uint32_t &ph_meta0, uint32_t &ph_meta1) {
        const int ctid = threadIdx.x;  // consumers are warps 0..NW-1; the producer is the last warp
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
            const int wbase = h->w_base[b], wend = h->w_end[b];
            const uint32_t tag_s = smem_u32(tag);
            const uint32_t tagv = tag_s - (uint32_t)(wbase >> 2);  // tag byte of document d: tagv + (d >> 2)
            const int last = h->w_last[b];
            const unsigned emask = h->w_emask[b], nmask = h->w_nmask[b];
            const int ME = h->w_ME[b], MN = h->w_MN[b];

            const int nlo = (int)(((long long)MN * cw) / NW), nhi = (int)(((long long)MN * (cw + 1)) / NW);

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
                if (m >= k && p.a.mask == nullptr && theta == 0.0f) {  // only when the k-th-largest table gave no start value
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
            PROF_CNT(17, ME); PROF_CNT(18, MN);
            PROF_CNT(19, MN >= 32768 ? 1 : 0); PROF_CNT(21, MN >= 32768 ? MN : 0);
            // this warp's share of the window's essential postings (flat index over the tokens of emask)
            const int lo = (int)(((long long)ME * cw) / NW), hi = (int)(((long long)ME * (cw + 1)) / NW);

            // ---- A: claim (and reset the hash left by the previous window)
            if (hash_dirty) {
                for (int i = ctid; i < kHash; i += CNT) h->hk[i] = -1;
                hash_dirty = false;
            }
            if (!(p.debug & 16))
                for_segs(b, emask, lo, hi, pool_s, [&](const Seg &g) {
                    pass_claim(g, theta, kept_posting(0.0f, g.rest, theta), tagv, lane);
                });
            PROF(3);
            bsync<true>();
            PROF(4);
            // ---- C: non-essential postings probe the tags
            if (nmask) {
                if (!(p.debug & 1))
                    for_nsegs(b, nmask, nlo, nhi, [&](int t, int off, int n) {
                        ne_probe_seg(t, h->t_beg[t], pool_s + 4u * (unsigned)(h->w_doc[b][t] + off), n, h->w_goff[b][t] + off, tagv, lane);
                    });
                PROF(5);
                bsync<true>();
                PROF(6);
            }
            // ---- D1: fetch the scores of the hash entries; resolve pass
            const bool hits = h->h_any != 0u;
            if (!(p.debug & 4))
                for_segs(b, emask, lo, hi, pool_s, [&](const Seg &g) {
                    pass_resolve(g, theta, kept_posting(0.0f, g.rest, theta), tagv, lane, b, T, pool_s);
                });
            asm volatile("cp.async.wait_all;" ::: "memory");  // this thread's score copies have landed (visible after the barrier)
            if (hits) hash_dirty = true;
            PROF(20);
            bsync<true>();
            PROF(8);
            // ---- D2: slow list (one thread per entry); clear the tags of the window
            {
                const int ns = (int)min(h->slow_cnt, (unsigned)kSlowCap);
                PROF_CNT(23, ns); PROF_CNT(30, ns >= 480 ? 1 : 0); PROF_CNT(31, ns >= 480 ? ns : 0);
                for (int i = ctid; i < ns; i += CNT) {
                    const int2 e = h->slow[i];
                    resolve_one(b, T, pool_s, theta, e.x, e.y, true);
                }
            }
            PROF(7);
            if (!(p.debug & 8)) {
                const int nz = (((wend - wbase) >> 2) + 16) >> 4;  // 16-byte words covering the window's tag bytes
                if (nz <= 8 * ME) {
                    for (int i = ctid; i < nz; i += CNT) sts128z(tag_s + 16u * (unsigned)i);
                } else {
                    for_segs(b, emask, lo, hi, pool_s, [&](const Seg &g) {
                        pass_clean(g, theta, kept_posting(0.0f, g.rest, theta), tagv, lane);
                    });
                }
            }
            PROF(9);
            bsync<true>();
            PROF(10);
            if (ctid == 0) { mbar_arrive(&h->bar_free[b]); h->h_any = 0u; h->h_over = 0u; h->slow_cnt = 0u; }
            if (lane == 0) h->h_cnt_w[cw] = 0u;
            if (h->cb_cnt >= (unsigned)trigger) { compact<true>(k, ctid, CNT); theta = fmaxf(theta, h->theta); PROF_CNT(22, 1); }
            PROF(11);
            if (last) break;
            b ^= 1;
        }
        PROF_FLUSH(ctid == 32 * (p.debug >> 8));
    }

    __device__ void run() {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int k = p.a.k;
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(&h->bar_data[i], 1); mbar_init(&h->bar_meta[i], 1); mbar_init(&h->bar_free[i], 1); }
            fence_mbar_init();
            h->h_any = 0u;
            h->h_over = 0u;
            for (int i = 0; i < kMaxWarps; ++i) h->h_cnt_w[i] = 0u;
            h->slow_cnt = 0u;
        }
        for (int i = tid * 16; i < p.tagcap; i += NT * 16) *reinterpret_cast<uint4 *>(tag + i) = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < kHash; i += NT) h->hk[i] = -1;
        __syncthreads();
        uint32_t ph_data0 = 0, ph_data1 = 0, ph_free0 = 0, ph_free1 = 0, ph_meta0 = 0, ph_meta1 = 0;
        PROF_DECL;
        for (;;) {
            PROF(12);
            if (tid == 0) h->q_slot = atomicAdd(&p.ctrl->fast_next, 1u);
            __syncthreads();
            const unsigned int slot = h->q_slot;
            if (slot >= (unsigned)p.a.n_queries) break;
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
                h->t_beg[tid] = beg;
                h->t_cur[tid] = beg;
                h->t_end[tid] = end;
                h->t_df[tid] = end - beg;
                h->t_max[tid] = mx;
                h->t_kth[tid] = kth;
            }
            if (tid == 0) { h->cb_cnt = 0; h->theta = 0.0f; }
            __syncthreads();
            Layout lay;
            lay.cap = 0; lay.off = 0; lay.ne = -1; lay.live = -1;  // forces the first share-out
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
__global__ void __launch_bounds__(NT, 1) retrieve_v7_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Fast<NT> f(p, smem);
    f.run();
}
// two CTAs per SM (256 threads each): two independent queries overlap each other's latencies
__global__ void __launch_bounds__(256, 2) retrieve_v7_kernel_x2(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Fast<256> f(p, smem);
    f.run();
}

}  // namespace v7

static int env_int7(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int launch_retrieve_v7(bm25cu_index *ix, const RetrieveArgs &a, Ctrl *d_ctrl, unsigned int *d_general_list, int *launches) {
    FastParams p;
    p.data = ix->d_data;
    p.indices = ix->d_indices;
    p.indptr = ix->d_indptr;
    p.colmax = ix->d_colmax;
    p.order = ix->order_valid ? ix->d_order.as<unsigned int>() : nullptr;
    p.qcount = nullptr;
    // the k-th-largest table holds ranks 10 / 100 / 1000: rank r bounds every k <= r (k = 1 uses the column maxima)
    p.colkth = nullptr;
    if (env_int7("BM25CU_SEED_KTH", 1) && ix->d_colkth && a.mask == nullptr) {  // a mask may remove the documents behind the bound
        if (a.k == 1) p.colkth = ix->d_colmax;
        else if (a.k <= 10) p.colkth = ix->d_colkth;
        else if (a.k <= 100) p.colkth = ix->d_colkth + (size_t)ix->n_vocab;
        else if (a.k <= 1000) p.colkth = ix->d_colkth + 2 * (size_t)ix->n_vocab;
    }
    p.prune = env_int7("BM25CU_PRUNE", 1);
    p.max_shift = 2;
    p.n_docs = ix->n_docs;
    p.n_vocab = ix->n_vocab;
    p.doc_base = ix->doc_base;
    p.a = a;
    p.ctrl = d_ctrl;
    p.general_list = d_general_list;
    int threads = env_int7("BM25CU_THREADS", 512);
    if (threads != 256 && threads != 1024) threads = 512;
    const int kcap = next_pow2(a.k < 1 ? 1 : a.k);
    const int cb_min = env_int7("BM25CU_CBMIN", 32);
    // 2 * next_pow2(k) keys for small k; four times for larger k (fewer cuts), never more than 4096 unless k needs it
    int cb_auto = kcap <= 64 ? 2 * kcap : 4 * kcap;
    if (cb_auto > 4096) cb_auto = 2 * kcap > 4096 ? 2 * kcap : 4096;
    p.cbcap = cb_auto < cb_min ? next_pow2(cb_min) : cb_auto;
    p.route_all_general = (!ix->nonneg) || (a.k > 2048) || env_int7("BM25CU_FORCE_GENERAL", 0) ||
                          (a.mask != nullptr && env_int7("BM25CU_MASK_FAST", 1) == 0);
    const int ctas_per_sm = env_int7("BM25CU_CTAS_PER_SM", 1) == 2 ? 2 : 1;
    if (ctas_per_sm == 2) threads = 256;
    const int budget = ix->max_smem_optin / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
    p.nwarps = threads / 32 - 1;
    p.debug = env_int7("BM25CU_DEBUG_SKIP", 0);
    p.hash_max = env_int7("BM25CU_HASHMAX", v7::kHashMax);
    p.theta_mode = 0;
    p.stages = 2;
    p.theta_io = nullptr;
    if (p.hash_max < 0 || p.hash_max > v7::kHashMax) p.hash_max = v7::kHashMax;
    p.hash_max /= (threads / 32 - 1);  // per consumer warp
    int pool = env_int7("BM25CU_POOL", ctas_per_sm == 2 ? 6144 : 14336);  // 4-byte slots per stage (two per essential posting)
    pool = (pool + 3) & ~3;
    const int min_pool = kMaxTok * 2 * kMinCap + 256;
    if (pool < min_pool) pool = min_pool;
    if (pool > 32768) pool = 32768;
    p.poolcap = pool;
    size_t fixed = v7::Fast<512>::fixed_smem(p);
    long long tagcap = (long long)budget - (long long)fixed - 64;
    if (tagcap < 4096) {
        p.route_all_general = 1;  // k too large for this shared-memory budget
        p.cbcap = 512;
        fixed = v7::Fast<512>::fixed_smem(p);
        tagcap = (long long)budget - (long long)fixed - 64;
    }
    const int tag_env = env_int7("BM25CU_TAG", 0);
    if (tag_env > 0 && tag_env < tagcap) tagcap = tag_env;
    p.tagcap = (int)(tagcap & ~15ll);
    p.trigger = env_int7("BM25CU_TRIGGER", p.cbcap / 2);
    if (p.trigger <= a.k) p.trigger = a.k + 1;
    if (p.trigger > p.cbcap) p.trigger = p.cbcap;
    p.spillcap = 2 * p.poolcap;
    const int grid = ix->sm_count * ctas_per_sm;
    int rc = ix->d_spill.reserve((size_t)grid * (size_t)p.spillcap * 8);
    if (rc) return rc;
    p.spill = ix->d_spill.as<unsigned long long>();
    const size_t smem = fixed + p.tagcap;
    cudaError_t e;
#define BM25CU_LAUNCH(NTV)                                                                                        \
    e = cudaFuncSetAttribute(v7::retrieve_v7_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) v7::retrieve_v7_kernel<NTV><<<grid, NTV, smem, ix->stream>>>(p);
    if (ctas_per_sm == 2) {
        e = cudaFuncSetAttribute(v7::retrieve_v7_kernel_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) v7::retrieve_v7_kernel_x2<<<grid, 256, smem, ix->stream>>>(p);
    } else
    switch (threads) {
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
