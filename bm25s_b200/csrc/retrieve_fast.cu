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

constexpr int kMaxTok = 64;    // token slots per query on this path (tag byte: slot, or 0x80 | slot)
constexpr int kMinCap = 16;    // minimum staged postings per token per window (multiple of 4)
constexpr int kPartCap = 4096; // floats of scratch for the contested-document partial scores

struct FastParams {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    int32_t n_docs, n_vocab, doc_base;
    RetrieveArgs a;
    Ctrl *ctrl;
    unsigned int *general_list;
    int tagcap;    // bytes of tag array == max doc span of a window
    int stagecap;  // postings per stage buffer
    int cbcap;     // candidate entries (power of two, >= 2 * next_pow2(k))
    int rcap;      // contested-doc list entries
    unsigned long long *spill;  // per-CTA global overflow of the candidate buffer
    int spillcap;
    int route_all_general;
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
    int t_cap[kMaxTok];
    int t_off[kMaxTok];      // element offset of the token's region inside a stage buffer
    int t_skip[2][kMaxTok];  // leading elements to skip (16-byte alignment of the global source)
    int t_n[2][kMaxTok];     // valid staged postings
    int w_cnt[2][kMaxTok];   // postings of the token inside the window of stage b
    int w_off[2][kMaxTok];   // t_off + t_skip of stage b
    int w_base[2];           // first document of the window
    int w_last[2];           // 1 if this is the query's last window
    int w_total[2];          // postings in the window
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
    SmemHdr *h;
    unsigned long long *cb;   // [cbcap]
    int *rlist;               // [rcap]
    float *part;              // [kPartCap]
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
        o += (size_t)kPartCap * 4;
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
        o += (size_t)p.cbcap * 8 + (size_t)p.rcap * 4 + (size_t)kPartCap * 4;
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
                    if (keep == (unsigned)k) h->theta = cand_score(cb[k - 1]);
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

    // producer: stage the next window of every live token into buffer b
    __device__ void issue_loads(int b, int T) {
        const int lane = threadIdx.x & 31;
        for (int t = lane; t < T; t += 32) {
            const long long cur = h->t_cur[t], rem = h->t_end[t] - cur;
            int n = 0, a = 0;
            if (rem > 0) {
                a = (int)(cur & 3);
                const int cap = h->t_cap[t];
                n = (int)min((long long)(cap - a), rem);
                const uint32_t bytes = (uint32_t)(((a + n + 3) & ~3) * 4);
                mbar_expect_tx(&h->bar_data[b], 2 * bytes);
                const size_t so = (size_t)b * p.stagecap + h->t_off[t];
                bulk_g2s(sdocs + so, p.indices + (cur - a), bytes, &h->bar_data[b]);
                bulk_g2s(sscores + so, p.data + (cur - a), bytes, &h->bar_data[b]);
            }
            h->t_skip[b][t] = a;
            h->t_n[b][t] = n;
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
    __device__ void producer(int T, uint32_t &ph_data0, uint32_t &ph_data1, uint32_t &ph_free0, uint32_t &ph_free1) {
        const int lane = threadIdx.x & 31;
        int b = 0;
        int nwin = 0;
        for (;;) {
            if (b == 0) { mbar_wait(&h->bar_data[0], ph_data0); ph_data0 ^= 1; }
            else { mbar_wait(&h->bar_data[1], ph_data1); ph_data1 ^= 1; }
            const int *bd = sdocs + (size_t)b * p.stagecap;
            // per lane: tokens lane and lane + 32
            int n0 = 0, n1 = 0, off0 = 0, off1 = 0, first = 0x7fffffff, e = 0x7fffffff, last0 = -1, last1 = -1;
            if (lane < T) {
                n0 = h->t_n[b][lane];
                off0 = h->t_off[lane] + h->t_skip[b][lane];
                if (n0 > 0) {
                    first = bd[off0];
                    last0 = bd[off0 + n0 - 1];
                    if (h->t_cur[lane] + n0 < h->t_end[lane]) e = last0 + 1;
                }
            }
            if (lane + 32 < T) {
                n1 = h->t_n[b][lane + 32];
                off1 = h->t_off[lane + 32] + h->t_skip[b][lane + 32];
                if (n1 > 0) {
                    first = min(first, bd[off1]);
                    last1 = bd[off1 + n1 - 1];
                    if (h->t_cur[lane + 32] + n1 < h->t_end[lane + 32]) e = min(e, last1 + 1);
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
            int c1 = (n1 > 0 && last1 < wend) ? n1 : 0;
            // tokens whose staged postings reach beyond the window end need a search (rare)
            unsigned int need0 = __ballot_sync(0xffffffffu, n0 > 0 && last0 >= wend);
            while (need0) {
                const int src = __ffs(need0) - 1;
                need0 &= need0 - 1;
                const int so = __shfl_sync(0xffffffffu, off0, src), sn = __shfl_sync(0xffffffffu, n0, src);
                const int c = warp_count_below(bd + so, sn, wend);
                if (lane == src) c0 = c;
            }
            unsigned int need1 = __ballot_sync(0xffffffffu, n1 > 0 && last1 >= wend);
            while (need1) {
                const int src = __ffs(need1) - 1;
                need1 &= need1 - 1;
                const int so = __shfl_sync(0xffffffffu, off1, src), sn = __shfl_sync(0xffffffffu, n1, src);
                const int c = warp_count_below(bd + so, sn, wend);
                if (lane == src) c1 = c;
            }
            bool more = false;
            if (lane < T) {
                h->w_cnt[b][lane] = c0;
                h->w_off[b][lane] = off0;
                const long long nc = h->t_cur[lane] + c0;
                h->t_cur[lane] = nc;
                more |= nc < h->t_end[lane];
            }
            if (lane + 32 < T) {
                h->w_cnt[b][lane + 32] = c1;
                h->w_off[b][lane + 32] = off1;
                const long long nc = h->t_cur[lane + 32] + c1;
                h->t_cur[lane + 32] = nc;
                more |= nc < h->t_end[lane + 32];
            }
            int tot = c0 + c1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const bool any_more = __any_sync(0xffffffffu, more);
            if (lane == 0) {
                h->w_base[b] = base;
                h->w_last[b] = any_more ? 0 : 1;
                h->w_total[b] = tot;
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
        // consume the release of the stages not waited for above (keeps barrier parities in step across queries)
        // windows nwin-2 (if any) and nwin-1 are still unconsumed
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

    // ------------------------------------------------------------------------------------------ consumer warps
    __device__ void consumer(int T, int k, uint32_t &ph_meta0, uint32_t &ph_meta1) {
        const int ctid = threadIdx.x - 32;
        const int trigger = p.cbcap / 2;
        float theta = 0.0f;
        int b = 0;
        bool first_window = true;
        for (;;) {
            if (b == 0) { mbar_wait(&h->bar_meta[0], ph_meta0); ph_meta0 ^= 1; }
            else { mbar_wait(&h->bar_meta[1], ph_meta1); ph_meta1 ^= 1; }
            const int *bd = sdocs + (size_t)b * p.stagecap;
            const float *bs = sscores + (size_t)b * p.stagecap;
            const int base = h->w_base[b];
            const int last = h->w_last[b];
            const int *wcnt = h->w_cnt[b];
            const int *woff = h->w_off[b];

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
                }
            }

            // ---- P1: every posting claims its document's tag byte
            for (int t = 0; t < T; ++t) {
                const int c = wcnt[t];
                const int *docs = bd + woff[t];
#pragma unroll 4
                for (int i = ctid; i < c; i += CNT) tag[docs[i] - base] = (unsigned char)t;
            }
            bsync<true>();
            // ---- P2: postings that lost the claim mark the document as contested
            for (int t = 0; t < T; ++t) {
                const int c = wcnt[t];
                const int *docs = bd + woff[t];
#pragma unroll 4
                for (int i = ctid; i < c; i += CNT) {
                    const int o = docs[i] - base;
                    if (tag[o] != (unsigned char)t) tag[o] = (unsigned char)(0x80 | t);
                }
            }
            bsync<true>();
            // ---- P3: uncontested owners score directly; the surviving marker owns the contested document
            for (int t = 0; t < T; ++t) {
                const int c = wcnt[t];
                const int *docs = bd + woff[t];
                const float *sc = bs + woff[t];
#pragma unroll 4
                for (int i = ctid; i < c; i += CNT) {
                    const int d = docs[i];
                    const unsigned char x = tag[d - base];
                    if (x == (unsigned char)t) {
                        const float s = sc[i];
                        if (s > theta) emit(s, d);
                    } else if (x == (unsigned char)(0x80 | t)) {
                        const unsigned int r = atomicAdd(&h->r_cnt, 1u);
                        if (r < (unsigned)p.rcap) rlist[r] = d;
                    }
                }
            }
            bsync<true>();
            // ---- P4: contested documents are summed in query-token order (the reference's float32 order).
            //      a) one thread per (document, token): binary search of the token's window postings
            //      b) one thread per document: ordered sum of the partial scores found
            const int R = min((int)h->r_cnt, p.rcap);
            if (R > 0) {
                const int chunk_docs = max(1, kPartCap / T);
                for (int r0 = 0; r0 < R; r0 += chunk_docs) {
                    const int nd = min(chunk_docs, R - r0);
                    const int pairs = nd * T;
                    for (int j = ctid; j < pairs; j += CNT) {
                        const int r = j / T, t = j - r * T;
                        const int d = rlist[r0 + r];
                        const int c = wcnt[t];
                        float v = -1.0f;
                        if (c > 0) {
                            const int *dt = bd + woff[t];
                            int lo = 0, hi = c;
                            while (lo < hi) { const int m = (lo + hi) >> 1; if (dt[m] < d) lo = m + 1; else hi = m; }
                            if (lo < c && dt[lo] == d) v = bs[woff[t] + lo];
                        }
                        part[j] = v;
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
            // all consumers are done with stage b and the tags of this window
            if (ctid == 0) { h->r_cnt = 0; mbar_arrive(&h->bar_free[b]); }
            if (h->cb_cnt >= (unsigned)trigger) { compact<true>(k, ctid, CNT); theta = fmaxf(theta, h->theta); }
            if (last) break;
            b ^= 1;
        }
    }

    __device__ void run() {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int k = p.a.k;
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) { mbar_init(&h->bar_data[i], 1); mbar_init(&h->bar_meta[i], 1); mbar_init(&h->bar_free[i], 1); }
            fence_mbar_init();
        }
        __syncthreads();
        uint32_t ph_data0 = 0, ph_data1 = 0, ph_free0 = 0, ph_free1 = 0, ph_meta0 = 0, ph_meta1 = 0;

        for (;;) {
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
                if ((unsigned)tok >= (unsigned)p.n_vocab) atomicOr(&p.ctrl->error, 1u);
                else { beg = p.indptr[tok]; end = p.indptr[tok + 1]; }
                h->t_cur[tid] = beg;
                h->t_end[tid] = end;
            }
            if (tid == 0) { h->cb_cnt = 0; h->theta = 0.0f; h->r_cnt = 0; }
            __syncthreads();
            if (warp == 0) {  // stage capacities proportional to document frequency
                long long df0 = (lane < T) ? h->t_end[lane] - h->t_cur[lane] : 0;
                long long df1 = (lane + 32 < T) ? h->t_end[lane + 32] - h->t_cur[lane + 32] : 0;
                long long tot = df0 + df1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                const long long freecap = (long long)p.stagecap - (long long)T * kMinCap;
                int c0 = 0, c1 = 0;
                if (lane < T) c0 = kMinCap + (int)(tot > 0 ? ((freecap * df0) / tot) & ~3ll : 0);
                if (lane + 32 < T) c1 = kMinCap + (int)(tot > 0 ? ((freecap * df1) / tot) & ~3ll : 0);
                int x = c0;  // exclusive scan of caps -> offsets (slots 0..31 then 32..63)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                const int tot0 = __shfl_sync(0xffffffffu, x, 31);
                int x1 = c1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x1, o); if (lane >= o) x1 += y; }
                if (lane < T) { h->t_cap[lane] = c0; h->t_off[lane] = x - c0; }
                if (lane + 32 < T) { h->t_cap[lane + 32] = c1; h->t_off[lane + 32] = tot0 + x1 - c1; }
                if (lane == 0) {
                    h->total_df = (unsigned long long)tot;
                    if (tot > 0) atomicAdd(&p.ctrl->posting_count, (unsigned long long)tot);
                }
                __syncwarp();
                if (tot > 0) issue_loads(0, T);
            }
            __syncthreads();
            if (h->total_df > 0) {
                if (warp == 0) producer(T, ph_data0, ph_data1, ph_free0, ph_free1);
                else consumer(T, k, ph_meta0, ph_meta1);
            }
            __syncthreads();
            finalize(q, T);
        }
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
    p.n_docs = ix->n_docs;
    p.n_vocab = ix->n_vocab;
    p.doc_base = ix->doc_base;
    p.a = a;
    p.ctrl = d_ctrl;
    p.general_list = d_general_list;
    const int threads = env_int("BM25CU_THREADS", 512);
    const int ctas_per_sm = env_int("BM25CU_CTAS_PER_SM", 1);
    const int kcap = next_pow2(a.k < 1 ? 1 : a.k);
    p.cbcap = 2 * kcap < 512 ? 512 : 2 * kcap;
    p.route_all_general = (!ix->nonneg) || (a.mask != nullptr) || (a.k > 2048) || env_int("BM25CU_FORCE_GENERAL", 0);
    const int budget = ix->max_smem_optin / (ctas_per_sm > 0 ? ctas_per_sm : 1) - (ctas_per_sm > 1 ? 1024 : 0);
    int stage = env_int("BM25CU_STAGE", ctas_per_sm > 1 ? 2048 : 4096);
    stage = (stage + 3) & ~3;
    if (stage < kMaxTok * kMinCap * 2) stage = kMaxTok * kMinCap * 2;
    p.stagecap = stage;
    p.rcap = stage / 2 + 64;
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
    switch (threads) {
        case 256:
            e = cudaFuncSetAttribute(retrieve_fast_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) retrieve_fast_kernel<256><<<grid, 256, smem, ix->stream>>>(p);
            break;
        case 1024:
            e = cudaFuncSetAttribute(retrieve_fast_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) retrieve_fast_kernel<1024><<<grid, 1024, smem, ix->stream>>>(p);
            break;
        default:
            e = cudaFuncSetAttribute(retrieve_fast_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) retrieve_fast_kernel<512><<<grid, 512, smem, ix->stream>>>(p);
            break;
    }
    BM25CU_CHECK(e);
    BM25CU_CHECK(cudaGetLastError());
    if (launches) ++*launches;
    return BM25CU_OK;
}

}  // namespace bm25cu
