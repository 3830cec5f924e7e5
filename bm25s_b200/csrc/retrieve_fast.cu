// K1: fused streaming scoring + top-k for a batch of queries (sm_100a).
//
// Replaces, per query, BM25._compute_relevance_from_scores (bm25s/__init__.py:312-318), the BM25L/BM25+ shift
// (bm25s/__init__.py:616-618) and selection.topk (bm25s/selection.py:14-37) without ever materialising a
// num_docs vector. One persistent CTA per SM pulls queries from a device-side queue. For a query it walks the
// document range in *windows*; a window's postings of every query token are streamed from HBM into a
// double-buffered shared-memory stage by 1-D bulk async copies (cp.async.bulk + mbarrier, the TMA engine),
// the next window being prefetched while the current one is processed.
//
// Exactness (bit-identical float32 scores): a document hit by ONE query token scores exactly that posting's value
// (0 + s = s), so no accumulator is needed for it. Ownership of a document inside a window is resolved through a
// one-byte tag per document in shared memory (three passes: write tag = token slot; losers mark 0x80|slot;
// winners/last-loser read back). Documents hit by >= 2 tokens ("contested", a few percent) are re-summed in the
// reference's order - token by token in query order - by one thread each, through binary searches of the staged
// window postings. Candidates above the running k-th best score go to a shared-memory candidate buffer that is
// compacted by a radix select on (score, ~doc) keys, which also makes ties deterministic (lower doc id first).
//
// The kernel requires non-negative finite scores (true for the five BM25 variants; checked at upload), at most
// 64 query tokens and no weight mask; anything else is routed to the dense general kernel (general.cu).
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace bm25cu {

constexpr int kMaxTok = 64;   // token slots per query on this path (tag byte: slot, or 0x80 | slot)
constexpr int kMinCap = 16;   // minimum staged postings per token per window (multiple of 4)

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
    int cbcap;     // candidate entries per half
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
    uint64_t mbar[2];
    long long t_cur[kMaxTok];
    long long t_end[kMaxTok];
    int t_cap[kMaxTok];
    int t_off[kMaxTok];       // element offset of the token's region inside a stage buffer
    int t_skip[2][kMaxTok];   // leading elements to skip (16-byte alignment of the global source)
    int t_n[2][kMaxTok];      // valid staged postings
    int t_first[kMaxTok];
    int t_e[kMaxTok];
    int t_cnt[kMaxTok];       // postings of the token inside the current window
    int w_base[kMaxTok];      // t_off + t_skip of the current buffer
    int prefix[kMaxTok + 1];
    unsigned int hist[256];
    SelectState<unsigned long long> sel;
    unsigned int q_slot;
    unsigned int cb_cnt;
    unsigned int new_cnt;
    unsigned int r_cnt;
    int cb_half;
    int last;
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
    SmemHdr *h;
    unsigned long long *cb;   // [2][cbcap]
    int *rlist;               // [rcap]
    int *sdocs;               // [2][stagecap]
    float *sscores;           // [2][stagecap]
    unsigned char *tag;       // [tagcap]
    const FastParams &p;
    unsigned long long *spill;

    __device__ Fast(const FastParams &pp, unsigned char *smem) : p(pp) {
        h = reinterpret_cast<SmemHdr *>(smem);
        size_t o = (sizeof(SmemHdr) + 15) & ~size_t(15);
        cb = reinterpret_cast<unsigned long long *>(smem + o);
        o += (size_t)2 * p.cbcap * 8;
        rlist = reinterpret_cast<int *>(smem + o);
        o += (size_t)p.rcap * 4;
        o = (o + 15) & ~size_t(15);
        sdocs = reinterpret_cast<int *>(smem + o);
        o += (size_t)2 * p.stagecap * 4;
        sscores = reinterpret_cast<float *>(smem + o);
        o += (size_t)2 * p.stagecap * 4;
        tag = smem + o;
        spill = p.spill + (size_t)blockIdx.x * p.spillcap;
    }

    __device__ __forceinline__ void emit(float s, int d) {
        const unsigned int slot = atomicAdd(&h->cb_cnt, 1u);
        const unsigned long long c = pack_cand(s, d);
        if (slot < (unsigned)p.cbcap) cb[(size_t)h->cb_half * p.cbcap + slot] = c;
        else if (slot - p.cbcap < (unsigned)p.spillcap) spill[slot - p.cbcap] = c;
    }

    __device__ __forceinline__ unsigned long long cand_at(unsigned int i, int half) const {
        return (i < (unsigned)p.cbcap) ? cb[(size_t)half * p.cbcap + i] : spill[i - p.cbcap];
    }

    // Keep the k best candidates (key order: score desc, doc asc), raise theta to the k-th best score.
    __device__ void compact(int k) {
        __syncthreads();
        const unsigned int n = min(h->cb_cnt, (unsigned)(p.cbcap + p.spillcap));
        const int half = h->cb_half;
        if (n <= (unsigned)k) {
            if (threadIdx.x == 0) h->cb_cnt = n;
            __syncthreads();
            return;
        }
        auto load = [&](long long i) { return cand_at((unsigned)i, half); };
        unsigned long long thr;
        long long need;
        block_radix_select<unsigned long long>(load, (long long)n, (long long)k, h->hist, &h->sel, ~0ull, thr, need);
        if (threadIdx.x == 0) h->new_cnt = 0;
        __syncthreads();
        unsigned long long *dst = cb + (size_t)(half ^ 1) * p.cbcap;
        for (unsigned int i = threadIdx.x; i < n; i += NT) {
            const unsigned long long c = cand_at(i, half);
            if (c >= thr) dst[atomicAdd(&h->new_cnt, 1u)] = c;  // keys are unique: exactly k survive
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            h->cb_cnt = h->new_cnt;
            h->cb_half = half ^ 1;
            h->theta = cand_score(thr);
        }
        __syncthreads();
    }

    // warp 0: stage the next window of every live token into buffer b
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
                mbar_expect_tx(&h->mbar[b], 2 * bytes);
                const size_t so = (size_t)b * p.stagecap + h->t_off[t];
                bulk_g2s(sdocs + so, p.indices + (cur - a), bytes, &h->mbar[b]);
                bulk_g2s(sscores + so, p.data + (cur - a), bytes, &h->mbar[b]);
            }
            h->t_skip[b][t] = a;
            h->t_n[b][t] = n;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&h->mbar[b]);
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

    __device__ void run() {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        constexpr int NW = NT / 32;
        const int k = p.a.k;
        if (tid == 0) {
            mbar_init(&h->mbar[0], 1);
            mbar_init(&h->mbar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        uint32_t phase0 = 0, phase1 = 0;
        const int trigger = min(p.cbcap / 2, max(4 * k, 256));

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
            if (tid == 0) { h->cb_cnt = 0; h->cb_half = 0; h->theta = 0.0f; h->r_cnt = 0; h->last = 0; }
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
            float theta = 0.0f;
            if (h->total_df > 0) {
                int b = 0;
                for (;;) {
                    if (b == 0) { mbar_wait(&h->mbar[0], phase0); phase0 ^= 1; }
                    else { mbar_wait(&h->mbar[1], phase1); phase1 ^= 1; }
                    const int *bd = sdocs + (size_t)b * p.stagecap;
                    const float *bs = sscores + (size_t)b * p.stagecap;
                    // ---- window bounds: first staged doc of every token, coverage end of truncated tokens
                    for (int t = warp; t < T; t += NW) {
                        if (lane == 0) {
                            const int n = h->t_n[b][t];
                            int first = 0x7fffffff, e = 0x7fffffff;
                            if (n > 0) {
                                const int *d = bd + h->t_off[t] + h->t_skip[b][t];
                                first = d[0];
                                if (h->t_cur[t] + n < h->t_end[t]) e = d[n - 1] + 1;
                            }
                            h->t_first[t] = first;
                            h->t_e[t] = e;
                        }
                    }
                    __syncthreads();
                    int base, wend;
                    {
                        int f = min(lane < T ? h->t_first[lane] : 0x7fffffff, lane + 32 < T ? h->t_first[lane + 32] : 0x7fffffff);
                        int e = min(lane < T ? h->t_e[lane] : 0x7fffffff, lane + 32 < T ? h->t_e[lane + 32] : 0x7fffffff);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
                            e = min(e, __shfl_xor_sync(0xffffffffu, e, o));
                        }
                        base = f;
                        const int cap_end = (base > 0x7fffffff - p.tagcap) ? 0x7fffffff : base + p.tagcap;
                        wend = min(e, cap_end);
                    }
                    for (int t = warp; t < T; t += NW) {
                        const int n = h->t_n[b][t];
                        int c = 0;
                        if (n > 0) {
                            const int *d = bd + h->t_off[t] + h->t_skip[b][t];
                            c = (d[n - 1] < wend) ? n : warp_count_below(d, n, wend);
                        }
                        if (lane == 0) h->t_cnt[t] = c;
                    }
                    __syncthreads();
                    {   // every warp derives the same prefix sums (identical values: benign duplicate writes)
                        const int c0 = lane < T ? h->t_cnt[lane] : 0, c1 = lane + 32 < T ? h->t_cnt[lane + 32] : 0;
                        int x = c0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                        const int tot0 = __shfl_sync(0xffffffffu, x, 31);
                        int x1 = c1;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x1, o); if (lane >= o) x1 += y; }
                        if (lane == 0) h->prefix[0] = 0;
                        if (lane < T) { h->prefix[lane + 1] = x; h->w_base[lane] = h->t_off[lane] + h->t_skip[b][lane]; }
                        if (lane + 32 < T) { h->prefix[lane + 33] = tot0 + x1; h->w_base[lane + 32] = h->t_off[lane + 32] + h->t_skip[b][lane + 32]; }
                        __syncwarp();
                        if (warp == 0) {  // advance cursors, prefetch the next window into the other buffer
                            bool more = false;
                            if (lane < T) { h->t_cur[lane] += c0; more |= h->t_cur[lane] < h->t_end[lane]; }
                            if (lane + 32 < T) { h->t_cur[lane + 32] += c1; more |= h->t_cur[lane + 32] < h->t_end[lane + 32]; }
                            const bool any_more = __any_sync(0xffffffffu, more);
                            __syncwarp();
                            if (lane == 0) h->last = any_more ? 0 : 1;
                            if (any_more) issue_loads(b ^ 1, T);
                        }
                    }
                    const int M = h->prefix[T];
                    // ---- P1: every posting claims its document's tag byte
                    for (int i = tid, t = 0; i < M; i += NT) {
                        while (i >= h->prefix[t + 1]) ++t;
                        const int d = bd[h->w_base[t] + (i - h->prefix[t])];
                        tag[d - base] = (unsigned char)t;
                    }
                    __syncthreads();
                    // ---- P2: postings that lost the claim mark the document as contested
                    for (int i = tid, t = 0; i < M; i += NT) {
                        while (i >= h->prefix[t + 1]) ++t;
                        const int d = bd[h->w_base[t] + (i - h->prefix[t])];
                        if (tag[d - base] != (unsigned char)t) tag[d - base] = (unsigned char)(0x80 | t);
                    }
                    __syncthreads();
                    // ---- P3: uncontested owners score directly; the surviving marker owns the contested document
                    for (int i = tid, t = 0; i < M; i += NT) {
                        while (i >= h->prefix[t + 1]) ++t;
                        const int off = h->w_base[t] + (i - h->prefix[t]);
                        const int d = bd[off];
                        const unsigned char x = tag[d - base];
                        if (x == (unsigned char)t) {
                            const float s = bs[off];
                            if (s > theta) emit(s, d);
                        } else if (x == (unsigned char)(0x80 | t)) {
                            const unsigned int r = atomicAdd(&h->r_cnt, 1u);
                            if (r < (unsigned)p.rcap) rlist[r] = d;
                        }
                    }
                    __syncthreads();
                    // ---- P4: contested documents are summed in query-token order (the reference's float32 order)
                    const int R = min((int)h->r_cnt, p.rcap);
                    for (int r = tid; r < R; r += NT) {
                        const int d = rlist[r];
                        float sum = 0.0f;
                        for (int t = 0; t < T; ++t) {
                            const int c = h->t_cnt[t];
                            if (c == 0) continue;
                            const int *dt = bd + h->w_base[t];
                            int lo = 0, hi = c;
                            while (lo < hi) { const int m = (lo + hi) >> 1; if (dt[m] < d) lo = m + 1; else hi = m; }
                            if (lo < c && dt[lo] == d) sum = __fadd_rn(sum, bs[h->w_base[t] + lo]);
                        }
                        if (sum > theta) emit(sum, d);
                    }
                    __syncthreads();
                    if (tid == 0) h->r_cnt = 0;
                    if (h->cb_cnt >= (unsigned)trigger) { compact(k); theta = h->theta; }
                    const int last = h->last;
                    __syncthreads();
                    if (last) break;
                    b ^= 1;
                }
            }
            finalize(q, T);
        }
    }

    // Sort the surviving candidates, pad with untouched documents if fewer than k matched, write the result row.
    __device__ void finalize(int q, int T) {
        const int tid = threadIdx.x;
        const int k = p.a.k;
        compact(k);
        const int n = (int)h->cb_cnt;  // <= k
        const int half = h->cb_half;
        unsigned long long *cur = cb + (size_t)half * p.cbcap;
        unsigned long long *oth = cb + (size_t)(half ^ 1) * p.cbcap;
        const int P = next_pow2(n > 0 ? n : 1);
        for (int i = n + tid; i < P; i += NT) cur[i] = 0ull;
        __syncthreads();
        sort_desc_keys(cur, P);
        const bool has_shift = (p.a.shift != nullptr) && T > 0;  // empty query: zeros, no shift (__init__.py:653-657)
        const float shift = has_shift ? p.a.shift[q] : 0.0f;
        int32_t *oi = p.a.out_ids + (size_t)q * k;
        float *os = p.a.out_scores + (size_t)q * k;
        for (int j = tid; j < n; j += NT) {
            const unsigned long long c = cur[j];
            float s = cand_score(c);
            if (has_shift) s = __fadd_rn(s, shift);
            oi[j] = cand_doc(c) + p.doc_base;
            os[j] = s;
        }
        if (n < k) {
            // padding: the (k - n) lowest document ids that are not candidates score 0 (+ shift)
            const int P2 = P;
            for (int i = tid; i < P2; i += NT) oth[i] = (i < n) ? (unsigned long long)(unsigned int)cand_doc(cur[i]) : ~0ull;
            __syncthreads();
            sort_asc_keys(oth, P2);
            const float pad_score = has_shift ? __fadd_rn(0.0f, shift) : 0.0f;
            for (int j = n + tid; j < k; j += NT) {
                const long long m = j - n;  // m-th (0-based) non-candidate document id
                int lo = 0, hi = n;         // first i with oth[i] - i > m
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((long long)oth[mid] - mid > m) hi = mid; else lo = mid + 1;
                }
                const long long x = m + lo;
                if (x < p.n_docs) { oi[j] = (int32_t)x + p.doc_base; os[j] = pad_score; }
                else { oi[j] = -1; os[j] = __int_as_float(0xff800000); }
            }
        }
        __syncthreads();
    }

    // keys-only bitonic sorts on shared memory (block-wide)
    __device__ void sort_desc_keys(unsigned long long *key, int P) {
        for (int size = 2; size <= P; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = threadIdx.x; t < (P >> 1); t += NT) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const unsigned long long a = key[lo], b2 = key[hi];
                    if (desc ? (a < b2) : (a > b2)) { key[lo] = b2; key[hi] = a; }
                }
            }
        __syncthreads();
    }
    __device__ void sort_asc_keys(unsigned long long *key, int P) {
        for (int size = 2; size <= P; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = threadIdx.x; t < (P >> 1); t += NT) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool asc = ((lo & size) == 0);
                    const unsigned long long a = key[lo], b2 = key[hi];
                    if (asc ? (a > b2) : (a < b2)) { key[lo] = b2; key[hi] = a; }
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
    size_t fixed = ((sizeof(SmemHdr) + 15) & ~size_t(15)) + (size_t)2 * p.cbcap * 8 + (((size_t)p.rcap * 4 + 15) & ~size_t(15)) +
                   (size_t)4 * p.stagecap * 4;
    long long tagcap = (long long)budget - (long long)fixed - 64;
    if (tagcap < 4096) {
        p.route_all_general = 1;  // k too large for this smem budget
        tagcap = 4096;
        p.cbcap = 512;
        fixed = ((sizeof(SmemHdr) + 15) & ~size_t(15)) + (size_t)2 * p.cbcap * 8 + (((size_t)p.rcap * 4 + 15) & ~size_t(15)) +
                (size_t)4 * p.stagecap * 4;
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
