// Shared device helpers of the K1 kernels: in-kernel phase profiling macros, launch parameters, mbarrier /
// bulk-async-copy PTX wrappers, candidate keys and explicit shared-space accessors.
#pragma once

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

constexpr int kMaxTok = 15;     // token slots per query on this path (4 bits of the tag byte)
constexpr int kTokArr = 16;
constexpr int kMinCap = 16;     // minimum staged postings per live token per window (multiple of 4)
constexpr int kMaxWarps = 32;
constexpr int kMaxChunk = 48;   // 32-posting chunks of essential postings per consumer warp and window (more: slice path)
constexpr unsigned kTagValid = 4u, kTagContested = 0x80u;

struct FastParams {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    const float *colmax;
    const unsigned int *order;  // queries in the order the CTAs pull them (longest first); nullptr: index order
    const unsigned int *qcount;  // non-null: `order` lists *qcount queries (the rest of the batch was served by retrieve_ms.cu)
    const float *colkth;  // column of the k-th-largest table that is valid for this k (nullptr: no seed)
    int32_t n_docs, n_vocab, doc_base;
    RetrieveArgs a;
    Ctrl *ctrl;
    unsigned int *general_list;
    int tagcap;    // bytes of the tag array; a window spans at most tagcap * G documents
    int poolcap;   // 4-byte slots per stage buffer (doc ids, and scores of the essential tokens)
    int cbcap;     // candidate entries (power of two, >= 2 * next_pow2(k))
    unsigned long long *spill;  // per-CTA global overflow of the candidate buffer
    int spillcap;
    int route_all_general;
    int prune;     // 0: every token is essential and nothing is dropped (exhaustive), 1: threshold pruning
    int max_shift; // largest log2(G)
    int nwarps;    // consumer warps of the launch configuration
    int debug;     // timing experiments only (BM25CU_DEBUG_SKIP bit mask; results are wrong when non-zero)
    float *theta_io;  // experiment: per-query final threshold written (theta_mode 1) / used as the start value (2)
    int theta_mode;
    int trigger;   // candidate count at a window end that starts a cut of the buffer
    int stages;    // stage buffers of the posting pool (2: the next window loads while this one is processed)
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

__device__ __forceinline__ unsigned long long pack_cand(float s, int d) {
    return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned int)(~d);
}
__device__ __forceinline__ float cand_score(unsigned long long c) { return __uint_as_float((unsigned int)(c >> 32)); }
__device__ __forceinline__ int cand_doc(unsigned long long c) { return (int)~(unsigned int)(c & 0xffffffffu); }

// Explicit shared-space accesses (32-bit shared addresses): the generic-pointer path costs an address-space
// conversion per access in these loops.
__device__ __forceinline__ unsigned lds32(uint32_t a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float ldsf(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds8(uint32_t a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts8(uint32_t a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// A posting (d, s) of an essential token is KEPT when s + rest (the other tokens' column maxima) can exceed theta.
__device__ __forceinline__ bool kept_posting(float s, float rest, float theta) {
    return __fmul_ru(__fadd_ru(s, rest), 1.00002f) > theta;
}

}  // namespace bm25cu
