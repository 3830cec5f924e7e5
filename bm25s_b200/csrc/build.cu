// GPU index build (K4-K8) - placeholder until the sort-based builder lands; every entry point fails loudly.
#include "common.cuh"

using namespace bm25cu;

struct bm25cu_builder { int unused; };

extern "C" {
int bm25cu_build_begin(const int32_t *, const int64_t *, int32_t, int32_t, int32_t, int32_t, bm25cu_builder **) {
    set_error("bm25cu_build_begin: GPU index build not implemented yet");
    return BM25CU_EINVAL;
}
int bm25cu_build_df_device(bm25cu_builder *, int32_t **) { set_error("not implemented"); return BM25CU_EINVAL; }
int bm25cu_build_df_get(bm25cu_builder *, int32_t *) { set_error("not implemented"); return BM25CU_EINVAL; }
int bm25cu_build_df_set(bm25cu_builder *, const int32_t *) { set_error("not implemented"); return BM25CU_EINVAL; }
int bm25cu_build_finish(bm25cu_builder *, int64_t, int64_t, int32_t, int32_t, double, double, double, int32_t, bm25cu_index **) {
    set_error("not implemented");
    return BM25CU_EINVAL;
}
int bm25cu_build_free(bm25cu_builder *) { return BM25CU_OK; }
int bm25cu_index_build(const int32_t *, const int64_t *, int32_t, int32_t, int32_t, int32_t, double, double, double,
                       int32_t, int32_t, bm25cu_index **) {
    set_error("bm25cu_index_build: GPU index build not implemented yet");
    return BM25CU_EINVAL;
}
}
