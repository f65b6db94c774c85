// bk_dist.cu — multi-GPU (row-partitioned) entry points.  Filled in by the distributed milestone;
// until then every entry reports BK_ERR_UNSUPPORTED so that callers fail loudly.
#include "bk_internal.cuh"

extern "C" int bk_dist_unique_id(void*) { return bk_fail(BK_ERR_UNSUPPORTED, "bk_dist: not built yet"); }
extern "C" int bk_dist_create(bk_handle*, const void*, int, int, int64_t, int64_t, int64_t, int64_t, const void*,
                              const void*, int, const void*, int, void*, bk_dist**) {
  return bk_fail(BK_ERR_UNSUPPORTED, "bk_dist: not built yet");
}
extern "C" int bk_dist_destroy(bk_dist*) { return BK_OK; }
extern "C" int bk_dist_spmv(bk_handle*, bk_dist*, const void*, void*, void*) {
  return bk_fail(BK_ERR_UNSUPPORTED, "bk_dist: not built yet");
}
extern "C" int bk_dist_cg(bk_handle*, bk_dist*, const void*, void*, int, double, double, int64_t, bk_result*, void*) {
  return bk_fail(BK_ERR_UNSUPPORTED, "bk_dist: not built yet");
}
