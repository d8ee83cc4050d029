// lowlevel.cu — batched low-level replans (placeholder until the kernel lands).
#include "common.cuh"
using namespace mrp;
extern "C" int mrp_lowlevel_batch(const mrp_map*, int, const int32_t*, int, const int32_t*, int,
                                  const int32_t*, int, const int32_t*, const int32_t*, int, int,
                                  int, const mrp_job*, int, const mrp_lowlevel_params*,
                                  mrp_path_info*, int32_t*, int32_t*) {
  return fail(MRP_ERR_UNSUPPORTED, "mrp_lowlevel_batch: not built yet");
}
