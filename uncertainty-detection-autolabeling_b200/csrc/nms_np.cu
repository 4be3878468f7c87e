// nms_np family on the device - placeholder until the kernels land.
#include "udal_common.cuh"

extern "C" int udal_nms_np(udal_ctx* ctx, const float* dets_host, int n, int method, float iou_thresh, float sigma,
                           float score_thresh, float* kept_host, int32_t* num_kept) {
  (void)ctx; (void)dets_host; (void)n; (void)method; (void)iou_thresh; (void)sigma; (void)score_thresh;
  (void)kept_host; (void)num_kept;
  udal_set_error("udal_nms_np is not available in this build");
  return UDAL_ERR_INVALID;
}
