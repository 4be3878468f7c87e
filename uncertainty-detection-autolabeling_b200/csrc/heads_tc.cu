// K1 (bf16 tensor-core mode) - placeholder translation unit until the tcgen05 sampler lands.
#include "udal_common.cuh"

int udal_heads_tc_prepare(udal_ctx* ctx, int head) {
  (void)ctx;
  (void)head;
  udal_set_error("heads_mode UDAL_HEADS_BF16_TC is not available in this build");
  return UDAL_ERR_INVALID;
}

int udal_heads_tc_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale,
                         float* const* cls_out, float* const* box_out) {
  (void)ctx; (void)feats; (void)batch; (void)scale; (void)cls_out; (void)box_out;
  udal_set_error("heads_mode UDAL_HEADS_BF16_TC is not available in this build");
  return UDAL_ERR_INVALID;
}
