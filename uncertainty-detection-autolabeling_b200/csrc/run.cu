// udal_run: BiFPN features -> detections in one call (heads + post-processing).
#include "udal_common.cuh"

extern "C" int udal_run(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
                        uint64_t seed, const float* image_scales, const udal_detections* out) {
  UDAL_REQUIRE(ctx && feats && out, "NULL argument");
  const udal_config& c = ctx->cfg;
  const int L = c.num_levels, T = c.mc_samples;
  const int ccls = c.anchors_per_loc * c.num_classes, cbox = udal_box_channels(ctx);
  const size_t P = (size_t)ctx->num_pixels;
  const size_t ncls = (size_t)(c.cls_mc ? T : 1) * batch * P * ccls;
  const size_t nbox = (size_t)(c.box_mc ? T : 1) * batch * P * cbox;
  float* buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_PRE_A, (ncls + nbox) * sizeof(float), (void**)&buf));
  float* cls[UDAL_MAX_LEVELS];
  float* box[UDAL_MAX_LEVELS];
  for (int l = 0; l < L; ++l) {
    cls[l] = buf + (size_t)(c.cls_mc ? T : 1) * batch * ctx->level_pix_off[l] * ccls;
    box[l] = buf + ncls + (size_t)(c.box_mc ? T : 1) * batch * ctx->level_pix_off[l] * cbox;
  }
  UDAL_TRY(udal_heads_sample(ctx, feats, batch, keep_masks, seed, cls, box));
  if (c.max_nms_inputs > 0) return udal_postprocess_per_class(ctx, cls, box, batch, image_scales, 0, out);
  return udal_postprocess_global(ctx, cls, box, batch, image_scales, out);
}
