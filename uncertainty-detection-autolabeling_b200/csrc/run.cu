// udal_run: BiFPN features -> detections in one call (heads + post-processing).
#include "udal_common.cuh"

int udal_heads_fused_ok(const udal_ctx* ctx);
int udal_run_global_fused(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                          const float* image_scales, const udal_detections* out);

int udal_heads_sample_fused(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                            const udal_prenms_out* pre);

// udal_run with 16-bit tensor-core heads whose predict layers are not fused with K2 (wide towers, class counts without a fused
// instantiation): the stand-alone decode runs in the arithmetic the fused kernels use (fp32 closed form, HBM bound) whatever
// decode_precision says - the head outputs carry 16-bit rounding, reproducing the float64 decode digit for digit buys nothing.
// The fp32 / fp32x3 modes (the 1e-4 contract) follow the configuration.
namespace {
struct DecodePrecisionScope {
  udal_ctx* c;
  int saved;
  explicit DecodePrecisionScope(udal_ctx* x) : c(x), saved(x->cfg.decode_precision) {
    if (c->cfg.heads_mode == UDAL_HEADS_BF16_TC || c->cfg.heads_mode == UDAL_HEADS_FP16_TC) c->cfg.decode_precision = UDAL_DECODE_FP32;
  }
  ~DecodePrecisionScope() { c->cfg.decode_precision = saved; }
};
}  // namespace

// per-level head outputs in one scratch block, every level starting on a 16-byte boundary (odd channel
// counts such as 63 = 9 anchors x 7 classes would otherwise misalign the following levels)
static int head_output_scratch(udal_ctx* ctx, int batch, float** cls, float** box) {
  const udal_config& c = ctx->cfg;
  const int L = c.num_levels, T = c.mc_samples;
  const int ccls = c.anchors_per_loc * c.num_classes, cbox = udal_box_channels(ctx);
  const size_t tc = (size_t)(c.cls_mc ? T : 1) * batch, tb = (size_t)(c.box_mc ? T : 1) * batch;
  size_t off_cls[UDAL_MAX_LEVELS], off_box[UDAL_MAX_LEVELS], total = 0;
  for (int l = 0; l < L; ++l) {
    const size_t px = (size_t)c.level_h[l] * c.level_w[l];
    off_cls[l] = total;
    total += (tc * px * ccls + 3) & ~(size_t)3;
  }
  for (int l = 0; l < L; ++l) {
    const size_t px = (size_t)c.level_h[l] * c.level_w[l];
    off_box[l] = total;
    total += (tb * px * cbox + 3) & ~(size_t)3;
  }
  float* buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_PRE_A, total * sizeof(float), (void**)&buf));
  for (int l = 0; l < L; ++l) {
    cls[l] = buf + off_cls[l];
    box[l] = buf + off_box[l];
  }
  return UDAL_OK;
}

// features -> per-anchor tensors (pre_nms of postprocess.py:144-339 in the max-reduce variant, starting at the BiFPN
// outputs) through exactly the kernels udal_run launches for this configuration: fused predict + K2 where they cover
// it, predict layers + decode_moments otherwise.  The parity tests compare these tensors with the oracle anchor by anchor.
extern "C" int udal_run_prenms(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
                               uint64_t seed, const udal_prenms_out* out) {
  UDAL_REQUIRE(ctx && feats && out, "NULL argument");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(ctx->cfg.max_nms_inputs == 0, "udal_run_prenms: the max-reduce variant (max_nms_inputs == 0)");
  if (udal_heads_fused_ok(ctx)) {
    UDAL_REQUIRE(out->mean_logits && out->std_logits && out->boxes && out->albox && out->mcbox && out->scores && out->classes,
                 "udal_run_prenms: the fused kernels write every per-anchor tensor");
    return udal_heads_sample_fused(ctx, feats, batch, keep_masks, seed, out);
  }
  float* cls[UDAL_MAX_LEVELS];
  float* box[UDAL_MAX_LEVELS];
  UDAL_TRY(head_output_scratch(ctx, batch, cls, box));
  UDAL_TRY(udal_heads_sample(ctx, feats, batch, keep_masks, seed, cls, box));
  DecodePrecisionScope prec(ctx);
  return udal_launch_decode_moments(ctx, cls, box, batch, out);
}

extern "C" int udal_run(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
                        uint64_t seed, const float* image_scales, const udal_detections* out) {
  UDAL_REQUIRE(ctx && feats && out, "NULL argument");
  if (udal_host_trace) udal_host_trace_mark("udal_run entry", 0);
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  if (udal_host_trace) udal_host_trace_mark("udal_run after cudaSetDevice", 0);
  // Pipelining across calls: the tail of the previous run (top-k / NMS / assemble) may still be in flight
  // on the post stream.  This run's heads start right away; its decode kernel writes scratch bank
  // `run_bank`, whose previous user (two runs ago) is waited for first.
  struct RunScope {
    udal_ctx* c;
    explicit RunScope(udal_ctx* x) : c(x) { c->in_run = true; }
    ~RunScope() {
      c->in_run = false;
      c->scratch_bank = 0;
    }
  } scope(ctx);
  const udal_config& c = ctx->cfg;
  ctx->run_pipelined = false;
  for (int b = 0; b < 2; ++b)
    if (ctx->post_pending[b] && cudaEventQuery(ctx->ev_post[b]) == cudaErrorNotReady) ctx->run_pipelined = true;
  (void)cudaGetLastError();  // cudaErrorNotReady is not an error
  if (udal_heads_fused_ok(ctx)) {
    // serving configuration: the predict layers write the per-anchor statistics straight into this run's
    // scratch bank, so its previous user (the tail of two runs ago) is waited for before the heads start
    const int bank = ctx->run_bank;
    if (ctx->post_pending[bank]) {
      UDAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_post[bank], 0));
      ctx->post_pending[bank] = false;
    }
    ctx->scratch_bank = bank;
    ctx->run_bank = bank ^ 1;
    if (udal_host_trace) udal_host_trace_mark("udal_run after bank wait", bank);
    return udal_run_global_fused(ctx, feats, batch, keep_masks, seed, image_scales, out);
  }
  float* cls[UDAL_MAX_LEVELS];
  float* box[UDAL_MAX_LEVELS];
  UDAL_TRY(head_output_scratch(ctx, batch, cls, box));
  UDAL_TRY(udal_heads_sample(ctx, feats, batch, keep_masks, seed, cls, box));
  DecodePrecisionScope prec(ctx);
  const int bank = ctx->run_bank;
  if (ctx->post_pending[bank]) {
    UDAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_post[bank], 0));
    ctx->post_pending[bank] = false;
  }
  ctx->scratch_bank = bank;
  ctx->run_bank = bank ^ 1;
  if (c.max_nms_inputs > 0) {
    // the per-class variant stays on one stream: join the other bank's tail as well
    if (ctx->post_pending[bank ^ 1]) {
      UDAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_post[bank ^ 1], 0));
      ctx->post_pending[bank ^ 1] = false;
    }
    return udal_postprocess_per_class(ctx, cls, box, batch, image_scales, 0, out);
  }
  return udal_postprocess_global(ctx, cls, box, batch, image_scales, out);
}
