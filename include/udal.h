/* libudal - C ABI of the B200-native uncertainty-sampling / post-processing hot path.
 *
 * Drop-in boundary for ONE path of continental/uncertainty-detection-autolabeling:
 *   BiFPN features -> T x class/box(+sigma) heads -> anchor decode with exact moment propagation
 *   -> MC mean/std -> top-k | max-reduce -> NMS -> detections.
 *
 * The reference has no FFI of its own (pure Python on TensorFlow); each entry point below cites
 * the reference function(s) (file:line under the reference's src/) whose arithmetic it replaces.
 * The Python mirror of the reference's modules (postprocess.py / anchors.py / nms_np.py /
 * utils_box.py / utils_extra.py) binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no framework types.  "device pointer" = CUDA global memory
 *     on the context's device (a DLPack / __cuda_array_interface__ data pointer passes as is).
 *   - every function returns 0 on success, a negative udal_status otherwise;
 *     udal_last_error() returns a thread-local message for the last failure.
 *   - all work is enqueued on the context's stream; functions return without synchronising
 *     unless stated.  One context per GPU, one host thread per context.
 *   - the library never frees caller memory and keeps no global state.
 *   - tensors are C-contiguous fp32, NHWC, levels ordered min_level..max_level.
 */
#ifndef UDAL_H_
#define UDAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDAL_MAX_LEVELS 8
#define UDAL_ABI_VERSION 1

typedef enum {
  UDAL_OK = 0,
  UDAL_ERR_INVALID = -1, /* bad argument / unsupported configuration (Python: ValueError) */
  UDAL_ERR_CUDA = -2,    /* CUDA runtime failure (Python: RuntimeError)                    */
  UDAL_ERR_STATE = -3,   /* call order (weights / anchors not set)                          */
  UDAL_ERR_NOMEM = -4
} udal_status;

/* decode method, reference utils_box.py:105-276 (decode_uncert) */
enum { UDAL_DECODE_LNORM = 0, UDAL_DECODE_NFLOW = 1, UDAL_DECODE_FALSEDEC = 2 };
/* arithmetic of the stand-alone decode + MC-moments kernel (udal_decode_moments and the postprocess entry points built on it):
 * FP64 reproduces the reference's NumPy/TF float64 decode value for value (utils_box.py:105-276); FP32 is the closed form in
 * fp32 (ex2.approx, series for exp(v) - 1, one-pass shifted variances) - the 1e-4 relative contract of BASELINE.json at the
 * speed of HBM.  udal_run's fused predict + decode kernels always use the fp32 form. */
enum { UDAL_DECODE_FP64 = 0, UDAL_DECODE_FP32 = 1 };
/* NMS method, reference postprocess.py:373-388 */
enum { UDAL_NMS_HARD = 0, UDAL_NMS_GAUSSIAN = 1 };
/* head-GEMM arithmetic */
enum {
  UDAL_HEADS_FP32 = 0,    /* CUDA-core fp32 towers: the 1e-4 parity mode */
  UDAL_HEADS_BF16_TC = 1, /* tcgen05, bf16 operands / activations (implicit-GEMM kernels) */
  UDAL_HEADS_FP16_TC = 2, /* tcgen05 pointwise + packed-fp16 depthwise, fp16 operands / activations: the precision the
                             reference's own GPU export runs in (mixed_float16, infer_lib.py:429-431); the benchmarked mode */
  UDAL_HEADS_FP32X3_TC = 3 /* fp32-accurate on the tensor cores (64-channel towers): fp32 depthwise on the CUDA cores, the
                              pointwise GEMM as three fp16 tcgen05 passes over (hi, lo) operand pairs with fp32 accumulation,
                              fp32 activations: the 1e-4 contract of UDAL_HEADS_FP32 at several times its speed */
};
enum { UDAL_HEAD_CLASS = 0, UDAL_HEAD_BOX = 1 };

/* The hot-path slice of the reference's params dict (hparams_config.py:183-370) plus geometry. */
typedef struct {
  int32_t abi_version;              /* UDAL_ABI_VERSION */
  int32_t device;                   /* CUDA ordinal */
  int32_t image_h, image_w;         /* utils.parse_image_size(params["image_size"]) */
  int32_t num_levels;               /* max_level - min_level + 1 */
  int32_t level_h[UDAL_MAX_LEVELS]; /* utils.get_feat_sizes(...)[min_level..max_level] */
  int32_t level_w[UDAL_MAX_LEVELS];
  int32_t anchors_per_loc;          /* A = num_scales * len(aspect_ratios) */
  int32_t num_classes;              /* C */
  int32_t num_filters;              /* F = fpn_num_filters */
  int32_t repeats;                  /* R = box_class_repeats */
  int32_t mc_samples;               /* T = mc_dropoutsamp */
  int32_t loss_attenuation;         /* box head emits 8A channels (4A box | 4A sigma) */
  int32_t cls_mc;                   /* bool(mc_classheadrate or mc_dropoutrate) */
  int32_t box_mc;                   /* bool(mc_boxheadrate or mc_dropoutrate) */
  float rate_class, rate_box;       /* SpatialDropout2D rates of the two heads */
  int32_t decode_method;            /* UDAL_DECODE_* */
  int32_t nms_method;               /* UDAL_NMS_* */
  float nms_iou_thresh;             /* after the reference's defaulting (postprocess.py:380-386) */
  float nms_score_thresh;           /* -INFINITY allowed */
  float nms_sigma_tf;               /* soft_nms_sigma passed to TF = sigma / 2; 0 for hard */
  int32_t nms_variant_old;          /* 0: TF >= 2.4 kernel semantics (pinned 2.10), 1: TF <= 2.3 */
  int32_t max_nms_inputs;           /* 0: max-reduce (serving), >0: top-k (eval) */
  int32_t max_output_size;          /* detections per image (100) */
  int32_t heads_mode;               /* UDAL_HEADS_* */
  int32_t prefilter_k;              /* global soft-NMS candidate pre-filter (0 = library default) */
  float inv_keep_class, inv_keep_box; /* fp32(1 / (1 - rate)) computed in double by the host, as TF does */
  int32_t decode_precision;         /* UDAL_DECODE_FP64 (default 0) | UDAL_DECODE_FP32: arithmetic of the stand-alone K2 kernel */
  int32_t reserved[4];
} udal_config;

typedef struct udal_ctx udal_ctx;

const char* udal_last_error(void);
int udal_abi_version(void);

/* ---- context, memory, stream ------------------------------------------------------------ */
int udal_create(const udal_config* cfg, udal_ctx** out);
int udal_destroy(udal_ctx* ctx);
/* use an external CUDA stream (cudaStream_t as void*); NULL restores the context's own stream */
int udal_set_stream(udal_ctx* ctx, void* cuda_stream);
int udal_sync(udal_ctx* ctx);
/* element type of the BiFPN feature maps handed to udal_heads_sample / udal_run / udal_run_prenms (their `feats` pointers are
 * typed float for the default): UDAL_FEAT_F16 = IEEE half [B,H_l,W_l,F], what the reference's GPU graphs produce under the
 * `mixed_precision` / mixed_float16 policy (src/hparams_config.py `mixed_precision`, src/utils_keras.py:142-160).  Needs
 * heads_mode UDAL_HEADS_FP16_TC and F = 64: the layer-0 kernel stages the fp16 tiles with TMA directly (half the bytes over
 * PCIe and HBM).  Sticky until changed. */
enum { UDAL_FEAT_F32 = 0, UDAL_FEAT_F16 = 1 };
int udal_set_feature_format(udal_ctx* ctx, int format);
/* ---- streaming front end ------------------------------------------------------------------
 * Keeps the GPU-side sequence of back-to-back udal_run calls (heads of run i+1 over the NMS tail of run i) when inputs
 * come from HOST buffers: uploads run on a copy stream into one of UDAL_STAGE_SLOTS caller-owned device buffer sets,
 * results are fetched behind the tail of their own run without joining it into the context's stream.
 *   udal_stage_begin(slot)      the copy stream waits until the run that last consumed `slot` is through its inputs
 *   udal_stage_h2d(...)         cudaMemcpyAsync on the copy stream (pinned host memory for a truly asynchronous copy)
 *   udal_stage_end(slot)        marks the slot's uploads
 *   udal_stage_acquire(slot)    the context's stream waits for them             (before udal_run)
 *   udal_stage_release(slot)    records "inputs consumed" on the context's stream (after udal_run)
 *   udal_fetch_d2h(...)         device -> host copy ordered behind the results of the last udal_run / postprocess call
 *   udal_fetch_mark(slot) / udal_fetch_wait(slot)   completion of the fetches enqueued so far / the host waits for it */
#define UDAL_STAGE_SLOTS 4
int udal_stage_begin(udal_ctx* ctx, int slot);
int udal_stage_h2d(udal_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int udal_stage_end(udal_ctx* ctx, int slot);
int udal_stage_acquire(udal_ctx* ctx, int slot);
int udal_stage_release(udal_ctx* ctx, int slot);
int udal_fetch_d2h(udal_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int udal_fetch_mark(udal_ctx* ctx, int slot);
int udal_fetch_wait(udal_ctx* ctx, int slot);
/* the stream the context currently enqueues on (cudaStream_t as void*): what a DLPack producer is handed in
 * __dlpack__(stream=...) */
int udal_get_stream(udal_ctx* ctx, void** cuda_stream);
/* orders all later work of the context after everything enqueued so far on producer_stream (a cudaStream_t, or the
 * special handles 0 / 1 = legacy default stream, 2 = per-thread default stream): zero-copy inputs written by another
 * framework's stream (the context's own streams are non-blocking and do not synchronise with the default stream) */
int udal_wait_stream(udal_ctx* ctx, void* producer_stream);
/* the same between two contexts: all later work of `ctx` after everything `producer` has enqueued so far (its udal_run
 * tails included).  Device arrays of one context handed to an entry point of another one (a sampler's detections into the
 * auto-label pass of a different context) are ordered this way by the Python layer. */
int udal_wait_context(udal_ctx* ctx, udal_ctx* producer);
int udal_malloc(udal_ctx* ctx, size_t bytes, void** dev_ptr);
int udal_free(udal_ctx* ctx, void* dev_ptr);
int udal_host_alloc(size_t bytes, void** pinned_ptr);
int udal_host_free(void* pinned_ptr);
int udal_memcpy_h2d(udal_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int udal_memcpy_d2h(udal_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int udal_memcpy_d2d(udal_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
int udal_memset(udal_ctx* ctx, void* dst_dev, int value, size_t bytes);
/* CUDA-event stopwatch on the context's stream (bench.py) */
int udal_timer_start(udal_ctx* ctx);
int udal_timer_stop(udal_ctx* ctx, float* elapsed_ms); /* synchronises */
int udal_device_count(int* count);
int udal_num_anchors(const udal_ctx* ctx, int64_t* n);
/* kernels this context has launched since creation (bench.py "gpu_launches") */
int udal_launch_count(const udal_ctx* ctx, int64_t* n);

/* ---- anchors ---------------------------------------------------------------------------- */
/* anchors.py:158-215 (Anchors._generate_boxes): the host mirror builds the float64 table the way
 * the reference does and uploads its float32 cast, [N,4] = (ymin, xmin, ymax, xmax). */
int udal_set_anchors(udal_ctx* ctx, const float* anchors_host, int64_t num_anchors);

/* ---- head sampler ----------------------------------------------------------------------- */
/* efficientdet_keras.py:353-513 (ClassNet), 516-692 (BoxNet); host pointers, copied.
 *   dw   [R][3][3][F]      depthwise kernels of the tower (shared over levels)
 *   pw   [R][F][F]         pointwise kernels   bias [R][F]
 *   bn_* [R][L][F]         per (repeat, level) gamma / beta / moving mean / moving variance
 *   dwp  [3][3][F], pwp [F][Cout], bp [Cout]   the predict layer (Cout = A*C or 4A / 8A)   */
int udal_set_head_weights(udal_ctx* ctx, int head, const float* dw, const float* pw,
                          const float* bias, const float* bn_gamma, const float* bn_beta,
                          const float* bn_mean, const float* bn_var, const float* dwp,
                          const float* pwp, const float* bp);

/* efficientdet_keras.py:979-1050 (MC loop) + utils_extra.py:201-217 (stack_mcpred), starting at
 * the BiFPN outputs.  feats[l]: device [B,H_l,W_l,F].  keep_masks: device uint8
 * [T,2,L,R,B,F] (1 = keep) or NULL to draw them in-kernel from Philox4x32-10 keyed by
 * (seed; t, head, level, repeat, b, f).  cls_out[l]: device [T,B,H_l,W_l,A*C];
 * box_out[l]: device [T,B,H_l,W_l,4A|8A].  A head without MC dropout writes T identical
 * samples only when T-stacking was requested by cfg (cls_mc / box_mc), else [B,...]. */
int udal_heads_sample(udal_ctx* ctx, const float* const* feats, int batch,
                      const uint8_t* keep_masks, uint64_t seed, float* const* cls_out,
                      float* const* box_out);

/* ---- pre-NMS: MC moments + decode ------------------------------------------------------- */
/* Per-anchor outputs of the serving variant (max_nms_inputs == 0), all device pointers;
 * any pointer may be NULL to skip that output.
 *   utils_extra.py:220-244 (get_mcuncert)  -> mean_logits, std_logits  [B,N,C]
 *   utils_box.py:105-276 (decode_uncert) / anchors.py:41-75 + postprocess.py:297-331
 *                                          -> boxes, albox, mcbox     [B,N,4]
 *   postprocess.py:123-135, 284            -> scores = sigmoid(max_c), classes = argmax_c [B,N] */
typedef struct {
  float* mean_logits;
  float* std_logits;
  float* boxes;
  float* albox;
  float* mcbox;
  float* scores;
  int32_t* classes;
} udal_prenms_out;

/* cls[l]: device [T,B,H,W,A*C] if cfg.cls_mc else [B,H,W,A*C]; box[l]: device
 * [T,B,H,W,4A|8A] if cfg.box_mc else [B,H,W,4A|8A].  postprocess.py:144-339 (pre_nms) without
 * top-k. */
int udal_decode_moments(udal_ctx* ctx, const float* const* cls, const float* const* box,
                        int batch, const udal_prenms_out* out);

/* postprocess.py:90-121 (topk_class_boxes, max_nms_inputs > 0): k largest of values[B,M] in
 * canonical order (value descending, index ascending).  idx_out [B,k] int32, val_out [B,k]. */
int udal_topk(udal_ctx* ctx, const float* values, int batch, int64_t m, int k, int32_t* idx_out,
              float* val_out);

/* pre_nms with top-k (postprocess.py:212-282, 297-331): mean logits -> top-k -> gather ->
 * decode -> moments on the k selected (anchor, class) pairs.  Outputs device [B,k,...]. */
typedef struct {
  float* mean_logits; /* [B,N,C]  classes_multi (un-gathered, postprocess.py:210-211) */
  int32_t* topk_idx;  /* [B,k]    flat index into N*C */
  float* boxes;       /* [B,k,4] */
  float* albox;       /* [B,k,4] or NULL */
  float* mcbox;       /* [B,k,4] or NULL */
  float* mcclass;     /* [B,k]   std of the selected logit, or NULL */
  float* scores;      /* [B,k] */
  int32_t* classes;   /* [B,k] */
} udal_prenms_topk_out;
int udal_prenms_topk(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                     const udal_prenms_topk_out* out);

/* ---- NMS -------------------------------------------------------------------------------- */
/* tf.raw_ops.NonMaxSuppressionV5 as called at postprocess.py:392-400, batched over independent
 * segments: boxes [S,n,4], scores [S,n] (device).  Thresholds come from cfg.  sel_idx [S,max_out]
 * int32 and sel_scores [S,max_out] are zero padded, valid [S] int32. */
int udal_nms_v5(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n,
                int32_t* sel_idx, float* sel_scores, int32_t* valid);

/* ---- BiFPN primitives (efficientdet_keras.py:51-350, 766-847) ---------------------------------
 * The node graph (fpn_configs.bifpn_config) is host logic - bifpn.py mirrors FNode / FPNCell / FPNCells on top of
 * these three calls.  All tensors device fp32 NHWC.
 * udal_conv1x1_bn : ResampleFeatureMap._maybe_apply_1x1 (305-311): in [NB,H,W,Cin] @ w [Cin,F] + bias [F], then
 *                   v * bn_scale + bn_shift when the tables are given (apply_bn_for_resampling).
 * udal_bifpn_fuse : FNode.call / fuse_features (86-125, 166-173): out [NB,H,W,F] = act(sum_i weight_i *
 *                   resample_i(in[i] [NB,in_h[i],in_w[i],F])); resampling by size as ResampleFeatureMap.call (313-350):
 *                   larger source = max (pool_avg: average) pooling with stride s = (in-1)/out+1, window s+1, SAME
 *                   padding; smaller = nearest neighbour; mode UDAL_FUSE_* with edge weights wsm[i] (scalar, or [F]
 *                   with per_channel); act != 0 applies swish to the fused value (OpAfterCombine, 229-236, activates
 *                   before its conv when conv_bn_act_pattern is off - the reference default).
 * udal_sepconv_bn : OpAfterCombine's SeparableConv2D + BN (200-236): depthwise 3x3 SAME [9][F] -> pointwise [F][Cout]
 *                   + bias -> act (UDAL_ACT_*). */
enum { UDAL_FUSE_SUM = 0, UDAL_FUSE_FASTATTN = 1, UDAL_FUSE_ATTN = 2 };
enum { UDAL_ACT_NONE = 0, UDAL_ACT_BN_SWISH = 1, UDAL_ACT_BN = 2 };
int udal_conv1x1_bn(udal_ctx* ctx, const float* in, int NB, int H, int W, int Cin, const float* w, const float* bias,
                    const float* bn_scale, const float* bn_shift, int F, float* out);
int udal_bifpn_fuse(udal_ctx* ctx, int n, const float* const* in, const int* in_h, const int* in_w, const float* const* wsm,
                    int mode, int per_channel, int pool_avg, int NB, int H, int W, int F, int act, float* out);
int udal_sepconv_bn(udal_ctx* ctx, const float* in, int NB, int H, int W, int F, int Cout, const float* dw, const float* pw,
                    const float* bias, const float* bn_scale, const float* bn_shift, int act, float* out);
/* The same separable conv + BN (+ swish) of a 64-channel node (F = Cout = 64: the D0 BiFPN) on the tensor cores, fp32 accurate:
 * fp32 depthwise on the CUDA cores, the pointwise GEMM as three fp16 tcgen05 passes over (hi, lo) operand pairs (the fp32x3
 * tower kernel on one map).  udal_sepconv_tc_prepare builds the device tables of one conv once (pointwise hi / lo images,
 * BN scale | folded bias; freed with the context); udal_sepconv_tc runs it: in / out [NB,H,W,64] fp32, 16-byte aligned,
 * dw [9][64], act = UDAL_ACT_BN | UDAL_ACT_BN_SWISH. */
int udal_sepconv_tc_prepare(udal_ctx* ctx, const float* pw, const float* bias, const float* bn_scale, const float* bn_shift,
                            void** table);
int udal_sepconv_tc(udal_ctx* ctx, const float* in, int NB, int H, int W, const float* dw, const void* table, int act, float* out);

/* ---- fused post-processing -------------------------------------------------------------- */
/* postprocess.py:472-621 (postprocess_global), serving variant.  image_scales: device [B] or
 * NULL.  Outputs (device): boxes [B,max_out,4*(1+has_al+has_mc)] = box|albox|mcbox,
 * scores [B,max_out], classes [B,max_out,1+C*has_mcclass] = class|mcclass-std, valid [B] int32,
 * logits [B,max_out,C]. */
typedef struct {
  float* boxes;
  float* scores;
  float* classes;
  int32_t* valid;
  float* logits;
} udal_detections;
int udal_postprocess_global(udal_ctx* ctx, const float* const* cls, const float* const* box,
                            int batch, const float* image_scales, const udal_detections* out);

/* postprocess.py:719-740 (postprocess_per_class), eval variant (cfg.max_nms_inputs > 0).
 * Outputs: boxes [B,max_out,4], scores, classes [B,max_out], valid [B], logits [B,max_out,C].
 * strict_reference != 0 reproduces the reference's logits chain (postprocess.py:659-666). */
int udal_postprocess_per_class(udal_ctx* ctx, const float* const* cls, const float* const* box,
                               int batch, const float* image_scales, int strict_reference,
                               const udal_detections* out);

/* postprocess.py:624-716 (per_class_nms) on already decoded candidates: boxes [B,K,4], scores
 * [B,K], classes [B,K] int32 in the candidate order the reference would pass (top-k order).
 * logits: [B,logit_rows,C] or NULL.  strict_reference != 0: the reference's logits chain; else
 * logits rows are taken at the candidate position. */
int udal_per_class_nms(udal_ctx* ctx, const float* boxes, const float* scores,
                       const int32_t* classes, int batch, int k, const float* image_scales,
                       const float* logits, int64_t logit_rows, int strict_reference,
                       const udal_detections* out);

/* postprocess.py:743-785 (generate_detections_from_nms_output): rows
 * [id, x1, y1, x2, y2, score, class, logits...] from [B,mo,box_stride] boxes (first 4 columns
 * ymin,xmin,ymax,xmax), scores [B,mo], classes [B,mo*class_stride] (first column), image ids [B],
 * original widths [B] (flip only), logits [B,mo,nlogits] or NULL.  out [B,mo,7+nlogits]. */
int udal_format_detections(udal_ctx* ctx, const float* boxes, int box_stride, const float* scores,
                           const float* classes, int class_stride, const float* image_ids,
                           const float* widths, int flip, const float* logits, int nlogits,
                           int batch, int max_out, float* out);
/* postprocess.py:874-887 (transform_detections): [id,x1,y1,x2,y2,score,class,...] ->
 * [id,x,y,w,h,score,class]; in [rows,in_cols], out [rows,7]. */
int udal_transform_detections(udal_ctx* ctx, const float* in, int64_t rows, int in_cols, float* out);

/* ---- calibrated uncertainties + auto-label threshold pass ------------------- */
enum {
  UDAL_CALIB_NONE = 0,
  UDAL_CALIB_TS_ALL = 1,           /* utils_box.py:424-426  uncert / temps[0] */
  UDAL_CALIB_TS_PERCOO = 2,        /* :441-452              uncert[:, j] / temps[j] */
  UDAL_CALIB_ISO_ALL = 3,          /* :420-422              table 0 */
  UDAL_CALIB_ISO_PERCOO = 4,       /* :428-439              table j */
  UDAL_CALIB_ISO_PERCLSCOO = 5,    /* :454-466              table (class - 1) * 4 + j */
  UDAL_CALIB_REL_ISO_PERCLSCOO = 6 /* :468-494              same on float16(uncert / [h,w,h,w]), times the norm */
};
typedef struct {
  int32_t calib_method_box;
  int32_t num_tables;      /* isotonic tables (sklearn X_thresholds_ / y_thresholds_), concatenated: */
  const float* table_x;    /*   knots of table t at [table_off[t], table_off[t+1]) - device pointers */
  const float* table_y;
  const int32_t* table_off;
  float temps[4];
  float class_temp;        /* logits / class_temp before the softmax (1 = uncalibrated; CalibrateClass ts_all) */
  float w_entropy;         /* opt_params of the selected uncertainties ("ENT", "ALBOX"); 0 = not selected */
  float w_albox;
  float threshold;         /* mean(opt_thrs) */
  float min_score;
  int32_t strict_reference; /* 1: after calibration every row uses the FIRST detection's std (infer_model.py:688-691) */
} udal_autolabel_params;
/* Per image: entropy of softmax(logits) (infer_model.py:585-595), calibrated aleatoric std
 * (utils_box.py:404-524), relative std (utils_box.py:279-292), opt_uncert = w_entropy * entropy +
 * w_albox * mean(relative std), decision[b] = all(opt_uncert[score > min_score] < threshold)
 * (infer_model.py:742-764; 1 = label automatically, 0 = examine).  boxes [B,M,box_stride] with the box at
 * columns 0..3 and the aleatoric std at albox_col..+3 (-1: none); classes [B,M,class_stride], id in column 0.
 * All pointers device memory. */
int udal_autolabel(udal_ctx* ctx, const float* boxes, int box_stride, int albox_col, const float* scores,
                   const float* classes, int class_stride, const float* logits, int num_classes, int batch,
                   int max_out, const udal_autolabel_params* prm, float* entropy, float* calib_albox,
                   float* rel_albox, float* opt_uncert, int32_t* decision);

/* utils_class.py:116-187 (CalibrateClass._perform_class_calib, no MC class uncertainty): calibrated class probabilities and
 * their entropy for `rows` detections.  method UDAL_CLASSCAL_*: temperature scaling with one temperature (temps[1]) or one per
 * class (temps[C]); isotonic regression on the softmax probabilities with one table or one per class (knots tx / ty, table t =
 * [off[t], off[t+1]); sklearn IsotonicRegression.predict, out_of_bounds="clip"), renormalised.  All pointers device. */
enum { UDAL_CLASSCAL_TS_ALL = 0, UDAL_CLASSCAL_TS_PERCLS = 1, UDAL_CLASSCAL_ISO_ALL = 2, UDAL_CLASSCAL_ISO_PERCLS = 3 };
int udal_calibrate_class(udal_ctx* ctx, const float* logits, long long rows, int C, int method, const float* temps,
                         const float* tx, const float* ty, const int32_t* off, float* probab, float* entropy);

/* layout helpers used by the Python mirror (device pointers):
 * out[r] = a[r] ++ b[r] for r < rows (tf.concat on the last axis) */
int udal_concat_channels(udal_ctx* ctx, const float* a, int ca, const float* b, int cb,
                         int64_t rows, float* out);
/* out[b][j][:] = src[b][idx[b][j]][:] (tf.gather / gather_nd batch_dims=1); rows out of range
 * read as zero (TF GPU rule).  mode 0: 32-bit copy; mode 1: int32 -> float32(value + 1)
 * (CLASS_OFFSET, postprocess.py:403). */
int udal_gather_rows(udal_ctx* ctx, const void* src, int batch, int64_t n_rows, int width,
                     const int32_t* idx, int m, int mode, void* out);

/* small device helpers behind the mirror's minor entry points (all pointers device memory):
 * postprocess.py:69-72 (clip_boxes): out = clip(boxes [rows,4], 0, [H,W,H,W]) */
int udal_clip_boxes(udal_ctx* ctx, const float* boxes, int64_t rows, float image_h, float image_w, float* out);
/* postprocess.py:123-135 (topk_class_boxes, max-reduce branch): reduce_max / argmax (first maximum) over the last axis */
int udal_max_reduce(udal_ctx* ctx, const float* x, int64_t rows, int c, float* max_out, int32_t* argmax_out);
/* postprocess.py:104-105: indices = idx // num_classes, classes = idx % num_classes */
int udal_divmod_i32(udal_ctx* ctx, const int32_t* idx, int64_t total, int d, int32_t* quot, int32_t* rem);
/* postprocess.py:284 (tf.math.sigmoid) elementwise */
int udal_sigmoid(udal_ctx* ctx, const float* x, int64_t total, float* y);

/* utils_box.py:162-184 (decode_uncert, method "sample"): pred / sigma / anchors [n,4] -> mean and population std of
 * the corners of n_samples decoded draws t + |sigma| z, in float64, rounded to fp32.  normals: device [n_samples,4,n]
 * standard-normal draws (injected, parity) or NULL: Philox4x32-10 + Box-Muller in-kernel (counter s*n + i, key seed). */
int udal_decode_sample(udal_ctx* ctx, const float* pred, const float* sigma, const float* anchors, int64_t n,
                       int n_samples, const float* normals, uint64_t seed, float* coords, float* stds);

/* heads + post-processing in one call: feats -> detections (variant by cfg.max_nms_inputs).
 * Replaces EfficientDetNet.call MC branch + postprocess_global / postprocess_per_class
 * (efficientdet_keras.py:999-1050, 1102-1116).  Asynchronous: outputs are valid after udal_sync (or any
 * later call on this context, all of which are ordered after it).  In the serving configuration (bf16
 * heads, A = 9, C = 7 or 8, loss attenuation, l-norm, MC dropout on both heads, global NMS) the predict layers
 * are fused with the decode / MC moments (no [T,...] head outputs in HBM; box quantities within ~1e-6
 * relative of the stand-alone fp64 decode), and the top-k / NMS / assemble tail runs on a second stream
 * underneath the next udal_run.  Debug switches exported as ints: udal_run_fused, udal_run_overlap. */
int udal_run(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
             uint64_t seed, const float* image_scales, const udal_detections* out);

/* The first half of udal_run on its own: feats -> per-anchor tensors (the max-reduce variant of pre_nms,
 * postprocess.py:144-339, applied to the T head samples of efficientdet_keras.py:999-1050), through exactly the
 * kernels udal_run launches for this configuration (fused predict + decode / MC moments where they cover it,
 * predict layers + udal_decode_moments otherwise).  Needs cfg.max_nms_inputs == 0; with the fused kernels every
 * pointer of `out` must be set.  The parity tests compare these tensors with the oracle anchor by anchor. */
int udal_run_prenms(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
                    uint64_t seed, const udal_prenms_out* out);

/* per-layer CUDA-event timing of the head sampler (bench.py roofline): enable, run
 * udal_heads_sample, then read ms[head * (R + 1) + layer]; *n = entries written (synchronises). */
int udal_profile_layers(udal_ctx* ctx, int enable);
int udal_get_layer_times(udal_ctx* ctx, float* ms, int cap, int* n);

/* bytes of device scratch the context currently holds (grows on demand, never shrinks) */
int udal_scratch_bytes(const udal_ctx* ctx, size_t* bytes);

/* ---- nms_np family ---------------------------------------------------------------------- */
/* nms_np.py:30-220 on the device: dets [n,5] = (x1,y1,x2,y2,score) host pointer in, kept rows
 * out in selection order; method: 0 hard, 1 diou, 2 linear, 3 gaussian.  Returns the number of
 * kept rows in *num_kept (synchronises). */
int udal_nms_np(udal_ctx* ctx, const float* dets_host, int n, int method, float iou_thresh,
                float sigma, float score_thresh, float* kept_host, int32_t* num_kept);

#ifdef __cplusplus
}
#endif
#endif /* UDAL_H_ */
