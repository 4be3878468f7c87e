// Internal definitions shared by the libudal translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "udal.h"

#define UDAL_NUM_SMS 148

struct udal_head_weights_dev {
  bool set = false;
  int cout = 0;
  float* dw = nullptr;    // [R][9][F]
  float* pw = nullptr;    // [R][F][F]
  float* bias = nullptr;  // [R][F]
  float* bn_scale = nullptr;  // [R][L][F]  gamma * rsqrt(var + eps)
  float* bn_shift = nullptr;  // [R][L][F]  beta - mean * scale
  float* dwp = nullptr;   // [9][F]
  float* pwp = nullptr;   // [F][cout]
  float* bp = nullptr;    // [cout]
  // bf16 tensor-core mode: per (repeat, level) folded pointwise weights, K-major
  void* pw_bf16 = nullptr;
  void* pwp_bf16 = nullptr;
  float* fold_bias = nullptr;
  int ig_rows = 0;       // weight rows per tap of the predict image in ig_w
  void* fused_w = nullptr;  // predict layer for the fused kernels: bf16 [9][fused_rows][64] image, then bias [80 | 96] fp32
  int fused_rows = 0;       // 72 (cout <= 72) or 96
  int pred_chunks = 1;   // > 1: predict layer with more than 80 channels, run as chunks of pred_chunk channels
  int pred_chunk = 0;
  // heads wider than 64 channels (heads_wide.cu): channels zero-padded to 128
  void* wide_w = nullptr;   // bf16 pointwise images [R + wide_chunks][2 atoms][128 n][64 k]
  float* wide_f = nullptr;  // depthwise [R + 1][9][128], then epilogue (scale | bias) [R][L][2][128] + [wide_chunks][2][128]
  int wide_chunks = 0;      // predict layer as chunks of <= 128 channels
  // fp32-accurate tensor-core mode (heads_wide.cu, X3): fp16 (hi, lo) images [R + x3_chunks][2][64][64], epilogue tables
  void* x3_w = nullptr;
  float* x3_f = nullptr;
  int x3_chunks = 0;
  void* l0_w = nullptr;     // 64-channel towers, layer 0 through heads_wide_kernel<64, fp32 in>: bf16 [64 n][64 k] image
  float* l0_ep = nullptr;   // ... and its epilogue tables [L][2][64] (BN scale | folded bias)
  // fp16 mode (heads_dw.cu): depthwise on the CUDA cores, pointwise on tcgen05
  void* dwh_w = nullptr;    // fp16 images: tower layers 2..R-1 [64][64] | predict chunks [dwh_chunks][dwh_rows][64] | fused predict [dwh_frows][64]
  float* dwh_f = nullptr;   // predict bias per chunk [dwh_chunks][dwh_rows] | fused bias [dwh_frows] | ones [96]
  int dwh_chunks = 0, dwh_chunk = 0, dwh_rows = 0, dwh_frows = 0;
  void* ig_w = nullptr;  // implicit-GEMM weight images: [(R-2)*L tower layers >= 2][9][64][64] then predict [9][Npad][64], bf16
};

struct udal_scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};

struct udal_ctx {
  udal_config cfg;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  int64_t num_anchors = 0;      // N
  int64_t num_pixels = 0;       // P
  int64_t level_pix_off[UDAL_MAX_LEVELS + 1];  // prefix of H_l*W_l
  float* anchors = nullptr;     // device [N,4]
  bool anchors_set = false;
  udal_head_weights_dev heads[2];
  int64_t launches = 0;
  // named scratch slots, grown on demand; the post-processing slots exist twice (banks) so that the
  // NMS tail of one udal_run may still read them while the next run's decode kernel writes the other set
  udal_scratch scratch[2 * 20];
  int scratch_bank = 0;         // bank udal_scratch_get resolves banked slots to
  // udal_run pipelining: top-k / NMS / assemble of run i on post_stream overlap the heads of run i+1
  cudaStream_t post_stream = nullptr;
  cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_post[2] = {nullptr, nullptr};
  bool post_pending[2] = {false, false};
  int run_bank = 0;
  bool in_run = false;
  bool run_pipelined = false;   // this udal_run was issued while the previous run's tail was still executing
  // fused udal_run: called between the class head and the box head (the scores exist, the boxes do not yet)
  int (*between_heads)(udal_ctx*, void*) = nullptr;
  void* between_heads_arg = nullptr;
  std::vector<void*> user_allocs;
  // work-item counters of the persistent head kernels (dynamic item claiming): one zeroed int per launch of a run
  int* work_counters = nullptr;
  int work_counter_next = 0;
  // streaming front end (udal_stage_* / udal_fetch_*): uploads on a copy stream, results fetched behind the run's tail
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_staged[UDAL_STAGE_SLOTS] = {}, ev_consumed[UDAL_STAGE_SLOTS] = {}, ev_fetched[UDAL_STAGE_SLOTS] = {};
  bool consumed_pending[UDAL_STAGE_SLOTS] = {};
  cudaStream_t last_tail_stream = nullptr;   // where the last udal_run / postprocess call left its results (null: ctx->stream)
  bool feat_f16 = false;         // udal_set_feature_format: the feats pointers of udal_run / udal_heads_sample are fp16
  bool profile_layers = false;
  std::vector<cudaEvent_t> layer_events;  // pairs (start, stop) in launch order
};

enum {
  SCR_TOPK_HIST = 0,
  SCR_TOPK_CAND,
  SCR_TOPK_META,
  SCR_PRE_A,
  SCR_PRE_B,
  SCR_NMS_A,
  SCR_NMS_B,
  SCR_HEADS_A,
  SCR_HEADS_B,
  SCR_HEADS_C,
  SCR_MISC,
  SCR_LEVEL_PTRS,
  SCR_POST_A,
  SCR_POST_B,
  SCR_POST_C,
  SCR_POST_D,
  SCR_POST_C2,
  SCR_HEADS_D,
};

void udal_set_error(const char* fmt, ...);
int udal_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int udal_scratch_get(udal_ctx* ctx, int slot, size_t bytes, void** out);
// makes the context's stream wait for every udal_run tail still in flight on post_stream (no-op inside
// udal_run).  Every entry point that enqueues work or copies on the context's stream calls it first.
int udal_join(udal_ctx* ctx);

#define UDAL_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return udal_cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// development (udal_host_trace = 1): reports host-side gaps > 0.3 ms between consecutive kernel launches (a launch or an
// API call in between that blocked on the device)
extern int udal_host_trace;
void udal_host_trace_mark(const char* file, int line);

#define UDAL_CHECK_LAUNCH(ctx)                    \
  do {                                            \
    (ctx)->launches++;                            \
    if (udal_host_trace) udal_host_trace_mark(__FILE__, __LINE__); \
    UDAL_CUDA(cudaGetLastError());                \
  } while (0)

#define UDAL_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      udal_set_error(__VA_ARGS__);     \
      return UDAL_ERR_INVALID;         \
    }                                  \
  } while (0)

#define UDAL_TRY(call)            \
  do {                            \
    int s__ = (call);             \
    if (s__ != UDAL_OK) return s__; \
  } while (0)

// CTAs of a persistent (one CTA per SM) head kernel.  The kernels claim their work items from a global counter
// (heads_umma.cuh), so the top-k / NMS tail of the previous udal_run on the post stream may occupy SMs at any time: a CTA
// that starts late claims fewer items.  udal_run_reserved_sms > 0 additionally leaves SMs free for that tail when
// udal_run calls arrive back to back (the static-striding kernels needed it; kept as a switch).
extern int udal_run_reserved_sms;
extern int udal_run_overlap;
constexpr int UDAL_WORK_COUNTERS = 64;
int udal_work_counters_reset(udal_ctx* ctx);        // zeroes the counters on the context's stream (start of a head-sampler run)
int udal_work_counter(udal_ctx* ctx, int** out);    // the next zeroed counter

static inline int udal_persistent_grid(const udal_ctx* ctx, int items) {
  int sms = UDAL_NUM_SMS;
  if (ctx->in_run && ctx->run_pipelined && udal_run_overlap && udal_run_reserved_sms > 0 &&
      udal_run_reserved_sms < UDAL_NUM_SMS / 2)
    sms -= udal_run_reserved_sms;
  return items < sms ? items : sms;
}

// level pointer tables passed to kernels by value
struct udal_level_ptrs {
  const float* p[UDAL_MAX_LEVELS];
};
struct udal_level_geom {
  int num_levels;
  int h[UDAL_MAX_LEVELS];
  int w[UDAL_MAX_LEVELS];
  int pix_off[UDAL_MAX_LEVELS + 1];  // prefix sums of h*w
};

static inline udal_level_geom udal_geom(const udal_ctx* ctx) {
  udal_level_geom g;
  g.num_levels = ctx->cfg.num_levels;
  for (int l = 0; l < UDAL_MAX_LEVELS; ++l) {
    g.h[l] = l < g.num_levels ? ctx->cfg.level_h[l] : 0;
    g.w[l] = l < g.num_levels ? ctx->cfg.level_w[l] : 0;
  }
  for (int l = 0; l <= UDAL_MAX_LEVELS; ++l) g.pix_off[l] = (int)ctx->level_pix_off[l < g.num_levels ? l : g.num_levels];
  return g;
}

static inline int udal_box_channels(const udal_ctx* ctx) {
  return 4 * ctx->cfg.anchors_per_loc * (ctx->cfg.loss_attenuation ? 2 : 1);
}

// orderable key: larger float <=> larger unsigned
__host__ __device__ static inline uint32_t udal_float_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ static inline float udal_key_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// internal entry points (defined across translation units)
int udal_launch_decode_moments(udal_ctx* ctx, const float* const* cls, const float* const* box,
                               int batch, const udal_prenms_out* out);
int udal_launch_logit_moments(udal_ctx* ctx, const float* const* cls, int batch, float* mean_logits,
                              float* std_logits);
int udal_launch_decode_gather(udal_ctx* ctx, const float* const* box, int batch, int k,
                              const int32_t* topk_idx, const float* topk_val, const float* std_logits,
                              const udal_prenms_topk_out* out);
int udal_launch_topk(udal_ctx* ctx, const float* values, int batch, int64_t m, int k,
                     int32_t* idx_out, float* val_out);
int udal_launch_nms_v5(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n,
                       int32_t* sel_idx, float* sel_scores, int32_t* valid);
int udal_nms_sorted(udal_ctx* ctx, const float* boxes, const float* scores, const int32_t* cand_idx,
                    const int32_t* seg_start, const int32_t* seg_count, const float* next_score,
                    int next_stride, int segments, int seg_n, int segs_per_image, int64_t img_stride,
                    int64_t total_cand, int32_t* sel_row, int32_t* sel_rank, float* sel_scores,
                    int32_t* valid, int32_t* flag);
int udal_nms_full(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n,
                  const int32_t* flag, int32_t* sel_row, float* sel_scores, int32_t* valid);
int udal_nms_prefilter_k(const udal_ctx* ctx, int n);
extern int udal_nms_cta;
int udal_nms_epoch_segments(udal_ctx* ctx, const float* boxes, const float* scores, const int32_t* cand_idx, const int32_t* seg_start,
                            const int32_t* seg_count, int segments, int segs_per_image, int64_t img_stride, int32_t* sel_row,
                            int32_t* sel_rank, float* sel_scores, int32_t* valid, int* handled);
int udal_nms_epoch(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n, int32_t* sel_idx,
                   float* sel_scores, int32_t* valid);
// top-K pre-filter of a global NMS (scratch of the context's current bank) and the selection that consumes it
struct udal_nms_plan {
  int kk = 0, kq = 0;
  int32_t* tk_idx = nullptr;
  float* tk_val = nullptr;
  int32_t* flag = nullptr;
  int32_t* starts = nullptr;
};
int udal_nms_prefilter(udal_ctx* ctx, const float* scores, int segments, int n, udal_nms_plan* plan);
int udal_nms_select(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n, const udal_nms_plan& plan,
                    int32_t* sel_idx, float* sel_scores, int32_t* valid);
