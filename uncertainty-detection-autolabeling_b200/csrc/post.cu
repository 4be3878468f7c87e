// Fused post-processing entry points: postprocess_global (serving), pre_nms with top-k and
// postprocess_per_class (eval).  Reference: src/postprocess.py:472-621, 144-339, 624-740.
#include <math_constants.h>

#include "udal_common.cuh"

namespace {

// ---- serving variant: gather the selected anchors into the detection tensors -----------------
struct AssembleParams {
  int batch, max_out, C;
  int64_t N;
  int has_al, has_mc, has_mcclass;
  float img_h, img_w;
  const int32_t* sel_idx;   // [B,max_out] anchor rows (0 padded)
  const float* sel_scores;  // [B,max_out]
  const int32_t* valid;     // [B]
  const float* boxes;       // [B,N,4]
  const float* albox;
  const float* mcbox;
  const int32_t* classes;   // [B,N]
  const float* mean_logits; // [B,N,C]
  const float* std_logits;  // [B,N,C]
  const float* scales;      // [B] or null
  udal_detections out;
};

__global__ void assemble_global_kernel(const AssembleParams p) {
  const int b = blockIdx.x;
  const float scale = p.scales ? p.scales[b] : 1.f;
  const int nb = 1 + p.has_al + p.has_mc;
  const int cw = 1 + (p.has_mcclass ? p.C : 0);
  for (int i = threadIdx.x; i < p.max_out; i += blockDim.x) {
    const size_t o = (size_t)b * p.max_out + i;
    const int64_t row = (int64_t)b * p.N + p.sel_idx[o];
    // postprocess.py:599-609: clip to the image, then scale boxes and box uncertainties
    float4 bx = reinterpret_cast<const float4*>(p.boxes)[row];
    bx.x = fminf(fmaxf(bx.x, 0.f), p.img_h);
    bx.y = fminf(fmaxf(bx.y, 0.f), p.img_w);
    bx.z = fminf(fmaxf(bx.z, 0.f), p.img_h);
    bx.w = fminf(fmaxf(bx.w, 0.f), p.img_w);
    float* ob = p.out.boxes + o * 4 * nb;
    if (p.scales) {
      bx.x = __fmul_rn(bx.x, scale);
      bx.y = __fmul_rn(bx.y, scale);
      bx.z = __fmul_rn(bx.z, scale);
      bx.w = __fmul_rn(bx.w, scale);
    }
    ob[0] = bx.x;
    ob[1] = bx.y;
    ob[2] = bx.z;
    ob[3] = bx.w;
    int col = 4;
    if (p.has_al) {
      const float4 v = reinterpret_cast<const float4*>(p.albox)[row];
      ob[col + 0] = p.scales ? __fmul_rn(v.x, scale) : v.x;
      ob[col + 1] = p.scales ? __fmul_rn(v.y, scale) : v.y;
      ob[col + 2] = p.scales ? __fmul_rn(v.z, scale) : v.z;
      ob[col + 3] = p.scales ? __fmul_rn(v.w, scale) : v.w;
      col += 4;
    }
    if (p.has_mc) {
      const float4 v = reinterpret_cast<const float4*>(p.mcbox)[row];
      ob[col + 0] = p.scales ? __fmul_rn(v.x, scale) : v.x;
      ob[col + 1] = p.scales ? __fmul_rn(v.y, scale) : v.y;
      ob[col + 2] = p.scales ? __fmul_rn(v.z, scale) : v.z;
      ob[col + 3] = p.scales ? __fmul_rn(v.w, scale) : v.w;
    }
    p.out.scores[o] = p.sel_scores[o];
    float* oc = p.out.classes + o * cw;
    oc[0] = (float)(p.classes[row] + 1);  // CLASS_OFFSET, postprocess.py:35, 403
    if (p.has_mcclass)
      for (int c = 0; c < p.C; ++c) oc[1 + c] = p.std_logits[row * p.C + c];
    if (p.out.logits)
      for (int c = 0; c < p.C; ++c) p.out.logits[o * p.C + c] = p.mean_logits[row * p.C + c];
  }
  if (threadIdx.x == 0) p.out.valid[b] = p.valid[b];
}

// ---- eval variant: split the k sorted candidates of an image into per-class segments ----------
__global__ void __launch_bounds__(256) class_partition_kernel(const int32_t* __restrict__ classes, int k, int C,
                                                              int32_t* __restrict__ seg_idx,
                                                              int32_t* __restrict__ seg_start,
                                                              int32_t* __restrict__ seg_count) {
  extern __shared__ int sh[];  // cnt[C], off[C]
  int* cnt = sh;
  int* off = sh + C;
  const int b = blockIdx.x;
  const int32_t* cls = classes + (size_t)b * k;
  for (int c = threadIdx.x; c < C; c += blockDim.x) cnt[c] = 0;
  __syncthreads();
  // rows whose class id is outside [0, C) belong to no segment: the reference loops over range(num_classes)
  // (postprocess.py:655-657) and never selects them
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int c = cls[j];
    if (c >= 0 && c < C) atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0;
    for (int c = 0; c < C; ++c) {
      off[c] = a;
      a += cnt[c];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = warp; c < C; c += nw) {
    int pos = off[c];
    if (lane == 0) {
      seg_start[b * C + c] = b * k + pos;
      seg_count[b * C + c] = cnt[c];
    }
    if (cnt[c] == 0) continue;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      const bool m = j < k && cls[j] == c;
      const unsigned int mask = __ballot_sync(0xffffffffu, m);
      if (m) seg_idx[(size_t)b * k + pos + __popc(mask & ((1u << lane) - 1u))] = j;
      pos += __popc(mask);
    }
  }
}

// ---- eval variant: concatenate the per-class results, zero pad, top max_out (sorted) ----------
struct MergeParams {
  int batch, k, C, max_out, p2, strict;
  int64_t N;
  const int32_t* sel_row;    // [B*C,max_out] position in the k-array
  const int32_t* sel_rank;   // [B*C,max_out] rank inside the class segment
  const float* sel_scores;   // [B*C,max_out]
  const int32_t* valid;      // [B*C]
  const int32_t* seg_count;  // [B*C]
  const float* dec_boxes;    // [B,k,4]
  const int32_t* topk_idx;   // [B,k] flat (anchor*C + class)
  const float* mean_logits;  // [B,N,C]
  const float* scales;
  float* chain;              // [B, C*max_out, C] scratch (strict logits chain)
  udal_detections out;
};

__global__ void __launch_bounds__(256) merge_per_class_kernel(const MergeParams p) {
  extern __shared__ unsigned long long keys[];  // [p2]
  __shared__ int cat_off[1025];
  const int b = blockIdx.x;
  const int C = p.C, mo = p.max_out;
  const int32_t* valid = p.valid + (size_t)b * C;
  if (threadIdx.x == 0) {
    int a = 0;
    for (int c = 0; c < C; ++c) {
      cat_off[c] = a;
      a += valid[c];
    }
    cat_off[C] = a;
  }
  __syncthreads();
  const int total_valid = cat_off[C];
  const int total = total_valid + mo;  // + max_out zero rows (postprocess.py:677-683)
  // entry id e -> (class, slot): stored in the low bits through the concat position
  for (int i = threadIdx.x; i < p.p2; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  for (int c = 0; c < C; ++c) {
    const int v = valid[c];
    for (int i = threadIdx.x; i < v; i += blockDim.x) {
      const float sc = p.sel_scores[((size_t)b * C + c) * mo + i];
      const unsigned int pos = (unsigned int)(cat_off[c] + i);
      keys[pos] = ((unsigned long long)udal_float_key(sc + 0.0f) << 32) | (unsigned long long)(~pos);
    }
  }
  for (int i = threadIdx.x; i < mo; i += blockDim.x) {
    const unsigned int pos = (unsigned int)(total_valid + i);
    keys[pos] = ((unsigned long long)udal_float_key(0.0f) << 32) | (unsigned long long)(~pos);
  }
  __syncthreads();
  for (unsigned int size = 2; size <= (unsigned int)p.p2; size <<= 1) {
    for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned int i = threadIdx.x; i < ((unsigned int)p.p2 >> 1); i += blockDim.x) {
        const unsigned int lo = 2 * i - (i & (stride - 1));
        const unsigned int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a < c) == desc) {
          keys[lo] = c;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  // strict logits chain (postprocess.py:659-666 under graph-mode / GPU-gather semantics)
  float* chain = p.chain ? p.chain + (size_t)b * C * mo * C : nullptr;
  if (p.out.logits && p.strict) {
    int prev_c = -1;
    for (int c = 0; c < C; ++c) {
      if (p.seg_count[(size_t)b * C + c] == 0) continue;
      const int v = valid[c];
      float* cur = chain + (size_t)c * mo * C;
      for (int t = threadIdx.x; t < v * C; t += blockDim.x) {
        const int e = t / C, col = t - e * C;
        float val = 0.f;
        if (prev_c < 0) {
          const int row = p.sel_row[((size_t)b * C + c) * mo + e];  // position in the k-array, read as an anchor id
          if ((int64_t)row < p.N) val = p.mean_logits[((size_t)b * p.N + row) * C + col];
        } else {
          const int rk = p.sel_rank[((size_t)b * C + c) * mo + e];
          if (rk < valid[prev_c]) val = chain[((size_t)prev_c * mo + rk) * C + col];
        }
        cur[t] = val;
      }
      prev_c = c;
      __syncthreads();
    }
  }
  const float scale = p.scales ? p.scales[b] : 1.f;
  for (int i = threadIdx.x; i < mo; i += blockDim.x) {
    const unsigned int pos = ~(unsigned int)(keys[i] & 0xffffffffull);
    const size_t o = (size_t)b * mo + i;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f, cl = 0.f;
    int c = -1, slot = 0;
    if ((int)pos < total_valid) {
      // locate the class of this concat position
      c = 0;
      while (c + 1 < C && cat_off[c + 1] <= (int)pos) ++c;
      slot = (int)pos - cat_off[c];
      const size_t so = ((size_t)b * C + c) * mo + slot;
      const int row = p.sel_row[so];
      bx = reinterpret_cast<const float4*>(p.dec_boxes)[(size_t)b * p.k + row];
      sc = p.sel_scores[so];
      cl = (float)(c + 1);
    }
    if (p.scales) {
      bx.x = __fmul_rn(bx.x, scale);
      bx.y = __fmul_rn(bx.y, scale);
      bx.z = __fmul_rn(bx.z, scale);
      bx.w = __fmul_rn(bx.w, scale);
    }
    reinterpret_cast<float4*>(p.out.boxes)[o] = bx;
    p.out.scores[o] = sc;
    p.out.classes[o] = cl;
    if (p.out.logits) {
      for (int col = 0; col < C; ++col) {
        float val = 0.f;
        if (c >= 0) {
          if (p.strict) {
            val = chain[((size_t)c * mo + slot) * C + col];
          } else {
            const int row = p.sel_row[((size_t)b * C + c) * mo + slot];
            // fused path: the logits of the selected anchor; mid-level API: the row aligned with boxes
            const int64_t lrow = p.topk_idx ? p.topk_idx[(size_t)b * p.k + row] / C : row;
            if (lrow < p.N) val = p.mean_logits[((size_t)b * p.N + lrow) * C + col];
          }
        }
        p.out.logits[o * C + col] = val;
      }
    }
  }
  if (threadIdx.x == 0) p.out.valid[b] = total_valid < mo ? total_valid : mo;
  (void)total;
}

}  // namespace

int udal_run_overlap = 1;  // 0: udal_run keeps its whole tail on the context's stream
int udal_run_reserved_sms = 0;     // SMs the persistent head kernels of a pipelined udal_run leave to the post stream (0: none - the
                                   // kernels claim their work items dynamically, a CTA whose SM is busy with the tail just claims fewer)
int udal_run_prefilter_on_main = 0;
int udal_run_debug_timeline = 0;   // development: print when the tail of every udal_run started / ended relative to its heads

extern "C" {

}  // extern "C"

// scratch of the global variant: per-anchor tensors written by K2 (or the fused predict kernels) and the
// NMS selection; resolved in the context's current scratch bank
struct GlobalScratch {
  udal_prenms_out pre;
  int32_t* sel_idx;
  float* sel_scores;
  int32_t* valid;
};

static int global_scratch(udal_ctx* ctx, int batch, GlobalScratch* g) {
  const udal_config& c = ctx->cfg;
  const int64_t N = ctx->num_anchors;
  const int C = c.num_classes, mo = c.max_output_size;
  const size_t bn = (size_t)batch * N;
  float* logit_buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_A, bn * C * 4 * 2, (void**)&logit_buf));
  char* anc_buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_B, bn * (16 * 3 + 4 + 4), (void**)&anc_buf));
  char* sel_buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_D, (size_t)batch * (mo * 8 + 4), (void**)&sel_buf));
  g->pre.mean_logits = logit_buf;
  g->pre.std_logits = logit_buf + bn * C;
  g->pre.boxes = (float*)anc_buf;
  g->pre.albox = (float*)(anc_buf + bn * 16);
  g->pre.mcbox = (float*)(anc_buf + bn * 32);
  g->pre.scores = (float*)(anc_buf + bn * 48);
  g->pre.classes = (int32_t*)(anc_buf + bn * 52);
  g->sel_idx = (int32_t*)sel_buf;
  g->sel_scores = (float*)(sel_buf + (size_t)batch * mo * 4);
  g->valid = (int32_t*)(sel_buf + (size_t)batch * mo * 8);
  return UDAL_OK;
}

// top-k pre-filter + NMS + assemble from the per-anchor tensors.  Inside a pipelined udal_run this tail (a
// few warps per image) moves to the post stream, where it overlaps the head sampler of the next run.
static int global_tail(udal_ctx* ctx, const GlobalScratch& g, int batch, const float* image_scales,
                       const udal_detections* out, const udal_nms_plan* plan = nullptr) {
  const udal_config& c = ctx->cfg;
  const int64_t N = ctx->num_anchors;
  const int C = c.num_classes, mo = c.max_output_size;
  const udal_prenms_out& pre = g.pre;
  const bool tail_on_post = ctx->in_run && udal_run_overlap && ctx->post_stream != nullptr;
  cudaStream_t main_stream = ctx->stream;
  const int bank = ctx->scratch_bank;
  static cudaEvent_t dbg[2][4];  // [bank]: heads done (main), tail start (post), tail end (post), run start
  static bool dbg_init = false, dbg_valid[2] = {false, false};
  if (udal_run_debug_timeline && tail_on_post) {
    if (!dbg_init) {
      for (auto& b : dbg)
        for (auto& e : b) cudaEventCreate(&e);
      dbg_init = true;
    }
    if (dbg_valid[bank] && cudaEventQuery(dbg[bank][2]) == cudaSuccess) {
      float a = 0, b2 = 0;
      cudaEventElapsedTime(&a, dbg[bank][0], dbg[bank][1]);
      cudaEventElapsedTime(&b2, dbg[bank][1], dbg[bank][2]);
      fprintf(stderr, "tail(bank %d): started %.3f ms after its heads ended, took %.3f ms\n", bank, a, b2);
    }
    cudaEventRecord(dbg[bank][0], main_stream);
  }
  if (tail_on_post) {
    UDAL_CUDA(cudaEventRecord(ctx->ev_pre[bank], main_stream));
    UDAL_CUDA(cudaStreamWaitEvent(ctx->post_stream, ctx->ev_pre[bank], 0));
    ctx->stream = ctx->post_stream;
    if (udal_run_debug_timeline) cudaEventRecord(dbg[bank][1], ctx->post_stream);
  }
  struct Restore {
    udal_ctx* c;
    cudaStream_t s;
    ~Restore() { c->stream = s; }
  } restore{ctx, main_stream};
  if (plan) UDAL_TRY(udal_nms_select(ctx, pre.boxes, pre.scores, batch, (int)N, *plan, g.sel_idx, g.sel_scores, g.valid));
  else UDAL_TRY(udal_launch_nms_v5(ctx, pre.boxes, pre.scores, batch, (int)N, g.sel_idx, g.sel_scores, g.valid));
  AssembleParams a;
  a.batch = batch;
  a.max_out = mo;
  a.C = C;
  a.N = N;
  a.has_al = c.loss_attenuation ? 1 : 0;
  a.has_mc = c.box_mc ? 1 : 0;
  a.has_mcclass = c.cls_mc ? 1 : 0;
  a.img_h = (float)c.image_h;
  a.img_w = (float)c.image_w;
  a.sel_idx = g.sel_idx;
  a.sel_scores = g.sel_scores;
  a.valid = g.valid;
  a.boxes = pre.boxes;
  a.albox = pre.albox;
  a.mcbox = pre.mcbox;
  a.classes = pre.classes;
  a.mean_logits = pre.mean_logits;
  a.std_logits = pre.std_logits;
  a.scales = image_scales;
  a.out = *out;
  assemble_global_kernel<<<batch, 128, 0, ctx->stream>>>(a);
  UDAL_CHECK_LAUNCH(ctx);
  if (udal_host_trace) udal_host_trace_mark("tail enqueued", 0);
  ctx->last_tail_stream = tail_on_post ? ctx->post_stream : nullptr;
  if (tail_on_post) {
    UDAL_CUDA(cudaEventRecord(ctx->ev_post[bank], ctx->post_stream));
    ctx->post_pending[bank] = true;
    if (udal_run_debug_timeline) {
      cudaEventRecord(dbg[bank][2], ctx->post_stream);
      dbg_valid[bank] = true;
    }
  }
  if (udal_host_trace) udal_host_trace_mark("udal_run end", 0);
  return UDAL_OK;
}

int udal_heads_sample_fused(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                            const udal_prenms_out* pre);

// udal_run, serving configuration: BiFPN features -> detections with the predict layers fused with K2
int udal_run_global_fused(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                          const float* image_scales, const udal_detections* out) {
  UDAL_REQUIRE(out->boxes && out->scores && out->classes && out->valid, "NULL output");
  GlobalScratch g;
  UDAL_TRY(global_scratch(ctx, batch, &g));
  // experiment (udal_run_prefilter_on_main, off: measured 3.62 ms vs 3.44 ms per step): the score pre-filter
  // on the main stream between the two heads, only the latency-bound selection on the post stream
  struct Hook {
    const GlobalScratch* g;
    int batch;
    udal_nms_plan plan;
    bool done = false;
  } hook{&g, batch};
  ctx->between_heads_arg = &hook;
  if (udal_run_prefilter_on_main) ctx->between_heads = [](udal_ctx* c, void* arg) -> int {
    Hook* h = static_cast<Hook*>(arg);
    UDAL_TRY(udal_nms_prefilter(c, h->g->pre.scores, h->batch, (int)c->num_anchors, &h->plan));
    h->done = true;
    return UDAL_OK;
  };
  const int st = udal_heads_sample_fused(ctx, feats, batch, keep_masks, seed, &g.pre);
  ctx->between_heads = nullptr;
  ctx->between_heads_arg = nullptr;
  UDAL_TRY(st);
  return global_tail(ctx, g, batch, image_scales, out, hook.done ? &hook.plan : nullptr);
}

extern "C" {

int udal_postprocess_global(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                            const float* image_scales, const udal_detections* out) {
  UDAL_REQUIRE(ctx && cls && box && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(out->boxes && out->scores && out->classes && out->valid, "NULL output");
  const udal_config& c = ctx->cfg;
  UDAL_REQUIRE(c.max_nms_inputs == 0,
               "postprocess_global with max_nms_inputs > 0 is not a functional combination in the reference "
               "(rank mismatch at postprocess.py:615-616); use postprocess_per_class");
  GlobalScratch g;
  UDAL_TRY(global_scratch(ctx, batch, &g));
  UDAL_TRY(udal_launch_decode_moments(ctx, cls, box, batch, &g.pre));
  return global_tail(ctx, g, batch, image_scales, out);
}

int udal_prenms_topk(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                     const udal_prenms_topk_out* out) {
  UDAL_REQUIRE(ctx && cls && box && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  const udal_config& c = ctx->cfg;
  const int k = c.max_nms_inputs;
  UDAL_REQUIRE(k > 0, "udal_prenms_topk needs max_nms_inputs > 0");
  const int64_t N = ctx->num_anchors;
  const int C = c.num_classes;
  UDAL_REQUIRE((int64_t)k <= N * C, "max_nms_inputs %d exceeds N*C", k);
  const size_t bn = (size_t)batch * N;
  float* mean_logits = out->mean_logits;
  float* std_logits = nullptr;
  const bool want_std = out->mcclass && c.cls_mc;
  if (!mean_logits || want_std) {
    float* buf;
    UDAL_TRY(udal_scratch_get(ctx, SCR_POST_A, bn * C * 4 * 2, (void**)&buf));
    if (!mean_logits) mean_logits = buf;
    if (want_std) std_logits = buf + bn * C;
  }
  UDAL_TRY(udal_launch_logit_moments(ctx, cls, batch, mean_logits, std_logits));
  char* tk;
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_C, (size_t)batch * k * 8, (void**)&tk));
  int32_t* idx = out->topk_idx ? out->topk_idx : (int32_t*)tk;
  float* val = (float*)(tk + (size_t)batch * k * 4);
  UDAL_TRY(udal_launch_topk(ctx, mean_logits, batch, N * C, k, idx, val));
  UDAL_TRY(udal_launch_decode_gather(ctx, box, batch, k, idx, val, std_logits, out));
  return UDAL_OK;
}

static int per_class_from_candidates(udal_ctx* ctx, const float* dec_boxes, const float* dec_scores,
                                     const int32_t* dec_classes, const int32_t* tk_idx, int batch, int k,
                                     const float* image_scales, const float* logits, int64_t logit_rows,
                                     int strict_reference, const udal_detections* out) {
  const udal_config& c = ctx->cfg;
  const int C = c.num_classes;
  const int mo = c.max_output_size;
  UDAL_REQUIRE(((uintptr_t)dec_boxes & 15) == 0, "boxes must be 16-byte aligned");
  int p2 = 1;
  while (p2 < C * mo + mo) p2 <<= 1;
  UDAL_REQUIRE(p2 <= 16384 && C <= 1024, "num_classes * max_output_size too large for the merge kernel");
  const size_t bk = (size_t)batch * k;
  const size_t S = (size_t)batch * C;
  char* buf;
  // seg_idx [bk] | seg_start [S] | seg_count [S] | sel_row [S,mo] | sel_rank [S,mo] | sel_scores [S,mo] | valid [S]
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_D, bk * 4 + S * 8 + S * mo * 12 + S * 4, (void**)&buf));
  int32_t* seg_idx = (int32_t*)buf;
  char* q = buf + bk * 4;
  int32_t* seg_start = (int32_t*)q;
  int32_t* seg_count = (int32_t*)(q + S * 4);
  int32_t* sel_row = (int32_t*)(q + S * 8);
  int32_t* sel_rank = (int32_t*)(q + S * 8 + S * mo * 4);
  float* sel_scores = (float*)(q + S * 8 + S * mo * 8);
  int32_t* valid = (int32_t*)(q + S * 8 + S * mo * 12);
  class_partition_kernel<<<batch, 256, 2 * C * sizeof(int), ctx->stream>>>(dec_classes, k, C, seg_idx, seg_start,
                                                                           seg_count);
  UDAL_CHECK_LAUNCH(ctx);
  UDAL_TRY(udal_nms_sorted(ctx, dec_boxes, dec_scores, seg_idx, seg_start, seg_count, nullptr, 0, (int)S, k, C, k,
                           (int64_t)bk, sel_row, sel_rank, sel_scores, valid, nullptr));
  MergeParams m;
  memset(&m, 0, sizeof(m));
  m.batch = batch;
  m.k = k;
  m.C = C;
  m.max_out = mo;
  m.p2 = p2;
  m.strict = strict_reference;
  m.N = logit_rows;
  m.sel_row = sel_row;
  m.sel_rank = sel_rank;
  m.sel_scores = sel_scores;
  m.valid = valid;
  m.seg_count = seg_count;
  m.dec_boxes = dec_boxes;
  m.topk_idx = tk_idx;
  m.mean_logits = logits;
  m.scales = image_scales;
  m.out = *out;
  if (!logits) m.out.logits = nullptr;
  if (m.out.logits && strict_reference) UDAL_TRY(udal_scratch_get(ctx, SCR_POST_B, S * mo * C * 4, (void**)&m.chain));
  const size_t smem = (size_t)p2 * 8;
  if (smem > 40 * 1024)
    UDAL_CUDA(cudaFuncSetAttribute(merge_per_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_per_class_kernel<<<batch, 256, smem, ctx->stream>>>(m);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_per_class_nms(udal_ctx* ctx, const float* boxes, const float* scores, const int32_t* classes, int batch,
                       int k, const float* image_scales, const float* logits, int64_t logit_rows,
                       int strict_reference, const udal_detections* out) {
  UDAL_REQUIRE(ctx && boxes && scores && classes && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(out->boxes && out->scores && out->classes && out->valid, "NULL output");
  UDAL_REQUIRE(batch > 0 && k > 0, "empty input");
  return per_class_from_candidates(ctx, boxes, scores, classes, nullptr, batch, k, image_scales, logits, logit_rows,
                                   strict_reference, out);
}

int udal_postprocess_per_class(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                               const float* image_scales, int strict_reference, const udal_detections* out) {
  UDAL_REQUIRE(ctx && cls && box && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(out->boxes && out->scores && out->classes && out->valid, "NULL output");
  const udal_config& c = ctx->cfg;
  const int k = c.max_nms_inputs;
  UDAL_REQUIRE(k > 0,
               "postprocess_per_class needs max_nms_inputs > 0 (the eval variant, eval.py:75)");
  const int C = c.num_classes;
  const int64_t N = ctx->num_anchors;
  const size_t bk = (size_t)batch * k;
  char* buf;
  // dec boxes [bk,4] | scores [bk] | classes [bk] | topk idx [bk]
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_C2, bk * 28, (void**)&buf));
  float* dec_boxes = (float*)buf;
  float* dec_scores = (float*)(buf + bk * 16);
  int32_t* dec_classes = (int32_t*)(buf + bk * 20);
  int32_t* tk_idx = (int32_t*)(buf + bk * 24);
  float* mean_logits;
  UDAL_TRY(udal_scratch_get(ctx, SCR_POST_A, (size_t)batch * N * C * 4 * 2, (void**)&mean_logits));
  udal_prenms_topk_out pre;
  memset(&pre, 0, sizeof(pre));
  pre.mean_logits = mean_logits;
  pre.topk_idx = tk_idx;
  pre.boxes = dec_boxes;
  pre.scores = dec_scores;
  pre.classes = dec_classes;
  UDAL_TRY(udal_prenms_topk(ctx, cls, box, batch, &pre));
  return per_class_from_candidates(ctx, dec_boxes, dec_scores, dec_classes, tk_idx, batch, k, image_scales,
                                   mean_logits, N, strict_reference, out);
}

// ---- small layout kernels used by the Python mirror ------------------------------------------
__global__ void concat_channels_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb,
                                       int64_t rows, float* __restrict__ out) {
  const int w = ca + cb;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * w) return;
  const int64_t r = i / w;
  const int c = (int)(i - r * w);
  out[i] = c < ca ? a[r * ca + c] : b[r * cb + (c - ca)];
}

int udal_concat_channels(udal_ctx* ctx, const float* a, int ca, const float* b, int cb, int64_t rows, float* out) {
  UDAL_REQUIRE(ctx && a && b && out && ca > 0 && cb > 0, "bad argument");
  UDAL_TRY(udal_join(ctx));
  const int64_t total = rows * (ca + cb);
  if (total == 0) return UDAL_OK;
  concat_channels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(a, ca, b, cb, rows, out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, int64_t n_rows, int width,
                                   const int32_t* __restrict__ idx, int m, int mode, uint32_t* __restrict__ out,
                                   int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % width);
  const int64_t j = (i / width) % m;
  const int64_t b = i / ((int64_t)width * m);
  const int64_t row = idx[b * m + j];
  uint32_t v = 0;
  const bool ok = row >= 0 && row < n_rows;
  if (ok) v = src[(b * n_rows + row) * width + c];
  if (mode == 1) v = __float_as_uint(ok ? (float)((int32_t)v + 1) : 0.f);
  out[i] = v;
}

int udal_gather_rows(udal_ctx* ctx, const void* src, int batch, int64_t n_rows, int width, const int32_t* idx, int m,
                     int mode, void* out) {
  UDAL_REQUIRE(ctx && src && idx && out && width > 0, "bad argument");
  UDAL_TRY(udal_join(ctx));
  const int64_t total = (int64_t)batch * m * width;
  if (total == 0) return UDAL_OK;
  gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)src, n_rows, width, idx,
                                                                                m, mode, (uint32_t*)out, total);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// ---- elementwise helpers of the mirror's small entry points (clip_boxes, topk_class_boxes, pre_nms(topk=False)) ----
__global__ void clip_boxes_kernel(const float* __restrict__ in, int64_t total, float h, float w, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // tf.clip_by_value(boxes, [0], [H, W, H, W]) = minimum(maximum(x, 0), hi)
  out[i] = fminf(fmaxf(in[i], 0.f), (i & 1) ? w : h);
}

int udal_clip_boxes(udal_ctx* ctx, const float* boxes, int64_t rows, float image_h, float image_w, float* out) {
  UDAL_REQUIRE(ctx && boxes && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (rows == 0) return UDAL_OK;
  clip_boxes_kernel<<<(unsigned)((rows * 4 + 255) / 256), 256, 0, ctx->stream>>>(boxes, rows * 4, image_h, image_w, out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void max_reduce_kernel(const float* __restrict__ x, int64_t rows, int c, float* __restrict__ mx,
                                  int32_t* __restrict__ arg) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = x + r * c;
  float best = p[0];
  int a = 0;
  for (int j = 1; j < c; ++j)
    if (p[j] > best) {  // tf.math.argmax: first maximum
      best = p[j];
      a = j;
    }
  mx[r] = best;
  arg[r] = a;
}

int udal_max_reduce(udal_ctx* ctx, const float* x, int64_t rows, int c, float* max_out, int32_t* argmax_out) {
  UDAL_REQUIRE(ctx && x && max_out && argmax_out && c > 0, "bad argument");
  UDAL_TRY(udal_join(ctx));
  if (rows == 0) return UDAL_OK;
  max_reduce_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(x, rows, c, max_out, argmax_out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void divmod_kernel(const int32_t* __restrict__ idx, int64_t total, int d, int32_t* __restrict__ q,
                              int32_t* __restrict__ r) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int32_t v = idx[i];
  q[i] = v / d;
  r[i] = v % d;
}

int udal_divmod_i32(udal_ctx* ctx, const int32_t* idx, int64_t total, int d, int32_t* quot, int32_t* rem) {
  UDAL_REQUIRE(ctx && idx && quot && rem && d > 0, "bad argument");
  UDAL_TRY(udal_join(ctx));
  if (total == 0) return UDAL_OK;
  divmod_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(idx, total, d, quot, rem);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void sigmoid_kernel(const float* __restrict__ x, int64_t total, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  y[i] = (float)(1.0 / (1.0 + exp(-(double)x[i])));  // the oracle's sigmoid: fp32(1 / (1 + exp(-fp64(x))))
}

int udal_sigmoid(udal_ctx* ctx, const float* x, int64_t total, float* y) {
  UDAL_REQUIRE(ctx && x && y, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (total == 0) return UDAL_OK;
  sigmoid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(x, total, y);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void format_detections_kernel(const float* boxes, int box_stride, const float* scores, const float* classes,
                                         int class_stride, const float* ids, const float* widths, int flip,
                                         const float* logits, int nlogits, int64_t rows, int max_out, float* out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t b = r / max_out;
  const float* bx = boxes + r * box_stride;
  float* o = out + r * (7 + nlogits);
  // image_ids * ones_like(scores)
  o[0] = __fmul_rn(ids[b], 1.0f);
  if (flip) {
    o[1] = __fsub_rn(widths[b], bx[3]);
    o[2] = bx[0];
    o[3] = __fsub_rn(widths[b], bx[1]);
    o[4] = bx[2];
  } else {
    o[1] = bx[1];
    o[2] = bx[0];
    o[3] = bx[3];
    o[4] = bx[2];
  }
  o[5] = scores[r];
  o[6] = classes[r * class_stride];
  for (int c = 0; c < nlogits; ++c) o[7 + c] = logits[r * nlogits + c];
}

int udal_format_detections(udal_ctx* ctx, const float* boxes, int box_stride, const float* scores,
                           const float* classes, int class_stride, const float* image_ids, const float* widths,
                           int flip, const float* logits, int nlogits, int batch, int max_out, float* out) {
  UDAL_REQUIRE(ctx && boxes && scores && classes && image_ids && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(!flip || widths, "flip needs the original image widths");
  if (!logits) nlogits = 0;
  const int64_t rows = (int64_t)batch * max_out;
  format_detections_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(
      boxes, box_stride, scores, classes, class_stride, image_ids, widths, flip, logits, nlogits, rows, max_out, out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

__global__ void transform_detections_kernel(const float* in, int64_t rows, int in_cols, float* out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* d = in + r * in_cols;
  float* o = out + r * 7;
  o[0] = d[0];
  o[1] = d[1];
  o[2] = d[2];
  o[3] = __fsub_rn(d[3], d[1]);
  o[4] = __fsub_rn(d[4], d[2]);
  o[5] = d[5];
  o[6] = d[6];
}

int udal_transform_detections(udal_ctx* ctx, const float* in, int64_t rows, int in_cols, float* out) {
  UDAL_REQUIRE(ctx && in && out && in_cols >= 7, "bad argument");
  UDAL_TRY(udal_join(ctx));
  if (rows == 0) return UDAL_OK;
  transform_detections_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(in, rows, in_cols, out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

}  // extern "C"
