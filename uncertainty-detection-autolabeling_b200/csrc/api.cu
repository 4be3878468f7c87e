// Context, memory and thin public wrappers of libudal (see include/udal.h).
#include <stdarg.h>

#include <algorithm>

#include <chrono>

#include "udal_common.cuh"

static thread_local char g_err[1024] = "";

void udal_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int udal_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  udal_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? UDAL_ERR_NOMEM : UDAL_ERR_CUDA;
}

static bool scratch_banked(int slot) {
  // the head sampler's slots are only ever touched on the context's own stream
  return !(slot == SCR_HEADS_A || slot == SCR_HEADS_B || slot == SCR_HEADS_C || slot == SCR_PRE_A || slot == SCR_LEVEL_PTRS);
}

int udal_host_trace = 0;
void udal_host_trace_mark(const char* file, int line) {
  static thread_local std::chrono::steady_clock::time_point last;
  static thread_local const char* last_file = nullptr;
  static thread_local int last_line = 0;
  const auto now = std::chrono::steady_clock::now();
  if (last_file) {
    const double ms = std::chrono::duration<double, std::milli>(now - last).count();
    if (ms > 0.3) fprintf(stderr, "host gap %.3f ms between %s:%d and %s:%d\n", ms, last_file, last_line, file, line);
  }
  last = now;
  last_file = file;
  last_line = line;
}

int udal_join(udal_ctx* ctx) {
  if (ctx->in_run) return UDAL_OK;
  // the current device is per host thread: a caller that drives several contexts from a thread pool
  // (scheduler.py, PipelinedSampler) must not have to remember cudaSetDevice
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  for (int b = 0; b < 2; ++b)
    if (ctx->post_pending[b]) {
      UDAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_post[b], 0));
      ctx->post_pending[b] = false;
    }
  ctx->scratch_bank = 0;
  ctx->run_bank = 0;
  ctx->last_tail_stream = nullptr;   // everything is ordered on the context's stream again
  return UDAL_OK;
}

int udal_scratch_get(udal_ctx* ctx, int slot, size_t bytes, void** out) {
  udal_scratch& s = ctx->scratch[slot + ((ctx->scratch_bank && scratch_banked(slot)) ? 20 : 0)];
  if (bytes == 0) bytes = 16;
  if (s.bytes < bytes) {
    if (s.ptr) {
      // the old block may still be in use by enqueued work
      UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
      UDAL_CUDA(cudaFree(s.ptr));
      s.ptr = nullptr;
      s.bytes = 0;
    }
    size_t want = bytes + bytes / 4;
    want = (want + 255) & ~(size_t)255;
    UDAL_CUDA(cudaMalloc(&s.ptr, want));
    s.bytes = want;
  }
  *out = s.ptr;
  return UDAL_OK;
}

int udal_work_counters_reset(udal_ctx* ctx) {
  if (!ctx->work_counters) UDAL_CUDA(cudaMalloc(&ctx->work_counters, UDAL_WORK_COUNTERS * sizeof(int)));
  UDAL_CUDA(cudaMemsetAsync(ctx->work_counters, 0, UDAL_WORK_COUNTERS * sizeof(int), ctx->stream));
  ctx->work_counter_next = 0;
  return UDAL_OK;
}

int udal_work_counter(udal_ctx* ctx, int** out) {
  UDAL_REQUIRE(ctx->work_counters && ctx->work_counter_next < UDAL_WORK_COUNTERS, "persistent-kernel work counters exhausted");
  *out = ctx->work_counters + ctx->work_counter_next++;
  return UDAL_OK;
}

extern "C" {

const char* udal_last_error(void) { return g_err; }
int udal_abi_version(void) { return UDAL_ABI_VERSION; }

int udal_device_count(int* count) {
  UDAL_REQUIRE(count, "NULL count");
  UDAL_CUDA(cudaGetDeviceCount(count));
  return UDAL_OK;
}

int udal_create(const udal_config* cfg, udal_ctx** out) {
  UDAL_REQUIRE(cfg && out, "udal_create: NULL argument");
  UDAL_REQUIRE(cfg->abi_version == UDAL_ABI_VERSION, "udal_create: ABI version %d, library is %d",
               cfg->abi_version, UDAL_ABI_VERSION);
  UDAL_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= UDAL_MAX_LEVELS, "num_levels %d outside [1,%d]",
               cfg->num_levels, UDAL_MAX_LEVELS);
  UDAL_REQUIRE(cfg->anchors_per_loc >= 1 && cfg->anchors_per_loc <= 64, "anchors_per_loc %d unsupported",
               cfg->anchors_per_loc);
  UDAL_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 1024, "num_classes %d unsupported", cfg->num_classes);
  UDAL_REQUIRE(cfg->mc_samples >= 1 && cfg->mc_samples <= 4096, "mc_samples %d unsupported", cfg->mc_samples);
  UDAL_REQUIRE(cfg->max_output_size >= 1 && cfg->max_output_size <= 2048, "max_output_size %d unsupported",
               cfg->max_output_size);
  UDAL_REQUIRE(cfg->decode_method >= UDAL_DECODE_LNORM && cfg->decode_method <= UDAL_DECODE_FALSEDEC,
               "decode method %d unsupported (the 'sample' method draws from tfp and is not offered)",
               cfg->decode_method);
  UDAL_REQUIRE(cfg->decode_precision == UDAL_DECODE_FP64 || cfg->decode_precision == UDAL_DECODE_FP32,
               "unknown decode_precision %d", cfg->decode_precision);
  UDAL_REQUIRE(cfg->nms_method == UDAL_NMS_HARD || cfg->nms_method == UDAL_NMS_GAUSSIAN,
               "Inference has invalid nms method %d", cfg->nms_method);
  UDAL_REQUIRE(cfg->max_nms_inputs >= 0 && cfg->max_nms_inputs <= 8192, "max_nms_inputs %d outside [0,8192]",
               cfg->max_nms_inputs);
  for (int l = 0; l < cfg->num_levels; ++l)
    UDAL_REQUIRE(cfg->level_h[l] > 0 && cfg->level_w[l] > 0, "level %d has an empty feature map", l);
  int ndev = 0;
  UDAL_CUDA(cudaGetDeviceCount(&ndev));
  UDAL_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "device %d not present (%d visible)", cfg->device, ndev);
  UDAL_CUDA(cudaSetDevice(cfg->device));
  udal_ctx* ctx = new udal_ctx();
  ctx->cfg = *cfg;
  int64_t off = 0;
  for (int l = 0; l <= UDAL_MAX_LEVELS; ++l) {
    ctx->level_pix_off[l] = off;
    if (l < cfg->num_levels) off += (int64_t)cfg->level_h[l] * cfg->level_w[l];
  }
  ctx->num_pixels = off;
  ctx->num_anchors = off * cfg->anchors_per_loc;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->post_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < UDAL_STAGE_SLOTS && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&ctx->ev_staged[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fetched[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_start);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_stop);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = cudaEventCreateWithFlags(&ctx->ev_pre[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_post[b], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    delete ctx;
    return udal_cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return UDAL_OK;
}

static void free_head(udal_head_weights_dev& h) {
  cudaFree(h.dw);
  cudaFree(h.pw);
  cudaFree(h.bias);
  cudaFree(h.bn_scale);
  cudaFree(h.bn_shift);
  cudaFree(h.dwp);
  cudaFree(h.pwp);
  cudaFree(h.bp);
  cudaFree(h.pw_bf16);
  cudaFree(h.pwp_bf16);
  cudaFree(h.fold_bias);
  cudaFree(h.ig_w);
  cudaFree(h.fused_w);
  cudaFree(h.wide_w);
  cudaFree(h.wide_f);
  cudaFree(h.l0_w);
  cudaFree(h.x3_w);
  cudaFree(h.x3_f);
  cudaFree(h.l0_ep);
  h = udal_head_weights_dev();
}

int udal_destroy(udal_ctx* ctx) {
  if (!ctx) return UDAL_OK;
  cudaSetDevice(ctx->cfg.device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->post_stream) cudaStreamSynchronize(ctx->post_stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  for (int i = 0; i < UDAL_STAGE_SLOTS; ++i) {
    if (ctx->ev_staged[i]) cudaEventDestroy(ctx->ev_staged[i]);
    if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
    if (ctx->ev_fetched[i]) cudaEventDestroy(ctx->ev_fetched[i]);
  }
  for (auto& s : ctx->scratch) cudaFree(s.ptr);
  for (void* p : ctx->user_allocs) cudaFree(p);
  cudaFree(ctx->anchors);
  cudaFree(ctx->work_counters);
  free_head(ctx->heads[0]);
  free_head(ctx->heads[1]);
  cudaEventDestroy(ctx->ev_start);
  cudaEventDestroy(ctx->ev_stop);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_pre[b]) cudaEventDestroy(ctx->ev_pre[b]);
    if (ctx->ev_post[b]) cudaEventDestroy(ctx->ev_post[b]);
  }
  if (ctx->post_stream) cudaStreamDestroy(ctx->post_stream);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return UDAL_OK;
}

int udal_set_stream(udal_ctx* ctx, void* cuda_stream) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return UDAL_OK;
}

#define UDAL_SLOT_OK(slot) UDAL_REQUIRE(ctx && (slot) >= 0 && (slot) < UDAL_STAGE_SLOTS, "bad staging slot %d", (slot))

int udal_stage_begin(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  if (ctx->consumed_pending[slot]) {
    UDAL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[slot], 0));
    ctx->consumed_pending[slot] = false;
  }
  return UDAL_OK;
}

int udal_stage_h2d(udal_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "NULL argument");
  if (bytes) UDAL_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  return UDAL_OK;
}

int udal_stage_end(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaEventRecord(ctx->ev_staged[slot], ctx->copy_stream));
  return UDAL_OK;
}

int udal_stage_acquire(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_staged[slot], 0));
  return UDAL_OK;
}

int udal_stage_release(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaEventRecord(ctx->ev_consumed[slot], ctx->stream));
  ctx->consumed_pending[slot] = true;
  return UDAL_OK;
}

int udal_fetch_d2h(udal_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "NULL argument");
  cudaStream_t st = ctx->last_tail_stream ? ctx->last_tail_stream : ctx->stream;
  if (bytes) UDAL_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
  return UDAL_OK;
}

int udal_fetch_mark(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaEventRecord(ctx->ev_fetched[slot], ctx->last_tail_stream ? ctx->last_tail_stream : ctx->stream));
  return UDAL_OK;
}

int udal_fetch_wait(udal_ctx* ctx, int slot) {
  UDAL_SLOT_OK(slot);
  UDAL_CUDA(cudaEventSynchronize(ctx->ev_fetched[slot]));
  return UDAL_OK;
}

int udal_set_feature_format(udal_ctx* ctx, int format) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_REQUIRE(format == UDAL_FEAT_F32 || format == UDAL_FEAT_F16, "unknown feature format %d", format);
  if (format == UDAL_FEAT_F16)
    UDAL_REQUIRE(ctx->cfg.heads_mode == UDAL_HEADS_FP16_TC && ctx->cfg.num_filters == 64,
                 "fp16 feature maps: heads_mode fp16 and fpn_num_filters 64 (the tensor-core layer-0 kernel reads them directly)");
  ctx->feat_f16 = format == UDAL_FEAT_F16;
  return UDAL_OK;
}

int udal_get_stream(udal_ctx* ctx, void** cuda_stream) {
  UDAL_REQUIRE(ctx && cuda_stream, "NULL argument");
  *cuda_stream = (void*)ctx->stream;
  return UDAL_OK;
}

int udal_wait_stream(udal_ctx* ctx, void* producer_stream) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  if ((cudaStream_t)producer_stream == ctx->stream) return UDAL_OK;
  // a fresh event per call: the wait is consumed when it is enqueued, the event can be destroyed right away
  cudaEvent_t ev;
  UDAL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, (cudaStream_t)producer_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);
  UDAL_CUDA(e);
  return UDAL_OK;
}

int udal_wait_context(udal_ctx* ctx, udal_ctx* producer) {
  UDAL_REQUIRE(ctx && producer, "NULL ctx");
  if (ctx == producer) return UDAL_OK;
  // everything the producer has enqueued - its udal_run tails included - before anything this context enqueues from now on
  UDAL_TRY(udal_join(producer));
  cudaEvent_t ev;
  UDAL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, producer->stream);
  if (e == cudaSuccess) e = cudaSetDevice(ctx->cfg.device);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);
  UDAL_CUDA(e);
  return UDAL_OK;
}

int udal_sync(udal_ctx* ctx) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  return UDAL_OK;
}

int udal_malloc(udal_ctx* ctx, size_t bytes, void** dev_ptr) {
  UDAL_REQUIRE(ctx && dev_ptr, "NULL argument");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  void* p = nullptr;
  UDAL_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
  ctx->user_allocs.push_back(p);
  *dev_ptr = p;
  return UDAL_OK;
}

int udal_free(udal_ctx* ctx, void* dev_ptr) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  if (!dev_ptr) return UDAL_OK;
  auto it = std::find(ctx->user_allocs.begin(), ctx->user_allocs.end(), dev_ptr);
  UDAL_REQUIRE(it != ctx->user_allocs.end(), "udal_free: pointer was not allocated by this context");
  ctx->user_allocs.erase(it);
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  UDAL_CUDA(cudaFree(dev_ptr));
  return UDAL_OK;
}

int udal_host_alloc(size_t bytes, void** pinned_ptr) {
  UDAL_REQUIRE(pinned_ptr, "NULL argument");
  UDAL_CUDA(cudaHostAlloc(pinned_ptr, bytes ? bytes : 16, cudaHostAllocDefault));
  return UDAL_OK;
}

int udal_host_free(void* pinned_ptr) {
  if (pinned_ptr) UDAL_CUDA(cudaFreeHost(pinned_ptr));
  return UDAL_OK;
}

int udal_memcpy_h2d(udal_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (bytes) UDAL_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return UDAL_OK;
}

int udal_memcpy_d2h(udal_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (bytes) UDAL_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return UDAL_OK;
}

int udal_memcpy_d2d(udal_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || (dst_dev && src_dev)), "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (bytes) UDAL_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return UDAL_OK;
}

int udal_memset(udal_ctx* ctx, void* dst_dev, int value, size_t bytes) {
  UDAL_REQUIRE(ctx && (bytes == 0 || dst_dev), "NULL argument");
  UDAL_TRY(udal_join(ctx));
  if (bytes) UDAL_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return UDAL_OK;
}

int udal_timer_start(udal_ctx* ctx) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  UDAL_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
  return UDAL_OK;
}

int udal_timer_stop(udal_ctx* ctx, float* elapsed_ms) {
  UDAL_REQUIRE(ctx && elapsed_ms, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
  UDAL_CUDA(cudaEventSynchronize(ctx->ev_stop));
  UDAL_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev_start, ctx->ev_stop));
  return UDAL_OK;
}

int udal_num_anchors(const udal_ctx* ctx, int64_t* n) {
  UDAL_REQUIRE(ctx && n, "NULL argument");
  *n = ctx->num_anchors;
  return UDAL_OK;
}

int udal_launch_count(const udal_ctx* ctx, int64_t* n) {
  UDAL_REQUIRE(ctx && n, "NULL argument");
  *n = ctx->launches;
  return UDAL_OK;
}

int udal_profile_layers(udal_ctx* ctx, int enable) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  ctx->profile_layers = enable != 0;
  for (cudaEvent_t e : ctx->layer_events) cudaEventDestroy(e);
  ctx->layer_events.clear();
  return UDAL_OK;
}

int udal_get_layer_times(udal_ctx* ctx, float* ms, int cap, int* n) {
  UDAL_REQUIRE(ctx && ms && n, "NULL argument");
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  const int pairs = (int)ctx->layer_events.size() / 2;
  *n = pairs < cap ? pairs : cap;
  for (int i = 0; i < *n; ++i) UDAL_CUDA(cudaEventElapsedTime(&ms[i], ctx->layer_events[2 * i], ctx->layer_events[2 * i + 1]));
  for (cudaEvent_t e : ctx->layer_events) cudaEventDestroy(e);
  ctx->layer_events.clear();
  return UDAL_OK;
}

int udal_scratch_bytes(const udal_ctx* ctx, size_t* bytes) {
  UDAL_REQUIRE(ctx && bytes, "NULL argument");
  size_t t = 0;
  for (const auto& s : ctx->scratch) t += s.bytes;
  *bytes = t;
  return UDAL_OK;
}

int udal_set_anchors(udal_ctx* ctx, const float* anchors_host, int64_t num_anchors) {
  UDAL_REQUIRE(ctx && anchors_host, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(num_anchors == ctx->num_anchors, "anchor table has %lld rows, the level geometry gives %lld",
               (long long)num_anchors, (long long)ctx->num_anchors);
  if (!ctx->anchors) UDAL_CUDA(cudaMalloc(&ctx->anchors, (size_t)num_anchors * 16));
  // pageable source: the copy is staged before the call returns
  UDAL_CUDA(cudaMemcpyAsync(ctx->anchors, anchors_host, (size_t)num_anchors * 16, cudaMemcpyHostToDevice, ctx->stream));
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->anchors_set = true;
  return UDAL_OK;
}

int udal_decode_moments(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                        const udal_prenms_out* out) {
  UDAL_REQUIRE(ctx && cls && box && out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  return udal_launch_decode_moments(ctx, cls, box, batch, out);
}

int udal_topk(udal_ctx* ctx, const float* values, int batch, int64_t m, int k, int32_t* idx_out, float* val_out) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  return udal_launch_topk(ctx, values, batch, m, k, idx_out, val_out);
}

int udal_nms_v5(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n, int32_t* sel_idx,
                float* sel_scores, int32_t* valid) {
  UDAL_REQUIRE(ctx && boxes && scores && sel_idx && sel_scores && valid, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  return udal_launch_nms_v5(ctx, boxes, scores, segments, n, sel_idx, sel_scores, valid);
}

}  // extern "C"
