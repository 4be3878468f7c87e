// K1 (fp32 reference-precision mode): the EfficientDet class / box(+sigma) head towers with
// MC-dropout, starting at the BiFPN outputs.
//
// Replaces (reference src/): efficientdet_keras.py:353-513 ClassNet, 516-692 BoxNet, the MC loop
// of EfficientDetNet.call 979-1050 and utils_extra.py:201-217 stack_mcpred.
//
// One launch per tower layer covers all pyramid levels, all images and all MC samples:
//   depthwise 3x3 (SAME, zero pad) -> pointwise 1x1 + bias -> per-level BN -> swish
// fused in one kernel; SpatialDropout2D is a per-(sample, image, channel) scale that the NEXT
// layer applies while it loads its input tile, so layer 0 (whose input, the BiFPN features, does
// not depend on the sample) is computed once per image instead of T times.
// This is the CUDA-core fp32 path (parity mode); heads_tc.cu holds the tcgen05 bf16 path.
#include "udal_common.cuh"

namespace {

constexpr int TH = 8, TW = 16, TPX = TH * TW;     // output tile: 128 pixels
constexpr int HALO_W = TW + 2, HALO_PX = (TH + 2) * (TW + 2);  // 10 x 18 = 180
constexpr int IN_STRIDE = HALO_PX + 1;             // odd stride: conflict-free channel-major tile
constexpr int NCHUNK = 64;                         // output channels per GEMM pass
constexpr int kThreads = 256;

struct LayerParams {
  int num_levels;
  int h[UDAL_MAX_LEVELS], w[UDAL_MAX_LEVELS];
  int tiles_x[UDAL_MAX_LEVELS];
  int tile_off[UDAL_MAX_LEVELS + 1];
  const float* in[UDAL_MAX_LEVELS];    // [NB_in, H, W, F]
  float* out[UDAL_MAX_LEVELS];         // [NB_out, H, W, Cout]
  const float* in_scale[UDAL_MAX_LEVELS];  // [NB_out, F] dropout scale of the producer layer, or null
  const float* bn_scale[UDAL_MAX_LEVELS];  // [Cout] or null (predict layer)
  const float* bn_shift[UDAL_MAX_LEVELS];
  const float* dw;    // [9][F]
  const float* pw;    // [F][Cout]
  const float* bias;  // [Cout]
  int F, Cout, batch;
  int in_per_sample;  // 1: input indexed by nb (= t*B+b); 0: by b = nb % batch
  int act;            // 1: BN + swish, 0: linear (predict), 2: BN only (BiFPN op_after_combine: the activation precedes the conv)
};

__device__ __forceinline__ float swish(float x) { return __fdiv_rn(x, 1.f + expf(-x)); }

__global__ void __launch_bounds__(kThreads) sepconv_layer_kernel(const LayerParams p) {
  extern __shared__ float smem[];
  const int F = p.F;
  float* in_tile = smem;                    // [F][IN_STRIDE]
  float* As = in_tile + (size_t)F * IN_STRIDE;  // [F][TPX]
  float* Bs = As + (size_t)F * TPX;          // [F][NCHUNK]
  const int tid = threadIdx.x;
  int l = 0;
#pragma unroll
  for (int i = 1; i < UDAL_MAX_LEVELS; ++i)
    if (i < p.num_levels && (int)blockIdx.x >= p.tile_off[i]) l = i;
  const int H = p.h[l], W = p.w[l];
  const int tile = blockIdx.x - p.tile_off[l];
  const int ty0 = (tile / p.tiles_x[l]) * TH, tx0 = (tile % p.tiles_x[l]) * TW;
  const int nb = blockIdx.y;
  const int in_img = p.in_per_sample ? nb : nb % p.batch;
  const float* in = p.in[l] + (size_t)in_img * H * W * F;
  const float* scale = p.in_scale[l] ? p.in_scale[l] + (size_t)nb * F : nullptr;

  // ---- load the input tile with halo, channel-major, zero padded, dropout scale applied ----
  for (int e = tid; e < HALO_PX * F; e += kThreads) {
    const int c = e % F, px = e / F;
    const int y = ty0 + px / HALO_W - 1, x = tx0 + px % HALO_W - 1;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      v = __ldg(in + ((size_t)y * W + x) * F + c);
      if (scale) v = __fmul_rn(v, __ldg(scale + c));
    }
    in_tile[c * IN_STRIDE + px] = v;
  }
  __syncthreads();

  // ---- depthwise 3x3 -> As[c][px] ----
  {
    const int px = tid % TPX;
    const int py = px / TW, pxx = px % TW;
    for (int c = tid / TPX; c < F; c += kThreads / TPX) {
      const float* t = in_tile + c * IN_STRIDE + py * HALO_W + pxx;
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc = fmaf(t[dy * HALO_W + dx], __ldg(p.dw + (dy * 3 + dx) * F + c), acc);
      As[c * TPX + px] = acc;
    }
  }

  // ---- pointwise GEMM [128 x F] x [F x Cout], NCHUNK output channels per pass ----
  const int tx = tid % 16, ty = tid / 16;  // tx: 4 output channels, ty: 8 pixels
  for (int n0 = 0; n0 < p.Cout; n0 += NCHUNK) {
    __syncthreads();
    for (int e = tid; e < F * NCHUNK; e += kThreads) {
      const int k = e / NCHUNK, n = e % NCHUNK;
      Bs[e] = (n0 + n < p.Cout) ? __ldg(p.pw + (size_t)k * p.Cout + n0 + n) : 0.f;
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < F; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + k * TPX + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + k * TPX + ty * 8 + 4);
      const float4 b = *reinterpret_cast<const float4*>(Bs + k * NCHUNK + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    // ---- epilogue: bias, BN, swish, store ----
    float* out = p.out[l] + (size_t)nb * H * W * p.Cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int px = ty * 8 + i;
      const int y = ty0 + px / TW, x = tx0 + px % TW;
      if (y >= H || x >= W) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= p.Cout) continue;
        float v = acc[i][j] + __ldg(p.bias + n);
        if (p.act) {
          v = fmaf(v, __ldg(p.bn_scale[l] + n), __ldg(p.bn_shift[l] + n));
          if (p.act == 1) v = swish(v);
        }
        out[((size_t)y * W + x) * p.Cout + n] = v;
      }
    }
  }
}

}  // namespace

// one separable conv layer on a single feature map (BiFPN, bifpn.cu): depthwise 3x3 SAME -> pointwise + bias -> act
int udal_sepconv_fp32_launch(udal_ctx* ctx, const float* in, int NB, int H, int W, int F, int Cout, const float* dw,
                             const float* pw, const float* bias, const float* bn_scale, const float* bn_shift, int act,
                             float* out) {
  LayerParams p;
  memset(&p, 0, sizeof(p));
  p.num_levels = 1;
  p.h[0] = H;
  p.w[0] = W;
  p.tiles_x[0] = (W + TW - 1) / TW;
  const int tiles = p.tiles_x[0] * ((H + TH - 1) / TH);
  for (int l = 1; l <= UDAL_MAX_LEVELS; ++l) p.tile_off[l] = tiles;
  p.in[0] = in;
  p.out[0] = out;
  p.bn_scale[0] = bn_scale;
  p.bn_shift[0] = bn_shift;
  p.dw = dw;
  p.pw = pw;
  p.bias = bias;
  p.F = F;
  p.Cout = Cout;
  p.batch = NB;
  p.in_per_sample = 1;
  p.act = act;
  const size_t smem = ((size_t)F * IN_STRIDE + (size_t)F * TPX + (size_t)F * NCHUNK) * sizeof(float);
  UDAL_REQUIRE(smem <= 227 * 1024, "separable conv with %d channels needs %zu bytes of shared memory", F, smem);
  UDAL_REQUIRE(NB <= 65535, "separable conv: %d images per launch", NB);
  UDAL_CUDA(cudaFuncSetAttribute(sepconv_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sepconv_layer_kernel<<<dim3(tiles, NB), kThreads, smem, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

namespace {

// ---- dropout scales ---------------------------------------------------------------------------
// scale[t][head][l][r][b][f] = keep ? 1/(1-rate) : 0   (SpatialDropout2D, noise shape [B,1,1,F])
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__global__ void dropout_scale_kernel(float* __restrict__ scale, const uint8_t* __restrict__ keep, int64_t total,
                                     int64_t per_head_block, float rate_class, float rate_box,
                                     float inv_class, float inv_box, uint64_t seed) {
  // layout [T][2][L*R*B*F]: per_head_block = L*R*B*F
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int head = (int)((i / per_head_block) % 2);
  const float rate = head == 0 ? rate_class : rate_box;
  bool k;
  if (keep) {
    k = keep[i] != 0;
  } else {
    uint32_t r[4];
    const uint64_t g = (uint64_t)i >> 2;
    philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const float u = (float)(r[i & 3] >> 8) * (1.0f / 16777216.0f);
    k = u >= rate;
  }
  scale[i] = k ? (head == 0 ? inv_class : inv_box) : 0.f;
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, int n,
                               float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float inv = gamma[i] / sqrtf(var[i] + 1e-3f);  // utils_keras.py:78 epsilon
  scale[i] = inv;
  shift[i] = beta[i] - mean[i] * inv;
}

int upload(udal_ctx* ctx, float** dst, const float* src, size_t n) {
  if (*dst) UDAL_CUDA(cudaFree(*dst));
  *dst = nullptr;
  UDAL_CUDA(cudaMalloc(dst, n * sizeof(float)));
  UDAL_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return UDAL_OK;
}

}  // namespace

int udal_heads_tc_prepare(udal_ctx* ctx, int head);  // heads_tc.cu
int udal_heads_x3_prepare(udal_ctx* ctx, int head);  // heads_wide.cu
int udal_heads_x3_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                         float* const* box_out);
int udal_heads_tc_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                         float* const* box_out, const udal_prenms_out* fused_pre);

extern "C" int udal_set_head_weights(udal_ctx* ctx, int head, const float* dw, const float* pw, const float* bias,
                                     const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                                     const float* bn_var, const float* dwp, const float* pwp, const float* bp) {
  UDAL_REQUIRE(ctx, "NULL ctx");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(head == UDAL_HEAD_CLASS || head == UDAL_HEAD_BOX, "head must be 0 (class) or 1 (box)");
  UDAL_REQUIRE(dw && pw && bias && bn_gamma && bn_beta && bn_mean && bn_var && dwp && pwp && bp, "NULL weight pointer");
  const udal_config& c = ctx->cfg;
  UDAL_REQUIRE(c.num_filters > 0 && c.num_filters % 4 == 0 && c.num_filters <= 256, "num_filters %d unsupported",
               c.num_filters);
  UDAL_REQUIRE(c.repeats >= 1 && c.repeats <= 8, "repeats %d unsupported", c.repeats);
  UDAL_CUDA(cudaSetDevice(c.device));
  udal_head_weights_dev& h = ctx->heads[head];
  const int F = c.num_filters, R = c.repeats, L = c.num_levels;
  h.cout = head == UDAL_HEAD_CLASS ? c.anchors_per_loc * c.num_classes : udal_box_channels(ctx);
  UDAL_TRY(upload(ctx, &h.dw, dw, (size_t)R * 9 * F));
  UDAL_TRY(upload(ctx, &h.pw, pw, (size_t)R * F * F));
  UDAL_TRY(upload(ctx, &h.bias, bias, (size_t)R * F));
  UDAL_TRY(upload(ctx, &h.dwp, dwp, (size_t)9 * F));
  UDAL_TRY(upload(ctx, &h.pwp, pwp, (size_t)F * h.cout));
  UDAL_TRY(upload(ctx, &h.bp, bp, (size_t)h.cout));
  const size_t nbn = (size_t)R * L * F;
  float* tmp;
  UDAL_TRY(udal_scratch_get(ctx, SCR_MISC, nbn * 4 * sizeof(float), (void**)&tmp));
  UDAL_CUDA(cudaMemcpyAsync(tmp, bn_gamma, nbn * 4, cudaMemcpyHostToDevice, ctx->stream));
  UDAL_CUDA(cudaMemcpyAsync(tmp + nbn, bn_beta, nbn * 4, cudaMemcpyHostToDevice, ctx->stream));
  UDAL_CUDA(cudaMemcpyAsync(tmp + 2 * nbn, bn_mean, nbn * 4, cudaMemcpyHostToDevice, ctx->stream));
  UDAL_CUDA(cudaMemcpyAsync(tmp + 3 * nbn, bn_var, nbn * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (h.bn_scale) UDAL_CUDA(cudaFree(h.bn_scale));
  if (h.bn_shift) UDAL_CUDA(cudaFree(h.bn_shift));
  h.bn_scale = h.bn_shift = nullptr;
  UDAL_CUDA(cudaMalloc(&h.bn_scale, nbn * 4));
  UDAL_CUDA(cudaMalloc(&h.bn_shift, nbn * 4));
  bn_fold_kernel<<<(int)((nbn + 255) / 256), 256, 0, ctx->stream>>>(tmp, tmp + nbn, tmp + 2 * nbn, tmp + 3 * nbn,
                                                                    (int)nbn, h.bn_scale, h.bn_shift);
  UDAL_CHECK_LAUNCH(ctx);
  // host sources are pageable: make sure the copies are done before the caller reuses them
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  h.set = true;
  if (c.heads_mode == UDAL_HEADS_FP32X3_TC) UDAL_TRY(udal_heads_x3_prepare(ctx, head));
  else if (c.heads_mode != UDAL_HEADS_FP32) UDAL_TRY(udal_heads_tc_prepare(ctx, head));
  return UDAL_OK;
}

static int run_tower_fp32(udal_ctx* ctx, int head, const float* const* feats, int batch, const float* scale_all,
                          float* const* outs) {
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  const int F = c.num_filters, R = c.repeats, L = c.num_levels, T = c.mc_samples, B = batch;
  const bool mc = head == UDAL_HEAD_CLASS ? c.cls_mc != 0 : c.box_mc != 0;
  const int NBt = mc ? T * B : B;  // (sample, image) pairs from layer 1 on
  // activation buffers: a0 [B,P,F] (sample-invariant layer 0) and two ping-pong [NBt,P,F]
  const size_t P = (size_t)ctx->num_pixels;
  float *a0, *pp;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_A, (size_t)B * P * F * 4, (void**)&a0));
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_B, 2 * (size_t)NBt * P * F * 4, (void**)&pp));
  LayerParams p;
  memset(&p, 0, sizeof(p));
  p.num_levels = L;
  int off = 0;
  for (int l = 0; l <= UDAL_MAX_LEVELS; ++l) {
    p.tile_off[l] = off;
    if (l < L) {
      p.h[l] = c.level_h[l];
      p.w[l] = c.level_w[l];
      p.tiles_x[l] = (p.w[l] + TW - 1) / TW;
      off += p.tiles_x[l] * ((p.h[l] + TH - 1) / TH);
    }
  }
  const int total_tiles = off;
  p.F = F;
  p.batch = B;
  const size_t smem = ((size_t)F * IN_STRIDE + (size_t)F * TPX + (size_t)F * NCHUNK) * sizeof(float);
  UDAL_REQUIRE(smem <= 227 * 1024, "num_filters %d needs %zu bytes of shared memory", F, smem);
  UDAL_CUDA(cudaFuncSetAttribute(sepconv_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int layer = 0; layer <= R; ++layer) {
    const bool predict = layer == R;
    p.dw = predict ? h.dwp : h.dw + (size_t)layer * 9 * F;
    p.pw = predict ? h.pwp : h.pw + (size_t)layer * F * F;
    p.bias = predict ? h.bp : h.bias + (size_t)layer * F;
    p.Cout = predict ? h.cout : F;
    p.act = predict ? 0 : 1;
    p.in_per_sample = layer >= 2 ? 1 : 0;
    const int nb_out = layer == 0 ? B : NBt;
    for (int l = 0; l < L; ++l) {
      const size_t lvl = (size_t)ctx->level_pix_off[l] * F;
      if (layer == 0) p.in[l] = feats[l];
      else if (layer == 1) p.in[l] = a0 + (size_t)B * lvl;
      else p.in[l] = pp + (size_t)((layer - 1) & 1) * NBt * P * F + (size_t)NBt * lvl;
      if (predict) p.out[l] = outs[l];
      else if (layer == 0) p.out[l] = a0 + (size_t)B * lvl;
      else p.out[l] = pp + (size_t)(layer & 1) * NBt * P * F + (size_t)NBt * lvl;
      // SpatialDropout2D of the producer layer (layer-1), applied while loading
      p.in_scale[l] = (mc && layer >= 1) ? scale_all + (((size_t)head * L + l) * R + (layer - 1)) * (size_t)NBt * F
                                         : nullptr;
      p.bn_scale[l] = predict ? nullptr : h.bn_scale + ((size_t)layer * L + l) * F;
      p.bn_shift[l] = predict ? nullptr : h.bn_shift + ((size_t)layer * L + l) * F;
    }
    dim3 grid(total_tiles, nb_out);
    sepconv_layer_kernel<<<grid, kThreads, smem, ctx->stream>>>(p);
    UDAL_CHECK_LAUNCH(ctx);
  }
  return UDAL_OK;
}

// scale buffer layout used by the towers: [2][L][R][T*B][F]  (head-major so a (head,l,r) slab is a
// contiguous [T*B, F] matrix); keep masks arrive as [T][2][L][R][B][F].
__global__ void scale_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int L, int R,
                                       int B, int F) {
  const int64_t total = (int64_t)T * 2 * L * R * B * F;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t r = i;
  const int f = (int)(r % F);
  r /= F;
  const int b = (int)(r % B);
  r /= B;
  const int rr = (int)(r % R);
  r /= R;
  const int l = (int)(r % L);
  r /= L;
  const int head = (int)(r % 2);
  const int t = (int)(r / 2);
  dst[((((size_t)head * L + l) * R + rr) * (size_t)(T * B) + (size_t)t * B + b) * F + f] = src[i];
}

// fused_pre != null: the predict layers are fused with K2 and write the per-anchor tensors of *fused_pre
// instead of the [T,...] head outputs (udal_run, serving configuration; heads_fused.cu)
static int heads_sample_impl(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                             float* const* cls_out, float* const* box_out, const udal_prenms_out* fused_pre) {
  UDAL_REQUIRE(ctx->heads[0].set && ctx->heads[1].set, "head weights not set (udal_set_head_weights)");
  UDAL_REQUIRE(batch > 0, "batch must be positive");
  const udal_config& c = ctx->cfg;
  for (int l = 0; l < c.num_levels; ++l)
    UDAL_REQUIRE(feats[l] && (fused_pre || (cls_out[l] && box_out[l])), "level %d pointer is NULL", l);
  const int T = c.mc_samples, L = c.num_levels, R = c.repeats, F = c.num_filters;
  if (udal_host_trace) udal_host_trace_mark("heads_sample_impl entry", 0);
  if (c.heads_mode != UDAL_HEADS_FP32) UDAL_TRY(udal_work_counters_reset(ctx));
  if (udal_host_trace) udal_host_trace_mark("after counters reset", 0);
  const int64_t total = (int64_t)T * 2 * L * R * batch * F;
  float* scale_raw;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_C, (size_t)total * 4 * 2, (void**)&scale_raw));
  float* scale = scale_raw + total;
  if (c.cls_mc || c.box_mc) {
    dropout_scale_kernel<<<(int)((total + 255) / 256), 256, 0, ctx->stream>>>(
        scale_raw, keep_masks, total, (int64_t)L * R * batch * F, c.rate_class, c.rate_box, c.inv_keep_class,
        c.inv_keep_box, seed);
    UDAL_CHECK_LAUNCH(ctx);
    scale_transpose_kernel<<<(int)((total + 255) / 256), 256, 0, ctx->stream>>>(scale_raw, scale, T, L, R, batch, F);
    UDAL_CHECK_LAUNCH(ctx);
  }
  if (c.heads_mode == UDAL_HEADS_FP32X3_TC) {
    UDAL_REQUIRE(!fused_pre, "the fused predict + decode kernels need heads_mode fp16 | bf16");
    return udal_heads_x3_sample(ctx, feats, batch, scale, cls_out, box_out);
  }
  if (c.heads_mode != UDAL_HEADS_FP32) return udal_heads_tc_sample(ctx, feats, batch, scale, cls_out, box_out, fused_pre);
  UDAL_REQUIRE(!fused_pre, "the fused predict + decode kernels need a tensor-core heads mode (fp16 | bf16)");
  UDAL_TRY(run_tower_fp32(ctx, UDAL_HEAD_CLASS, feats, batch, scale, cls_out));
  UDAL_TRY(run_tower_fp32(ctx, UDAL_HEAD_BOX, feats, batch, scale, box_out));
  return UDAL_OK;
}

extern "C" int udal_heads_sample(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks,
                                 uint64_t seed, float* const* cls_out, float* const* box_out) {
  UDAL_REQUIRE(ctx && feats && cls_out && box_out, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  return heads_sample_impl(ctx, feats, batch, keep_masks, seed, cls_out, box_out, nullptr);
}

int udal_heads_sample_fused(udal_ctx* ctx, const float* const* feats, int batch, const uint8_t* keep_masks, uint64_t seed,
                            const udal_prenms_out* pre) {
  return heads_sample_impl(ctx, feats, batch, keep_masks, seed, nullptr, nullptr, pre);
}
