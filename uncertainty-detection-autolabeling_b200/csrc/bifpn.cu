// BiFPN (the producer of the hot path's input): device primitives of FPNCells.
//
// Replaces (reference src/efficientdet_keras.py): ResampleFeatureMap 238-350 (1x1 conv + BN when the channel count
// differs, max / average pooling with TF "SAME" padding, nearest-neighbour up-sampling), FNode.fuse_features 86-125
// (fastattn / attn / sum and their per-channel variants) and OpAfterCombine 176-236 (swish -> separable 3x3 conv ->
// BN).  The graph itself (which node reads which, fpn_configs.bifpn_config) is host logic and lives in bifpn.py, as it
// does in the reference.  fp32 in and out; 64-filter nodes run the separable conv on the tensor cores (udal_sepconv_tc,
// heads_wide.cu: the fp32-accurate fp32x3 tower kernel on one map), other widths the fp32 tower kernel of heads_fp32.cu.
//
//   udal_conv1x1_bn   [NB,H,W,Cin] -> [NB,H,W,F]: x @ w + b, then BN (scale, shift) when given
//   udal_bifpn_fuse   out = act( sum_i weight_i * resample_i(in_i) ), resample_i chosen by the source / target sizes:
//                     same size - copy; larger source - pooling window (stride s = (in-1)/out+1, size s+1, SAME padding:
//                     pad_before = total/2, padded cells never win a max and do not count in an average);
//                     smaller source - nearest neighbour, src = min(floor(dst * (float)in / out), in - 1)
//                     (tf.compat.v1.image.resize_nearest_neighbor, align_corners = False)
//   udal_sepconv_bn   depthwise 3x3 SAME -> pointwise + bias -> BN               (heads_fp32.cu kernel)
#include "udal_common.cuh"

int udal_sepconv_fp32_launch(udal_ctx* ctx, const float* in, int NB, int H, int W, int F, int Cout, const float* dw,
                             const float* pw, const float* bias, const float* bn_scale, const float* bn_shift, int act,
                             float* out);

namespace {

__device__ __forceinline__ float swish_f(float x) { return __fdiv_rn(x, 1.f + expf(-x)); }

__global__ void conv1x1_bn_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                  const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int64_t pixels,
                                  int Cin, int F, float* __restrict__ out) {
  // block = 8 pixels x F outputs (F <= 512 handled by the stride loop); the pixel rows are staged in shared memory
  extern __shared__ float s_in[];  // [8][Cin]
  const int64_t p0 = (int64_t)blockIdx.x * 8;
  const int npx = (int)min((int64_t)8, pixels - p0);
  for (int e = threadIdx.x; e < npx * Cin; e += blockDim.x) s_in[e] = __ldg(in + p0 * Cin + e);
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = 0; k < Cin; ++k) {
      const float wk = __ldg(w + (size_t)k * F + f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(i < npx ? s_in[i * Cin + k] : 0.f, wk, acc[i]);
    }
    const float b = bias ? __ldg(bias + f) : 0.f;
    const float sc = bn_scale ? __ldg(bn_scale + f) : 1.f, sh = bn_scale ? __ldg(bn_shift + f) : 0.f;
    for (int i = 0; i < npx; ++i) {
      float v = acc[i] + b;
      if (bn_scale) v = fmaf(v, sc, sh);
      out[(p0 + i) * F + f] = v;
    }
  }
}

// Register-tiled form for F = 64 (the D0 BiFPN), Cin % 4 == 0, 16-byte aligned maps: block = 128 pixels, thread = 8 pixels x 4
// outputs (32 accumulators); the input rows go through shared memory in chunks of 32 channels, per 4 channels a thread issues
// 8 broadcast LDS.128 + 4 LDG.128 (weights, L1 resident) for 128 FMAs (the kernel above: 1 load per FMA).
constexpr int kC1Px = 128, kC1K = 32;
__global__ void __launch_bounds__(256) conv1x1_bn64_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ bn_scale,
                                                          const float* __restrict__ bn_shift, int64_t pixels, int Cin,
                                                          float* __restrict__ out) {
  __shared__ __align__(16) float s_in[kC1Px * kC1K];   // [128 px][32 k]
  const int tid = threadIdx.x, g = tid >> 4, t = tid & 15;   // group of 8 pixels, thread's 4 outputs 4 t ..
  const int64_t p0 = (int64_t)blockIdx.x * kC1Px;
  const int npx = (int)min((int64_t)kC1Px, pixels - p0);
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < Cin; k0 += kC1K) {
    const int kn = min(kC1K, Cin - k0);   // multiple of 4
    __syncthreads();
    for (int e = tid; e < kC1Px * (kC1K / 4); e += 256) {
      const int px = e >> 3, k4 = (e & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < npx && k4 < kn) v = __ldg(reinterpret_cast<const float4*>(in + (p0 + px) * Cin + k0 + k4));
      *reinterpret_cast<float4*>(s_in + px * kC1K + k4) = v;
    }
    __syncthreads();
    for (int kk = 0; kk < kn; kk += 4) {
      float4 wk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wk[j] = __ldg(reinterpret_cast<const float4*>(w + (size_t)(k0 + kk + j) * 64) + t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(s_in + (g * 8 + i) * kC1K + kk);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i].x = fmaf(xv[j], wk[j].x, acc[i].x);
          acc[i].y = fmaf(xv[j], wk[j].y, acc[i].y);
          acc[i].z = fmaf(xv[j], wk[j].z, acc[i].z);
          acc[i].w = fmaf(xv[j], wk[j].w, acc[i].w);
        }
      }
    }
  }
  const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bn_scale) {
    sc = __ldg(reinterpret_cast<const float4*>(bn_scale) + t);
    sh = __ldg(reinterpret_cast<const float4*>(bn_shift) + t);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int px = g * 8 + i;
    if (px >= npx) continue;
    float4 v = make_float4(acc[i].x + b.x, acc[i].y + b.y, acc[i].z + b.z, acc[i].w + b.w);
    if (bn_scale) v = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    *reinterpret_cast<float4*>(out + (p0 + px) * 64 + 4 * t) = v;
  }
}

struct FuseParams {
  int n;                       // inputs (1..3)
  const float* in[3];          // [NB, h_i, w_i, F]
  int h[3], w[3];
  const float* wsm[3];         // edge weight of input i: scalar or [F] (per_channel); null for "sum"
  int mode;                    // UDAL_FUSE_*
  int per_channel;
  int pool_avg;
  int NB, H, W, F;
  int act;                     // 1: swish on the fused value (OpAfterCombine applies it before its conv)
  float* out;                  // [NB,H,W,F]
};

__device__ __forceinline__ float resample_at(const float* __restrict__ src, int h, int w, int H, int W, int y, int x, int F,
                                             int f, int pool_avg) {
  if (h == H && w == W) return __ldg(src + ((size_t)y * w + x) * F + f);
  if (h > H && w > W) {
    // pooling, TF "SAME": out = ceil(in / s) must equal the target (checked on the host)
    const int sy = (h - 1) / H + 1, sx = (w - 1) / W + 1, ky = sy + 1, kx = sx + 1;
    const int pad_y = max((H - 1) * sy + ky - h, 0) / 2, pad_x = max((W - 1) * sx + kx - w, 0) / 2;
    const int y0 = y * sy - pad_y, x0 = x * sx - pad_x;
    float best = -3.402823466e38f, sum = 0.f;
    int cnt = 0;
    for (int dy = 0; dy < ky; ++dy) {
      const int yy = y0 + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = 0; dx < kx; ++dx) {
        const int xx = x0 + dx;
        if (xx < 0 || xx >= w) continue;
        const float v = __ldg(src + ((size_t)yy * w + xx) * F + f);
        best = fmaxf(best, v);
        sum += v;
        ++cnt;
      }
    }
    return pool_avg ? __fdiv_rn(sum, (float)cnt) : best;
  }
  // nearest neighbour (h <= H and w <= W)
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int yy = min((int)floorf((float)y * sy), h - 1), xx = min((int)floorf((float)x * sx), w - 1);
  return __ldg(src + ((size_t)yy * w + xx) * F + f);
}

__global__ void bifpn_fuse_kernel(const FuseParams p) {
  const int64_t total = (int64_t)p.NB * p.H * p.W * p.F;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % p.F);
  int64_t r = i / p.F;
  const int x = (int)(r % p.W);
  r /= p.W;
  const int y = (int)(r % p.H);
  const int nb = (int)(r / p.H);
  float v[3], ew[3];
  for (int k = 0; k < p.n; ++k) {
    v[k] = resample_at(p.in[k] + (size_t)nb * p.h[k] * p.w[k] * p.F, p.h[k], p.w[k], p.H, p.W, y, x, p.F, f, p.pool_avg);
    ew[k] = p.wsm[k] ? __ldg(p.wsm[k] + (p.per_channel ? f : 0)) : 1.f;
  }
  float acc;
  if (p.mode == UDAL_FUSE_FASTATTN) {
    // nodes[i] * relu(w_i) / (sum_j relu(w_j) + 0.0001), then add_n (left to right)
    float ws = 0.f;
    for (int k = 0; k < p.n; ++k) {
      ew[k] = fmaxf(ew[k], 0.f);
      ws = k == 0 ? ew[0] : __fadd_rn(ws, ew[k]);
    }
    const float den = __fadd_rn(ws, 0.0001f);
    acc = __fdiv_rn(__fmul_rn(v[0], ew[0]), den);
    for (int k = 1; k < p.n; ++k) acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(v[k], ew[k]), den));
  } else if (p.mode == UDAL_FUSE_ATTN) {
    // softmax over the edge weights, reduce_sum(nodes * w)
    float mx = ew[0];
    for (int k = 1; k < p.n; ++k) mx = fmaxf(mx, ew[k]);
    float e[3], se = 0.f;
    for (int k = 0; k < p.n; ++k) {
      e[k] = expf(ew[k] - mx);
      se += e[k];
    }
    acc = __fmul_rn(v[0], __fdiv_rn(e[0], se));
    for (int k = 1; k < p.n; ++k) acc = __fadd_rn(acc, __fmul_rn(v[k], __fdiv_rn(e[k], se)));
  } else {
    acc = v[0];
    for (int k = 1; k < p.n; ++k) acc = __fadd_rn(acc, v[k]);
  }
  p.out[i] = p.act ? swish_f(acc) : acc;
}

// The same fusion, four channels per thread (F % 4 == 0, 16-byte aligned maps): grid = (16-pixel row segments, rows, images)
// - no index division at all -, the resampling rule of every input resolved on the host (Resample: same / pooling window /
// nearest scale), 16-byte loads and stores, the edge weights normalised once per thread (w_i / (sum + 0.0001) resp. the
// softmax) and applied by FMA, swish through ex2 / rcp.  The scalar kernel above divides every value by the weight sum as the
// reference does (IEEE); this one differs from it by a few ulp per node - the contract of this row is the oracle's 2e-4.
// (The scalar kernel took 3.9 ms of an 8.4 ms FPNCells call at D0 1280x384, batch 64; HBM time of its bytes: 0.3 ms.)
struct Resample {
  int kind;                    // 0 same size, 1 pooling, 2 nearest neighbour
  int sy, sx, ky, kx, pad_y, pad_x;
  float fy, fx;                // nearest: (float)h / H, (float)w / W
};
struct Fuse4Params {
  FuseParams f;
  Resample rs[3];
};

__device__ __forceinline__ float4 resample_at4(const float* __restrict__ src, const Resample& rs, int h, int w, int y, int x, int F,
                                               int f, int pool_avg) {
  if (rs.kind == 0) return __ldg(reinterpret_cast<const float4*>(src + ((size_t)y * w + x) * F + f));
  if (rs.kind == 1) {
    const int y0 = y * rs.sy - rs.pad_y, x0 = x * rs.sx - rs.pad_x;
    float4 best = make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int dy = 0; dy < rs.ky; ++dy) {
      const int yy = y0 + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = 0; dx < rs.kx; ++dx) {
        const int xx = x0 + dx;
        if (xx < 0 || xx >= w) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)yy * w + xx) * F + f));
        best = make_float4(fmaxf(best.x, v.x), fmaxf(best.y, v.y), fmaxf(best.z, v.z), fmaxf(best.w, v.w));
        sum = make_float4(sum.x + v.x, sum.y + v.y, sum.z + v.z, sum.w + v.w);
        ++cnt;
      }
    }
    if (!pool_avg) return best;
    const float fc = (float)cnt;
    return make_float4(__fdiv_rn(sum.x, fc), __fdiv_rn(sum.y, fc), __fdiv_rn(sum.z, fc), __fdiv_rn(sum.w, fc));
  }
  const int yy = min((int)floorf((float)y * rs.fy), h - 1), xx = min((int)floorf((float)x * rs.fx), w - 1);
  return __ldg(reinterpret_cast<const float4*>(src + ((size_t)yy * w + xx) * F + f));
}

__device__ __forceinline__ float swish_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

// normalised edge weights of a node with scalar weights (the reference default): one division per thread, not four
__device__ __forceinline__ void fuse_weights_scalar(const FuseParams& p, float (&wn)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) wn[k] = (k < p.n && p.wsm[k]) ? __ldg(p.wsm[k]) : 1.f;
  if (p.mode == UDAL_FUSE_FASTATTN) {
    float ws = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (k < p.n) {
        wn[k] = fmaxf(wn[k], 0.f);
        ws += wn[k];
      }
    const float inv = __fdiv_rn(1.f, ws + 0.0001f);
#pragma unroll
    for (int k = 0; k < 3; ++k) wn[k] *= inv;
  } else if (p.mode == UDAL_FUSE_ATTN) {
    float mx = wn[0];
#pragma unroll
    for (int k = 1; k < 3; ++k)
      if (k < p.n) mx = fmaxf(mx, wn[k]);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (k < p.n) {
        wn[k] = expf(wn[k] - mx);
        se += wn[k];
      }
    const float inv = __fdiv_rn(1.f, se);
#pragma unroll
    for (int k = 0; k < 3; ++k) wn[k] *= inv;
  }
}

__global__ void __launch_bounds__(256, 4) bifpn_fuse4_kernel(const Fuse4Params q) {
  const FuseParams& p = q.f;
  const int F4 = p.F >> 2;
  const int x = blockIdx.x * 16 + (threadIdx.x >> 4);   // 16 threads (64 channels per pass) per pixel
  const int y = blockIdx.y, nb = blockIdx.z;
  if (x >= p.W) return;
  const size_t px = ((size_t)nb * p.H + y) * p.W + x;
  float ws[3] = {1.f, 1.f, 1.f};
  if (!p.per_channel) fuse_weights_scalar(p, ws);
  for (int f4 = threadIdx.x & 15; f4 < F4; f4 += 16) {
    const int f = 4 * f4;
    float4 v[3], wn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k >= p.n) break;
      v[k] = resample_at4(p.in[k] + (size_t)nb * p.h[k] * p.w[k] * p.F, q.rs[k], p.h[k], p.w[k], y, x, p.F, f, p.pool_avg);
      wn[k] = make_float4(ws[k], ws[k], ws[k], ws[k]);
    }
    if (p.per_channel) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k < p.n) wn[k] = p.wsm[k] ? __ldg(reinterpret_cast<const float4*>(p.wsm[k] + f)) : make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.mode == UDAL_FUSE_FASTATTN) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k >= p.n) break;
          wn[k] = make_float4(fmaxf(wn[k].x, 0.f), fmaxf(wn[k].y, 0.f), fmaxf(wn[k].z, 0.f), fmaxf(wn[k].w, 0.f));
          sum = make_float4(sum.x + wn[k].x, sum.y + wn[k].y, sum.z + wn[k].z, sum.w + wn[k].w);
        }
        const float4 inv = make_float4(__fdiv_rn(1.f, sum.x + 0.0001f), __fdiv_rn(1.f, sum.y + 0.0001f), __fdiv_rn(1.f, sum.z + 0.0001f),
                                       __fdiv_rn(1.f, sum.w + 0.0001f));
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < p.n) wn[k] = make_float4(wn[k].x * inv.x, wn[k].y * inv.y, wn[k].z * inv.z, wn[k].w * inv.w);
      } else if (p.mode == UDAL_FUSE_ATTN) {
        float4 mx = wn[0];
#pragma unroll
        for (int k = 1; k < 3; ++k)
          if (k < p.n) mx = make_float4(fmaxf(mx.x, wn[k].x), fmaxf(mx.y, wn[k].y), fmaxf(mx.z, wn[k].z), fmaxf(mx.w, wn[k].w));
        float4 se = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k >= p.n) break;
          wn[k] = make_float4(expf(wn[k].x - mx.x), expf(wn[k].y - mx.y), expf(wn[k].z - mx.z), expf(wn[k].w - mx.w));
          se = make_float4(se.x + wn[k].x, se.y + wn[k].y, se.z + wn[k].z, se.w + wn[k].w);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < p.n) wn[k] = make_float4(__fdiv_rn(wn[k].x, se.x), __fdiv_rn(wn[k].y, se.y), __fdiv_rn(wn[k].z, se.z), __fdiv_rn(wn[k].w, se.w));
      }
    }
    float4 acc = make_float4(v[0].x * wn[0].x, v[0].y * wn[0].y, v[0].z * wn[0].z, v[0].w * wn[0].w);
#pragma unroll
    for (int k = 1; k < 3; ++k)
      if (k < p.n) acc = make_float4(fmaf(v[k].x, wn[k].x, acc.x), fmaf(v[k].y, wn[k].y, acc.y), fmaf(v[k].z, wn[k].z, acc.z), fmaf(v[k].w, wn[k].w, acc.w));
    if (p.act) acc = make_float4(swish_fast(acc.x), swish_fast(acc.y), swish_fast(acc.z), swish_fast(acc.w));
    *reinterpret_cast<float4*>(p.out + px * p.F + f) = acc;
  }
}

}  // namespace

extern "C" {

int udal_conv1x1_bn(udal_ctx* ctx, const float* in, int NB, int H, int W, int Cin, const float* w, const float* bias,
                    const float* bn_scale, const float* bn_shift, int F, float* out) {
  UDAL_REQUIRE(ctx && in && w && out, "NULL argument");
  UDAL_REQUIRE(NB > 0 && H > 0 && W > 0 && Cin > 0 && F > 0, "udal_conv1x1_bn: bad sizes");
  UDAL_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "udal_conv1x1_bn: BN scale and shift go together");
  UDAL_REQUIRE((size_t)8 * Cin * sizeof(float) <= 48 * 1024, "udal_conv1x1_bn: %d input channels", Cin);
  UDAL_TRY(udal_join(ctx));
  const int64_t pixels = (int64_t)NB * H * W;
  const uintptr_t al = (uintptr_t)in | (uintptr_t)w | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)bn_scale | (uintptr_t)bn_shift;
  if (F == 64 && (Cin & 3) == 0 && (al & 15) == 0) {
    conv1x1_bn64_kernel<<<(unsigned)((pixels + kC1Px - 1) / kC1Px), 256, 0, ctx->stream>>>(in, w, bias, bn_scale, bn_shift, pixels,
                                                                                          Cin, out);
    UDAL_CHECK_LAUNCH(ctx);
    return UDAL_OK;
  }
  const int threads = F >= 128 ? 128 : 64;
  conv1x1_bn_kernel<<<(unsigned)((pixels + 7) / 8), threads, (size_t)8 * Cin * sizeof(float), ctx->stream>>>(
      in, w, bias, bn_scale, bn_shift, pixels, Cin, F, out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_bifpn_fuse(udal_ctx* ctx, int n, const float* const* in, const int* in_h, const int* in_w, const float* const* wsm,
                    int mode, int per_channel, int pool_avg, int NB, int H, int W, int F, int act, float* out) {
  UDAL_REQUIRE(ctx && in && in_h && in_w && out, "NULL argument");
  UDAL_REQUIRE(n >= 1 && n <= 3, "udal_bifpn_fuse: %d inputs (1..3)", n);
  UDAL_REQUIRE(mode == UDAL_FUSE_SUM || mode == UDAL_FUSE_FASTATTN || mode == UDAL_FUSE_ATTN, "unknown weight_method %d", mode);
  UDAL_REQUIRE(NB > 0 && H > 0 && W > 0 && F > 0, "udal_bifpn_fuse: bad sizes");
  FuseParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  for (int k = 0; k < n; ++k) {
    UDAL_REQUIRE(in[k], "udal_bifpn_fuse: input %d is NULL", k);
    const int h = in_h[k], w = in_w[k];
    const bool same = h == H && w == W, down = h > H && w > W, up = h <= H && w <= W;
    // efficientdet_keras.py:345-349
    UDAL_REQUIRE(same || down || up, "Incompatible Resampling : feat shape %dx%d target_shape: %dx%d", h, w, H, W);
    if (down) {
      const int sy = (h - 1) / H + 1, sx = (w - 1) / W + 1;
      UDAL_REQUIRE((h + sy - 1) / sy == H && (w + sx - 1) / sx == W,
                   "pooling %dx%d with stride %dx%d does not give the target %dx%d", h, w, sy, sx, H, W);
    }
    p.in[k] = in[k];
    p.h[k] = h;
    p.w[k] = w;
    p.wsm[k] = (wsm && mode != UDAL_FUSE_SUM) ? wsm[k] : nullptr;
    UDAL_REQUIRE(mode == UDAL_FUSE_SUM || p.wsm[k], "udal_bifpn_fuse: edge weight %d is NULL", k);
  }
  p.mode = mode;
  p.per_channel = per_channel;
  p.pool_avg = pool_avg;
  p.NB = NB;
  p.H = H;
  p.W = W;
  p.F = F;
  p.act = act;
  p.out = out;
  UDAL_TRY(udal_join(ctx));
  const int64_t total = (int64_t)NB * H * W * F;
  bool vec = (F & 3) == 0 && ((uintptr_t)out & 15) == 0 && (int64_t)NB * H * W < (1ll << 31);
  for (int k = 0; k < n; ++k) vec = vec && ((uintptr_t)in[k] & 15) == 0 && (!p.wsm[k] || !per_channel || ((uintptr_t)p.wsm[k] & 15) == 0);
  vec = vec && H <= 65535 && NB <= 65535;
  if (vec) {
    Fuse4Params q;
    q.f = p;
    for (int k = 0; k < n; ++k) {
      Resample& r = q.rs[k];
      memset(&r, 0, sizeof(r));
      const int h = p.h[k], w = p.w[k];
      if (h == H && w == W) r.kind = 0;
      else if (h > H && w > W) {
        r.kind = 1;
        r.sy = (h - 1) / H + 1;
        r.sx = (w - 1) / W + 1;
        r.ky = r.sy + 1;
        r.kx = r.sx + 1;
        const int ty = (H - 1) * r.sy + r.ky - h, tx = (W - 1) * r.sx + r.kx - w;
        r.pad_y = (ty > 0 ? ty : 0) / 2;
        r.pad_x = (tx > 0 ? tx : 0) / 2;
      } else {
        r.kind = 2;
        r.fy = (float)h / (float)H;
        r.fx = (float)w / (float)W;
      }
    }
    bifpn_fuse4_kernel<<<dim3((unsigned)((W + 15) / 16), (unsigned)H, (unsigned)NB), 256, 0, ctx->stream>>>(q);
  } else {
    bifpn_fuse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p);
  }
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_sepconv_bn(udal_ctx* ctx, const float* in, int NB, int H, int W, int F, int Cout, const float* dw, const float* pw,
                    const float* bias, const float* bn_scale, const float* bn_shift, int act, float* out) {
  UDAL_REQUIRE(ctx && in && dw && pw && bias && out, "NULL argument");
  UDAL_REQUIRE(NB > 0 && H > 0 && W > 0 && F > 0 && Cout > 0, "udal_sepconv_bn: bad sizes");
  UDAL_REQUIRE(act == UDAL_ACT_NONE || (bn_scale && bn_shift), "udal_sepconv_bn: BN tables missing");
  UDAL_TRY(udal_join(ctx));
  return udal_sepconv_fp32_launch(ctx, in, NB, H, W, F, Cout, dw, pw, bias, bn_scale, bn_shift, act, out);
}

}  // extern "C"
