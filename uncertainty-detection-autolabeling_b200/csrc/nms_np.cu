// The reference's NumPy NMS family (src/nms_np.py) on the device.
//
// Replaces nms_np.py:30-89 diou_nms, :92-129 hard_nms, :132-194 soft_nms (linear / gaussian /
// traditional), all with the "+1" pixel convention of that file, fp32 arithmetic in NumPy's
// evaluation order.  One CTA per call (the reference calls these once per image and class on
// <= 5000 candidates); the soft variant emulates the in-place array algorithm literally
// (arg-max, swap with row 0, decay, stable compaction) so the visiting order - and with it the
// product order of the decay weights - is the reference's.
#include <math_constants.h>

#include "udal_common.cuh"

namespace {

constexpr int kT = 1024;
constexpr int kSmemN = 16384;  // sort keys + alive flags of up to this many boxes live in shared memory, more in global scratch

enum { NP_HARD = 0, NP_DIOU = 1, NP_LINEAR = 2, NP_GAUSSIAN = 3, NP_SOFT_HARD = 4 };

struct Det {
  float x1, y1, x2, y2, s;
};

__device__ __forceinline__ float inter_p1(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2,
                                          float by2) {
  const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 1.0f));
  const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 1.0f));
  return __fmul_rn(w, h);
}

// greedy family: hard / diou.  order[] = indices by score descending (ties: higher index first)
__global__ void __launch_bounds__(kT) nms_np_greedy_kernel(const float* __restrict__ dets, int n, int diou, float thr,
                                                           unsigned long long* __restrict__ gkeys,
                                                           float* __restrict__ kept, int* __restrict__ num_kept) {
  extern __shared__ unsigned long long skeys[];  // [p2] then alive bytes (n <= kSmemN; larger inputs: gkeys, global)
  __shared__ int sh_p, sh_cnt;
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  unsigned long long* keys = gkeys ? gkeys : skeys;
  unsigned char* alive = reinterpret_cast<unsigned char*>(keys + p2);
  const int tid = threadIdx.x;
  for (int i = tid; i < p2; i += kT) {
    unsigned long long k = 0ull;
    if (i < n) k = ((unsigned long long)udal_float_key(dets[i * 5 + 4] + 0.0f) << 32) | (unsigned int)(i + 1);
    keys[i] = k;
  }
  __syncthreads();
  for (unsigned int size = 2; size <= (unsigned int)p2; size <<= 1)
    for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned int i = tid; i < ((unsigned int)p2 >> 1); i += kT) {
        const unsigned int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a < c) == desc) {
          keys[lo] = c;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < n; i += kT) alive[i] = 1;
  if (tid == 0) {
    sh_p = 0;
    sh_cnt = 0;
  }
  __syncthreads();
  while (true) {
    if (tid == 0) {
      int p = sh_p;
      while (p < n && !alive[p]) ++p;
      sh_p = p;
    }
    __syncthreads();
    const int p = sh_p;
    if (p >= n) break;
    const int i = (int)(keys[p] & 0xffffffffull) - 1;
    const float ix1 = dets[i * 5], iy1 = dets[i * 5 + 1], ix2 = dets[i * 5 + 2], iy2 = dets[i * 5 + 3];
    const float ia = __fmul_rn(__fadd_rn(__fsub_rn(ix2, ix1), 1.0f), __fadd_rn(__fsub_rn(iy2, iy1), 1.0f));
    if (tid == 0) {
      for (int c = 0; c < 5; ++c) kept[sh_cnt * 5 + c] = dets[i * 5 + c];
      ++sh_cnt;
      alive[p] = 0;
    }
    for (int q = p + 1 + tid; q < n; q += kT) {
      if (!alive[q]) continue;
      const int j = (int)(keys[q] & 0xffffffffull) - 1;
      const float jx1 = dets[j * 5], jy1 = dets[j * 5 + 1], jx2 = dets[j * 5 + 2], jy2 = dets[j * 5 + 3];
      const float ja = __fmul_rn(__fadd_rn(__fsub_rn(jx2, jx1), 1.0f), __fadd_rn(__fsub_rn(jy2, jy1), 1.0f));
      const float inter = inter_p1(ix1, iy1, ix2, iy2, jx1, jy1, jx2, jy2);
      float metric = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, ja), inter));
      if (diou) {
        const float ex1 = fminf(ix1, jx1), ex2 = fmaxf(ix2, jx2), ey1 = fminf(iy1, jy1), ey2 = fmaxf(iy2, jy2);
        const float dx = __fsub_rn(ex2, ex1), dy = __fsub_rn(ey2, ey1);
        const float diag = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float cx = __fsub_rn(__fmul_rn(__fadd_rn(ix1, ix2), 0.5f), __fmul_rn(__fadd_rn(jx1, jx2), 0.5f));
        const float cy = __fsub_rn(__fmul_rn(__fadd_rn(iy1, iy2), 0.5f), __fmul_rn(__fadd_rn(jy1, jy2), 0.5f));
        const float dist = __fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy));
        metric = __fsub_rn(metric, __fdiv_rn(dist, __fadd_rn(diag, 1e-10f)));
      }
      if (!(metric <= thr)) alive[q] = 0;
    }
    __syncthreads();
  }
  if (tid == 0) *num_kept = sh_cnt;
}

// soft family, literal array emulation.  work buffers: [2][n][6] floats (x1,y1,x2,y2,score,area)
__global__ void __launch_bounds__(kT) nms_np_soft_kernel(const float* __restrict__ dets, int n, int method, float thr,
                                                         float sigma, float score_thr, float* __restrict__ work,
                                                         float* __restrict__ kept, int* __restrict__ num_kept) {
  __shared__ float red_s[kT / 32];
  __shared__ int red_i[kT / 32];
  __shared__ int sh_top, sh_scan[kT / 32], sh_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* cur = work;
  float* nxt = work + (size_t)n * 6;
  for (int i = tid; i < n; i += kT) {
    const float x1 = dets[i * 5], y1 = dets[i * 5 + 1], x2 = dets[i * 5 + 2], y2 = dets[i * 5 + 3];
    cur[i * 6 + 0] = x1;
    cur[i * 6 + 1] = y1;
    cur[i * 6 + 2] = x2;
    cur[i * 6 + 3] = y2;
    cur[i * 6 + 4] = dets[i * 5 + 4];
    cur[i * 6 + 5] = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.0f), __fadd_rn(__fsub_rn(y2, y1), 1.0f));
  }
  __syncthreads();
  int m = n, cnt = 0;
  while (m > 0) {
    // argmax (first occurrence)
    float bs = -CUDART_INF_F;
    int bi = 0x7fffffff;
    bool have = false;
    for (int i = tid; i < m; i += kT) {
      const float s = cur[i * 6 + 4];
      if (!have || s > bs) {  // NaN never wins unless it is the first element (np.argmax returns the first NaN; not modelled)
        bs = s;
        bi = i;
        have = true;
      }
    }
    if (!have) bi = 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || os > bs || (os == bs && oi < bi))) {
        bs = os;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_s[warp] = bs;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kT / 32; ++w)
        if (red_i[w] != 0x7fffffff && (bi == 0x7fffffff || red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi))) {
          bs = red_s[w];
          bi = red_i[w];
        }
      sh_top = bi;
      sh_base = 0;
    }
    __syncthreads();
    const int top = sh_top;
    if (tid < 6 && top != 0) {  // dets[[0, top]] = dets[[top, 0]]
      const float a = cur[tid], b = cur[top * 6 + tid];
      cur[tid] = b;
      cur[top * 6 + tid] = a;
    }
    __syncthreads();
    const float x1 = cur[0], y1 = cur[1], x2 = cur[2], y2 = cur[3], a0 = cur[5];
    if (tid < 5) kept[cnt * 5 + tid] = cur[tid];
    ++cnt;
    // decay rows 1..m-1 and compact the survivors (stable) into nxt
    for (int c0 = 1; c0 < m; c0 += kT) {
      const int i = c0 + tid;
      bool keep = false;
      float row[6];
      if (i < m) {
#pragma unroll
        for (int c = 0; c < 6; ++c) row[c] = cur[i * 6 + c];
        const float inter = inter_p1(x1, y1, x2, y2, row[0], row[1], row[2], row[3]);
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a0, row[5]), inter));
        float wgt;
        if (method == NP_LINEAR) wgt = iou > thr ? __fsub_rn(1.0f, iou) : 1.0f;
        else if (method == NP_GAUSSIAN) wgt = expf(__fdiv_rn(-__fmul_rn(iou, iou), sigma));
        else wgt = iou > thr ? 0.0f : 1.0f;
        row[4] = __fmul_rn(row[4], wgt);
        keep = row[4] >= score_thr;
      }
      const unsigned int mask = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) sh_scan[warp] = __popc(mask);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < kT / 32; ++w) {
        const int c = sh_scan[w];
        if (w < warp) before += c;
        total += c;
      }
      const int base = sh_base;
      if (keep) {
        const int dst = base + before + __popc(mask & ((1u << lane) - 1u));
#pragma unroll
        for (int c = 0; c < 6; ++c) nxt[dst * 6 + c] = row[c];
      }
      __syncthreads();
      if (tid == 0) sh_base = base + total;
      __syncthreads();
    }
    m = m > 1 ? sh_base : 0;
    float* t = cur;
    cur = nxt;
    nxt = t;
    __syncthreads();
  }
  if (tid == 0) *num_kept = cnt;
}

}  // namespace

extern "C" int udal_nms_np(udal_ctx* ctx, const float* dets_host, int n, int method, float iou_thresh, float sigma,
                           float score_thresh, float* kept_host, int32_t* num_kept) {
  UDAL_REQUIRE(ctx && num_kept, "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(method >= NP_HARD && method <= NP_SOFT_HARD, "Unknown NMS method: %d", method);
  UDAL_REQUIRE(n >= 0 && n <= (1 << 24), "nms_np: %d boxes", n);
  *num_kept = 0;
  if (n == 0) return UDAL_OK;
  UDAL_REQUIRE(dets_host && kept_host, "NULL argument");
  float* buf;
  UDAL_TRY(udal_scratch_get(ctx, SCR_MISC, (size_t)n * (5 + 5 + 12) * sizeof(float) + 16, (void**)&buf));
  float* d_dets = buf;
  float* d_kept = buf + (size_t)n * 5;
  float* d_work = buf + (size_t)n * 10;
  int* d_cnt = reinterpret_cast<int*>(buf + (size_t)n * 22);
  UDAL_CUDA(cudaMemcpyAsync(d_dets, dets_host, (size_t)n * 5 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (method == NP_HARD || method == NP_DIOU) {
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    size_t smem = (size_t)p2 * 8 + (size_t)n + 16;
    unsigned long long* gkeys = nullptr;
    if (n > kSmemN) {  // the reference has no size limit: sort keys and alive flags in global scratch
      UDAL_TRY(udal_scratch_get(ctx, SCR_NMS_A, smem, (void**)&gkeys));
      smem = 0;
    }
    if (smem > 48 * 1024)
      UDAL_CUDA(cudaFuncSetAttribute(nms_np_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_np_greedy_kernel<<<1, kT, smem, ctx->stream>>>(d_dets, n, method == NP_DIOU ? 1 : 0, iou_thresh, gkeys, d_kept, d_cnt);
  } else {
    nms_np_soft_kernel<<<1, kT, 0, ctx->stream>>>(d_dets, n, method, iou_thresh, sigma, score_thresh, d_work, d_kept,
                                                  d_cnt);
  }
  UDAL_CHECK_LAUNCH(ctx);
  int cnt = 0;
  UDAL_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cnt > 0)
    UDAL_CUDA(cudaMemcpyAsync(kept_host, d_kept, (size_t)cnt * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  *num_kept = cnt;
  return UDAL_OK;
}
