// K2: MC moments of the class logits + anchor decode with exact moment propagation + MC moments
// of the decoded boxes, one pass over HBM.
//
// Replaces (reference src/): utils_extra.py:220-244 get_mcuncert, postprocess.py:75-87
// merge_class_box_level_outputs, postprocess.py:123-135 max-reduce, postprocess.py:284 sigmoid,
// utils_box.py:105-276 decode_uncert, anchors.py:41-75 decode_box_outputs and the MC reductions
// of postprocess.py:297-331.
//
// Arithmetic contract (bit-for-bit what oracle/ref_np.py does, up to the last-ulp behaviour of
// exp): decode in fp64 from fp32 inputs with separately rounded mul/add, results rounded to fp32;
// MC mean = sequential fp32 sum in sample order / T; MC std = two-pass population std in fp32.
#include "fast_math64.cuh"
#include "udal_common.cuh"
#include "decode_math.cuh"

#include <algorithm>
#include <type_traits>

namespace {

constexpr int kThreads = 256;

struct DecodeParams {
  udal_level_ptrs cls;
  udal_level_ptrs box;
  udal_level_geom geom;
  int tile_off[UDAL_MAX_LEVELS + 1];  // prefix of tiles per level
  int tile_px;                         // pixels per tile
  int stage_off;                       // float offset of the sample staging area in dynamic smem
  int batch, A, C, BC;                 // BC = box channels (4A or 8A)
  int Tc, Tb;                          // samples on the class / box inputs (1 = no MC axis)
  int cls_mc, box_mc, la, method;
  const float* anchors;                // [N,4]
  int64_t N;
  udal_prenms_out out;
};

__device__ __forceinline__ float seq_mean(const float* v, int T) {
  float acc = v[0];
  for (int t = 1; t < T; ++t) acc = __fadd_rn(acc, v[t]);
  return __fdiv_rn(acc, (float)T);
}

template <int TMAX>
__device__ __forceinline__ void moments_reg(const float (&v)[TMAX], int T, float& mean, float& sd) {
  float acc = v[0];
#pragma unroll
  for (int t = 1; t < TMAX; ++t)
    if (t < T) acc = __fadd_rn(acc, v[t]);
  const float fT = (float)T;
  mean = __fdiv_rn(acc, fT);
  float d0 = __fsub_rn(v[0], mean);
  float s = __fmul_rn(d0, d0);
#pragma unroll
  for (int t = 1; t < TMAX; ++t)
    if (t < T) {
      float d = __fsub_rn(v[t], mean);
      s = __fadd_rn(s, __fmul_rn(d, d));
    }
  sd = __fsqrt_rn(__fdiv_rn(s, fT));
}

__device__ __forceinline__ int find_level(const int* off, int nl, int v) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < UDAL_MAX_LEVELS; ++i)
    if (i < nl && v >= off[i]) l = i;
  return l;
}

// -------------------------------------------------------------------------------------------
// logits: elementwise MC mean / std over a contiguous run of `count` floats.
// src(t) = base + t * t_stride; dst offsets are shared by mean/std.
// -------------------------------------------------------------------------------------------
template <int TMAX, bool TWO_PASS = false>
__device__ __forceinline__ void logits_run(const float* __restrict__ base, size_t t_stride, int T,
                                           int count, float* __restrict__ mean_out,
                                           float* __restrict__ std_out, float* smem_mean, int tid,
                                           int nthreads) {
  if (TMAX == 0 || TWO_PASS) {
    // many samples: two passes over global memory (second pass hits L1/L2).  For 17..24 samples this keeps the
    // kernel under 85 registers, i.e. 3 resident CTAs per SM instead of 2
    for (int e = tid; e < count; e += nthreads) {
      float acc = __ldg(base + e);
      for (int t = 1; t < T; ++t) acc = __fadd_rn(acc, __ldg(base + (size_t)t * t_stride + e));
      const float mean = __fdiv_rn(acc, (float)T);
      float s = 0.f;
      for (int t = 0; t < T; ++t) {
        float d = __fsub_rn(__ldg(base + (size_t)t * t_stride + e), mean);
        s = t == 0 ? __fmul_rn(d, d) : __fadd_rn(s, __fmul_rn(d, d));
      }
      const float sd = __fsqrt_rn(__fdiv_rn(s, (float)T));
      if (mean_out) mean_out[e] = mean;
      if (std_out) std_out[e] = sd;
      if (smem_mean) smem_mean[e] = mean;
    }
    return;
  }
  constexpr int TM = TMAX == 0 ? 1 : TMAX;
  constexpr int U = TM <= 16 ? 2 : 1;  // elements in flight per thread (2 x T independent loads)
  for (int e0 = tid; e0 < count; e0 += U * nthreads) {
    float v[U][TM];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * nthreads;
#pragma unroll
      for (int t = 0; t < TM; ++t)
        if (t < T && e < count) v[u][t] = __ldg(base + (size_t)t * t_stride + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * nthreads;
      if (e < count) {
        float mean, sd;
        moments_reg<TM>(v[u], T, mean, sd);
        if (mean_out) mean_out[e] = mean;
        if (std_out) std_out[e] = sd;
        if (smem_mean) smem_mean[e] = mean;
      }
    }
  }
}

// FAST = the serving configuration (loss attenuation, l-norm decode, MC dropout on both heads,
// T == TMAX): every mode switch folds away at compile time.
// TCH: samples per load chunk of the 17..32-sample path.  LEAN (17..24 samples): logits in two passes, 3 CTAs per SM.
template <int TMAX, bool FAST, int TCH_ = 4, bool LEAN = false>
__global__ void __launch_bounds__(kThreads, TMAX == 0 ? 1 : (TMAX > 16 ? (LEAN ? 3 : 2) : 4)) decode_moments_kernel(const DecodeParams p) {
  const int k_la = FAST ? 1 : p.la;
  const int k_method = FAST ? (int)UDAL_DECODE_LNORM : p.method;
  const int k_box_mc = FAST ? 1 : p.box_mc;
  const int k_cls_mc = FAST ? 1 : p.cls_mc;
  const int k_Tb = FAST ? TMAX : p.Tb;
  const int k_Tc = FAST ? TMAX : p.Tc;
  extern __shared__ float smem_mean[];  // [tile anchors * C] then the per-thread sample staging
  float* smem_stage = smem_mean + p.stage_off;
  __shared__ double smem_tbl[64];
  if (threadIdx.x < 64) smem_tbl[threadIdx.x] = kExp2Table[threadIdx.x];
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = find_level(p.tile_off, p.geom.num_levels, tile);
  const int hw = p.geom.h[l] * p.geom.w[l];
  const int p0 = (tile - p.tile_off[l]) * p.tile_px;
  const int npx = min(p.tile_px, hw - p0);
  const int A = p.A, C = p.C;
  const int64_t anchor0 = (int64_t)A * (p.geom.pix_off[l] + p0);  // first global anchor of the tile

  // ---- phase 1: class logits (elementwise) ------------------------------------------------
  {
    const int count = npx * A * C;
    const size_t plane = (size_t)hw * A * C;
    const float* base = p.cls.p[l] + ((size_t)b * hw + p0) * A * C;
    const size_t t_stride = (size_t)p.batch * plane;
    float* mo = p.out.mean_logits ? p.out.mean_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
    float* so = (p.out.std_logits && k_cls_mc) ? p.out.std_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
    logits_run<TMAX, LEAN>(base, t_stride, k_Tc, count, mo, so, smem_mean, threadIdx.x, kThreads);
  }
  __syncthreads();

  // ---- phase 2: one thread per (anchor, axis): even lanes decode y (ymin, ymax), odd lanes x ----
  const int axis = threadIdx.x & 1;
  for (int i0 = 0; i0 < npx * A; i0 += kThreads / 2) {
    const int i = i0 + (threadIdx.x >> 1);
    const bool live = i < npx * A;
    const int ii = live ? i : 0;
    const int px = p0 + ii / A;
    const int a = ii - (ii / A) * A;
    const int64_t n = anchor0 + ii;
    const float4 anc = __ldg(reinterpret_cast<const float4*>(p.anchors) + n);
    const float a_lo = axis ? anc.y : anc.x, a_hi = axis ? anc.w : anc.z;
    const size_t plane = (size_t)hw * p.BC;
    const float* bb = p.box.p[l] + ((size_t)b * hw + px) * p.BC + a * 4;
    const size_t t_stride = (size_t)p.batch * plane;
    const int T = k_Tb;
    float m_lo, m_hi, sd_lo = 0.f, sd_hi = 0.f, al_lo = 0.f, al_hi = 0.f;
    if (TMAX > 16) {
      // 17..32 samples: the loads go out in chunks of TCH samples (2 x TCH independent 16-byte requests per thread;
      // all 2 x T at once would need 256 registers), each chunk is parked in shared memory and decoded by a rolled
      // loop; the decoded corners of all samples stay in shared memory for the two-pass standard deviation.
      // (2 T + 4 TCH) floats per thread instead of 4 T, so that several CTAs fit an SM (the fp64 decode chains
      // need the warps: the kernel is issue / latency bound, not HBM bound).
      constexpr int TCH = TCH_;
      float* sd = smem_stage + threadIdx.x;                      // (lo, hi) of sample t at sd[(t * 2 + k) * kThreads]
      float* sr = smem_stage + 2 * T * kThreads + threadIdx.x;   // raw chunk: slot (u, k) at sr[(u * 4 + k) * kThreads]
      float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll 1
      for (int t0 = 0; t0 < T; t0 += TCH) {
        {
          float4 tt[TCH], sg[TCH];
#pragma unroll
          for (int u = 0; u < TCH; ++u)
            if (t0 + u < T) {
              tt[u] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)(t0 + u) * t_stride));
              if (k_la) sg[u] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)(t0 + u) * t_stride + 4 * A));
            }
#pragma unroll
          for (int u = 0; u < TCH; ++u)
            if (t0 + u < T) {
              sr[(u * 4 + 0) * kThreads] = axis ? tt[u].y : tt[u].x;
              sr[(u * 4 + 1) * kThreads] = axis ? tt[u].w : tt[u].z;
              if (k_la) {
                sr[(u * 4 + 2) * kThreads] = axis ? sg[u].y : sg[u].x;
                sr[(u * 4 + 3) * kThreads] = axis ? sg[u].w : sg[u].z;
              }
            }
        }
        const int tn = min(TCH, T - t0);
#pragma unroll 1
        for (int u = 0; u < tn; ++u) {
          const int t = t0 + u;
          const float t_c = sr[(u * 4 + 0) * kThreads], t_s = sr[(u * 4 + 1) * kThreads];
          float lo, hi;
          if (k_la) {
            float s_lo, s_hi;
            decode_axis_la(k_method, smem_tbl, a_lo, a_hi, t_c, t_s, sr[(u * 4 + 2) * kThreads], sr[(u * 4 + 3) * kThreads], lo,
                           hi, s_lo, s_hi);
            al_lo = t == 0 ? s_lo : __fadd_rn(al_lo, s_lo);
            al_hi = t == 0 ? s_hi : __fadd_rn(al_hi, s_hi);
          } else {
            decode_axis_plain(a_lo, a_hi, t_c, t_s, lo, hi);
          }
          sd[(t * 2 + 0) * kThreads] = lo;
          sd[(t * 2 + 1) * kThreads] = hi;
          sum_lo = t == 0 ? lo : __fadd_rn(sum_lo, lo);
          sum_hi = t == 0 ? hi : __fadd_rn(sum_hi, hi);
        }
      }
      if (k_box_mc) {
        const float fT = (float)T;
        m_lo = __fdiv_rn(sum_lo, fT);
        m_hi = __fdiv_rn(sum_hi, fT);
        float q_lo = 0.f, q_hi = 0.f;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          const float d0 = __fsub_rn(sd[(t * 2 + 0) * kThreads], m_lo);
          const float d1 = __fsub_rn(sd[(t * 2 + 1) * kThreads], m_hi);
          q_lo = t == 0 ? __fmul_rn(d0, d0) : __fadd_rn(q_lo, __fmul_rn(d0, d0));
          q_hi = t == 0 ? __fmul_rn(d1, d1) : __fadd_rn(q_hi, __fmul_rn(d1, d1));
        }
        sd_lo = __fsqrt_rn(__fdiv_rn(q_lo, fT));
        sd_hi = __fsqrt_rn(__fdiv_rn(q_hi, fT));
        al_lo = __fdiv_rn(al_lo, fT);
        al_hi = __fdiv_rn(al_hi, fT);
      } else {
        m_lo = sum_lo;
        m_hi = sum_hi;
      }
    } else if (TMAX != 0) {
      constexpr int TM = TMAX == 0 ? 1 : TMAX;
      // (a) every load of this thread is issued first (2 x T independent 16-byte requests in
      //     flight) and parked in shared memory; (b) a rolled loop decodes sample by sample.
      float* st = smem_stage + threadIdx.x;  // slot (t, k) at st[(t * 4 + k) * kThreads]
      {
        float4 tt[TM], sg[TM];
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (t < T) {
            tt[t] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride));
            if (k_la) sg[t] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride + 4 * A));
          }
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (t < T) {
            st[(t * 4 + 0) * kThreads] = axis ? tt[t].y : tt[t].x;
            st[(t * 4 + 1) * kThreads] = axis ? tt[t].w : tt[t].z;
            if (k_la) {
              st[(t * 4 + 2) * kThreads] = axis ? sg[t].y : sg[t].x;
              st[(t * 4 + 3) * kThreads] = axis ? sg[t].w : sg[t].z;
            }
          }
      }
      float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const float t_c = st[(t * 4 + 0) * kThreads], t_s = st[(t * 4 + 1) * kThreads];
        float lo, hi;
        if (k_la) {
          float s_lo, s_hi;
          decode_axis_la(k_method, smem_tbl, a_lo, a_hi, t_c, t_s, st[(t * 4 + 2) * kThreads],
                         st[(t * 4 + 3) * kThreads], lo, hi, s_lo, s_hi);
          al_lo = t == 0 ? s_lo : __fadd_rn(al_lo, s_lo);
          al_hi = t == 0 ? s_hi : __fadd_rn(al_hi, s_hi);
        } else {
          decode_axis_plain(a_lo, a_hi, t_c, t_s, lo, hi);
        }
        st[(t * 4 + 0) * kThreads] = lo;
        st[(t * 4 + 1) * kThreads] = hi;
        sum_lo = t == 0 ? lo : __fadd_rn(sum_lo, lo);
        sum_hi = t == 0 ? hi : __fadd_rn(sum_hi, hi);
      }
      if (k_box_mc) {
        const float fT = (float)T;
        m_lo = __fdiv_rn(sum_lo, fT);
        m_hi = __fdiv_rn(sum_hi, fT);
        float q_lo = 0.f, q_hi = 0.f;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          const float d0 = __fsub_rn(st[(t * 4 + 0) * kThreads], m_lo);
          const float d1 = __fsub_rn(st[(t * 4 + 1) * kThreads], m_hi);
          q_lo = t == 0 ? __fmul_rn(d0, d0) : __fadd_rn(q_lo, __fmul_rn(d0, d0));
          q_hi = t == 0 ? __fmul_rn(d1, d1) : __fadd_rn(q_hi, __fmul_rn(d1, d1));
        }
        sd_lo = __fsqrt_rn(__fdiv_rn(q_lo, fT));
        sd_hi = __fsqrt_rn(__fdiv_rn(q_hi, fT));
        al_lo = __fdiv_rn(al_lo, fT);
        al_hi = __fdiv_rn(al_hi, fT);
      } else {
        m_lo = sum_lo;
        m_hi = sum_hi;
      }
    } else {
      // many samples: decode twice (sum pass, deviation pass)
      float s_lo_sum = 0.f, s_hi_sum = 0.f, q_lo = 0.f, q_hi = 0.f;
      m_lo = m_hi = 0.f;
      for (int pass = 0; pass < (k_box_mc ? 2 : 1); ++pass) {
        for (int t = 0; t < T; ++t) {
          const float4 tt = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride));
          const float t_c = axis ? tt.y : tt.x, t_s = axis ? tt.w : tt.z;
          float lo, hi, s_lo = 0.f, s_hi = 0.f;
          if (k_la) {
            const float4 sg = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride + 4 * A));
            decode_axis_la(k_method, smem_tbl, a_lo, a_hi, t_c, t_s, axis ? sg.y : sg.x, axis ? sg.w : sg.z, lo, hi, s_lo, s_hi);
          } else {
            decode_axis_plain(a_lo, a_hi, t_c, t_s, lo, hi);
          }
          if (pass == 0) {
            s_lo_sum = t == 0 ? lo : __fadd_rn(s_lo_sum, lo);
            s_hi_sum = t == 0 ? hi : __fadd_rn(s_hi_sum, hi);
            al_lo = t == 0 ? s_lo : __fadd_rn(al_lo, s_lo);
            al_hi = t == 0 ? s_hi : __fadd_rn(al_hi, s_hi);
          } else {
            const float d0 = __fsub_rn(lo, m_lo), d1 = __fsub_rn(hi, m_hi);
            q_lo = t == 0 ? __fmul_rn(d0, d0) : __fadd_rn(q_lo, __fmul_rn(d0, d0));
            q_hi = t == 0 ? __fmul_rn(d1, d1) : __fadd_rn(q_hi, __fmul_rn(d1, d1));
          }
        }
        if (pass == 0) {
          m_lo = k_box_mc ? __fdiv_rn(s_lo_sum, (float)T) : s_lo_sum;
          m_hi = k_box_mc ? __fdiv_rn(s_hi_sum, (float)T) : s_hi_sum;
        }
      }
      if (k_box_mc) {
        sd_lo = __fsqrt_rn(__fdiv_rn(q_lo, (float)T));
        sd_hi = __fsqrt_rn(__fdiv_rn(q_hi, (float)T));
        al_lo = __fdiv_rn(al_lo, (float)T);
        al_hi = __fdiv_rn(al_hi, (float)T);
      }
    }
    // pair exchange: the y lane assembles (ymin, xmin, ymax, xmax) rows
    const float o_m_lo = __shfl_xor_sync(0xffffffffu, m_lo, 1), o_m_hi = __shfl_xor_sync(0xffffffffu, m_hi, 1);
    const float o_sd_lo = __shfl_xor_sync(0xffffffffu, sd_lo, 1), o_sd_hi = __shfl_xor_sync(0xffffffffu, sd_hi, 1);
    const float o_al_lo = __shfl_xor_sync(0xffffffffu, al_lo, 1), o_al_hi = __shfl_xor_sync(0xffffffffu, al_hi, 1);
    if (!live) continue;
    const size_t o = (size_t)b * p.N + n;
    if (axis == 0) {
      if (p.out.boxes) reinterpret_cast<float4*>(p.out.boxes)[o] = make_float4(m_lo, o_m_lo, m_hi, o_m_hi);
      if (p.out.mcbox && k_box_mc)
        reinterpret_cast<float4*>(p.out.mcbox)[o] = make_float4(sd_lo, o_sd_lo, sd_hi, o_sd_hi);
    } else {
      if (p.out.albox && k_la)
        reinterpret_cast<float4*>(p.out.albox)[o] = make_float4(o_al_lo, al_lo, o_al_hi, al_hi);
      if (p.out.scores || p.out.classes) {
        const float* ml = smem_mean + (size_t)ii * C;
        float best = ml[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
          const float v = ml[c];
          if (v > best) {
            best = v;
            arg = c;
          }
        }
        if (p.out.scores) p.out.scores[o] = sigmoid_ref(best);
        if (p.out.classes) p.out.classes[o] = arg;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// fp32 closed-form variant (udal_config.decode_precision = UDAL_DECODE_FP32): the same tile structure, arithmetic of the fused
// predict + decode kernels (heads_dw.cu): ex2.approx, series for exp(v) - 1 (the cancellation of utils_box.py:151-152 at small
// variances), one-pass statistics around the first sample (mean = sequential sum / T exactly as the fp64 kernel; variance =
// (sum d^2 - (sum d)^2 / T) / T with d = x - x_0, so the subtraction cancels against the spread, not the magnitude).  No
// staging in shared memory: TCH samples (2 TCH independent 16-byte loads per thread) are in flight per thread, three CTAs
// per SM - the kernel moves its algorithmic bytes once and is bound by HBM.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float f32_expm1(float v) {  // v = sigma^2 >= 0
  if (v < 0.25f) {
    float q = fmaf(v, 1.f / 720.f, 1.f / 120.f);
    q = fmaf(q, v, 1.f / 24.f);
    q = fmaf(q, v, 1.f / 6.f);
    q = fmaf(q, v, 0.5f);
    q = fmaf(q, v, 1.f);
    return q * v;
  }
  return f32_exp(v) - 1.f;
}
__device__ __forceinline__ float f32_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// one axis of one sample.  sa / ca / sa2: anchor size, centre, size^2
__device__ __forceinline__ void f32_decode_axis(int method, int la, float sa, float ca, float sa2, float t_c, float t_s, float s_c,
                                                float s_s, float& lo, float& hi, float& sd_lo, float& sd_hi) {
  const float c = fmaf(t_c, sa, ca);
  if (!la) {  // anchors.py:41-75
    const float half = 0.5f * f32_exp(t_s) * sa;
    lo = c - half;
    hi = c + half;
    sd_lo = sd_hi = 0.f;
    return;
  }
  const float vs = s_s * s_s, vc = s_c * s_c;
  if (method == UDAL_DECODE_FALSEDEC) {  // utils_box.py:186-266: the variances decoded like offsets
    const float half = 0.5f * f32_exp(t_s) * sa;
    lo = c - half;
    hi = c + half;
    const float dhalf = 0.5f * f32_exp(vs) * sa;
    const float dc = fmaf(vc, sa, ca);
    sd_lo = f32_sqrt(fabsf(dc - dhalf));
    sd_hi = f32_sqrt(dc + dhalf);
    return;
  }
  // l-norm (utils_box.py:125-160) and n-flow (same closed forms through tfp)
  const float e = f32_exp(fmaf(0.5f, vs, t_s));
  const float half = 0.5f * e * sa;
  lo = c - half;
  hi = c + half;
  // Var(centre) + Var(size) / 4, Var(size) = (exp(v) - 1) exp(2 t + v) sa^2
  sd_lo = sd_hi = f32_sqrt(sa2 * fmaf(0.25f * f32_expm1(vs), e * e, vc));
}

struct F32Stat {  // running statistics of one quantity over the samples
  float sum, x0, s1, s2;
  __device__ __forceinline__ void add(float x, int t) {
    if (t == 0) {
      sum = x;
      x0 = x;
      s1 = 0.f;
      s2 = 0.f;
    } else {
      sum += x;
      const float d = x - x0;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  __device__ __forceinline__ float mean(float t) const { return __fdiv_rn(sum, t); }  // the fp64 kernel's mean, bit for bit
  __device__ __forceinline__ float sd(float inv_t) const { return f32_sqrt(fmaxf(fmaf(-s1 * inv_t, s1, s2), 0.f) * inv_t); }
};

template <int TCH>
__global__ void __launch_bounds__(kThreads, 3) decode_moments_f32_kernel(const DecodeParams p) {
  extern __shared__ float smem_mean[];  // [tile anchors * C]
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = find_level(p.tile_off, p.geom.num_levels, tile);
  const int hw = p.geom.h[l] * p.geom.w[l];
  const int p0 = (tile - p.tile_off[l]) * p.tile_px;
  const int npx = min(p.tile_px, hw - p0);
  const int A = p.A, C = p.C;
  const int64_t anchor0 = (int64_t)A * (p.geom.pix_off[l] + p0);

  // ---- phase 1: class logits, four consecutive logits per thread where the tile is 16-byte aligned ----
  {
    const int T = p.Tc;
    const float fT = (float)T, inv_t = 1.f / fT;
    const int count = npx * A * C;
    const size_t plane = (size_t)hw * A * C;
    const float* base = p.cls.p[l] + ((size_t)b * hw + p0) * A * C;
    const size_t t_stride = (size_t)p.batch * plane;
    float* mo = p.out.mean_logits ? p.out.mean_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
    float* so = (p.out.std_logits && p.cls_mc) ? p.out.std_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
    const bool vec = (((uintptr_t)base | (uintptr_t)(t_stride * 4) | (uintptr_t)mo | (uintptr_t)so) & 15) == 0;
    const int nvec = vec ? count >> 2 : 0;
    for (int e4 = threadIdx.x; e4 < nvec; e4 += kThreads) {
      F32Stat st[4];
      const float4* src = reinterpret_cast<const float4*>(base) + e4;
#pragma unroll 1
      for (int t0 = 0; t0 < T; t0 += 2 * TCH) {
        float4 v[2 * TCH];
#pragma unroll
        for (int u = 0; u < 2 * TCH; ++u)
          if (t0 + u < T) v[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)(t0 + u) * t_stride));
#pragma unroll
        for (int u = 0; u < 2 * TCH; ++u)
          if (t0 + u < T) {
            st[0].add(v[u].x, t0 + u);
            st[1].add(v[u].y, t0 + u);
            st[2].add(v[u].z, t0 + u);
            st[3].add(v[u].w, t0 + u);
          }
      }
      const float4 m = make_float4(st[0].mean(fT), st[1].mean(fT), st[2].mean(fT), st[3].mean(fT));
      if (mo) reinterpret_cast<float4*>(mo)[e4] = m;
      if (so) reinterpret_cast<float4*>(so)[e4] = make_float4(st[0].sd(inv_t), st[1].sd(inv_t), st[2].sd(inv_t), st[3].sd(inv_t));
      reinterpret_cast<float4*>(smem_mean)[e4] = m;
    }
    for (int e = nvec * 4 + threadIdx.x; e < count; e += kThreads) {  // unaligned tiles / tail
      F32Stat st;
      for (int t = 0; t < T; ++t) st.add(__ldg(base + (size_t)t * t_stride + e), t);
      const float m = st.mean(fT);
      if (mo) mo[e] = m;
      if (so) so[e] = st.sd(inv_t);
      smem_mean[e] = m;
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per (anchor, axis) ----
  const int axis = threadIdx.x & 1;
  const int T = p.Tb;
  const float fT = (float)T, inv_t = 1.f / fT;
  const int method = p.method, la = p.la;
  for (int i0 = 0; i0 < npx * A; i0 += kThreads / 2) {
    const int i = i0 + (threadIdx.x >> 1);
    const bool live = i < npx * A;
    const int ii = live ? i : 0;
    const int px = p0 + ii / A;
    const int a = ii - (ii / A) * A;
    const int64_t n = anchor0 + ii;
    const float4 anc = __ldg(reinterpret_cast<const float4*>(p.anchors) + n);
    const float a_lo = axis ? anc.y : anc.x, a_hi = axis ? anc.w : anc.z;
    const float sa = a_hi - a_lo, ca = 0.5f * (a_lo + a_hi), sa2 = sa * sa;
    const float* bb = p.box.p[l] + ((size_t)b * hw + px) * p.BC + a * 4;
    const size_t t_stride = (size_t)p.batch * hw * p.BC;
    F32Stat s_lo, s_hi;
    float al_lo = 0.f, al_hi = 0.f;
#pragma unroll 1
    for (int t0 = 0; t0 < T; t0 += TCH) {
      float4 tt[TCH], sg[TCH];
#pragma unroll
      for (int u = 0; u < TCH; ++u)
        if (t0 + u < T) {
          tt[u] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)(t0 + u) * t_stride));
          if (la) sg[u] = __ldg(reinterpret_cast<const float4*>(bb + (size_t)(t0 + u) * t_stride + 4 * A));
        }
#pragma unroll
      for (int u = 0; u < TCH; ++u)
        if (t0 + u < T) {
          float lo, hi, d_lo, d_hi;
          f32_decode_axis(method, la, sa, ca, sa2, axis ? tt[u].y : tt[u].x, axis ? tt[u].w : tt[u].z,
                          la ? (axis ? sg[u].y : sg[u].x) : 0.f, la ? (axis ? sg[u].w : sg[u].z) : 0.f, lo, hi, d_lo, d_hi);
          s_lo.add(lo, t0 + u);
          s_hi.add(hi, t0 + u);
          al_lo += d_lo;
          al_hi += d_hi;
        }
    }
    const float m_lo = s_lo.mean(fT), m_hi = s_hi.mean(fT);
    const float sd_lo = s_lo.sd(inv_t), sd_hi = s_hi.sd(inv_t);
    al_lo *= inv_t;
    al_hi *= inv_t;
    const float o_m_lo = __shfl_xor_sync(0xffffffffu, m_lo, 1), o_m_hi = __shfl_xor_sync(0xffffffffu, m_hi, 1);
    const float o_sd_lo = __shfl_xor_sync(0xffffffffu, sd_lo, 1), o_sd_hi = __shfl_xor_sync(0xffffffffu, sd_hi, 1);
    const float o_al_lo = __shfl_xor_sync(0xffffffffu, al_lo, 1), o_al_hi = __shfl_xor_sync(0xffffffffu, al_hi, 1);
    if (!live) continue;
    const size_t o = (size_t)b * p.N + n;
    if (axis == 0) {
      if (p.out.boxes) reinterpret_cast<float4*>(p.out.boxes)[o] = make_float4(m_lo, o_m_lo, m_hi, o_m_hi);
      if (p.out.mcbox && p.box_mc) reinterpret_cast<float4*>(p.out.mcbox)[o] = make_float4(sd_lo, o_sd_lo, sd_hi, o_sd_hi);
    } else {
      if (p.out.albox && la) reinterpret_cast<float4*>(p.out.albox)[o] = make_float4(o_al_lo, al_lo, o_al_hi, al_hi);
      if (p.out.scores || p.out.classes) {
        const float* ml = smem_mean + (size_t)ii * C;
        float best = ml[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
          const float v = ml[c];
          if (v > best) {
            best = v;
            arg = c;
          }
        }
        if (p.out.scores) p.out.scores[o] = __frcp_rn(1.f + f32_exp(-best));
        if (p.out.classes) p.out.classes[o] = arg;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// Persistent, TMA-staged form of the fp32 kernel (the one UDAL_DECODE_FP32 launches whenever the configuration allows):
// CTAs sized to the SM count loop over (tile of 16 pixels, image) items; a producer warp streams the item's samples through
// a ring of shared-memory stages with bulk copies (cp.async.bulk: one contiguous run of logits + one of box outputs per
// sample, completion on an mbarrier); the consumers fold each stage into running statistics held in registers and release
// it.  The ring keeps stages x ~10 KB x CTAs per SM in flight whatever the consumers do, and it runs ahead across items: the
// loads of the next tiles land while this one is finalised.  Consumers: thread = anchor, BOTH axes as packed fp32 pairs
// ((ty,tx), (th,tw), (sy,sx), (sh,sw) are adjacent in the reference's layout: every decode step is one f32x2 instruction),
// logits as EPT2 strided pairs - half the instructions per byte of the per-tile kernel, which is what bounded it.
// Tiles whose runs are not 16-byte aligned (odd pixel counts of the smallest levels) are copied by the producer warp itself.
// -------------------------------------------------------------------------------------------
constexpr int kStreamPx = 16, kStreamMaxStages = 8;

__device__ __forceinline__ uint32_t dm_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dm_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void dm_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void dm_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dm_bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// packed fp32 pairs: two independent IEEE operations per instruction
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_sub(f2_t a, f2_t b) {
  f2_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t f2_bcast(float v) { return f2_pack(v, v); }

struct F32Stat2 {  // F32Stat of two quantities at once
  f2_t sum, x0, s1, s2;
  template <bool FIRST>
  __device__ __forceinline__ void add(f2_t x) {
    if (FIRST) {
      sum = x;
      x0 = x;
      s1 = 0ull;
      s2 = 0ull;
    } else {
      sum = f2_add(sum, x);
      const f2_t d = f2_sub(x, x0);
      s1 = f2_add(s1, d);
      s2 = f2_fma(d, d, s2);
    }
  }
  __device__ __forceinline__ void mean(float t, float& a, float& b) const {
    float sa, sb;
    f2_unpack(sum, sa, sb);
    a = t == 1.f ? sa : __fdiv_rn(sa, t);
    b = t == 1.f ? sb : __fdiv_rn(sb, t);
  }
  __device__ __forceinline__ void sd(float inv_t, float& a, float& b) const {
    float p1, q1, p2, q2;
    f2_unpack(s1, p1, q1);
    f2_unpack(s2, p2, q2);
    a = f32_sqrt(fmaxf(fmaf(-p1 * inv_t, p1, p2), 0.f) * inv_t);
    b = f32_sqrt(fmaxf(fmaf(-q1 * inv_t, q1, q2), 0.f) * inv_t);
  }
};

// the same interface for a single sample (no MC axis on either head): the value itself, no statistics - 40 registers less
struct F32One2 {
  f2_t sum;
  template <bool FIRST>
  __device__ __forceinline__ void add(f2_t x) { sum = x; }
  __device__ __forceinline__ void mean(float, float& a, float& b) const { f2_unpack(sum, a, b); }
  __device__ __forceinline__ void sd(float, float& a, float& b) const { a = b = 0.f; }
};

// exp(v) - 1 for both halves: degree-8 series below 1 (relative truncation < 3e-6), ex2 above
__device__ __forceinline__ f2_t f2_expm1(f2_t v) {
  f2_t q = f2_fma(v, f2_bcast(1.f / 362880.f), f2_bcast(1.f / 40320.f));
  q = f2_fma(q, v, f2_bcast(1.f / 5040.f));
  q = f2_fma(q, v, f2_bcast(1.f / 720.f));
  q = f2_fma(q, v, f2_bcast(1.f / 120.f));
  q = f2_fma(q, v, f2_bcast(1.f / 24.f));
  q = f2_fma(q, v, f2_bcast(1.f / 6.f));
  q = f2_fma(q, v, f2_bcast(0.5f));
  q = f2_fma(q, v, f2_bcast(1.f));
  q = f2_mul(q, v);
  float v0, v1;
  f2_unpack(v, v0, v1);
  if (v0 >= 1.f || v1 >= 1.f) {  // sigma >= 1: rare
    float q0, q1;
    f2_unpack(q, q0, q1);
    if (v0 >= 1.f) q0 = f32_exp(v0) - 1.f;
    if (v1 >= 1.f) q1 = f32_exp(v1) - 1.f;
    q = f2_pack(q0, q1);
  }
  return q;
}

struct StreamParams {
  DecodeParams d;       // tile_px = kStreamPx, tile_off for it
  int items, tiles;     // (tile, image) items; tiles per image
  int cls_floats;       // floats reserved for the logits run of a stage (16 A C, rounded up to 4)
  int stage_floats;     // floats per stage
  int stages;           // ring depth (<= kStreamMaxStages)
  int consumers;        // consumer threads: 16 A rounded up to whole warps
};

// consumer state of one item
template <int EPT2, bool SINGLE>
struct StreamState {
  typedef typename std::conditional<SINGLE, F32One2, F32Stat2>::type Stat;
  Stat st[EPT2];      // logit pairs q = tid + j * consumers
  Stat lo, hi;        // (ymin, xmin), (ymax, xmax)
  f2_t al;            // sum of the aleatoric std (y, x)
};

template <int EPT2, bool FIRST, bool SINGLE>
__device__ __forceinline__ void stream_consume(StreamState<EPT2, SINGLE>& S, const float* __restrict__ sc, const float* __restrict__ sb,
                                               int tid, int nct, bool do_cls, bool do_box, int la, int sig_off, f2_t sa, f2_t hsa,
                                               f2_t ca, f2_t saq) {
  if (do_cls) {
#pragma unroll
    for (int j = 0; j < EPT2; ++j)
      S.st[j].template add<FIRST>(*reinterpret_cast<const f2_t*>(sc + 2 * (tid + j * nct)));
  }
  if (do_box) {
    const ulonglong2 tt = *reinterpret_cast<const ulonglong2*>(sb);  // (ty, tx), (th, tw)
    const f2_t c = f2_fma(tt.x, sa, ca);
    f2_t arg = tt.y, vs = 0ull, vc = 0ull;
    if (la) {
      const ulonglong2 sg = *reinterpret_cast<const ulonglong2*>(sb + sig_off);  // (sy, sx), (sh, sw)
      vc = f2_mul(sg.x, sg.x);
      vs = f2_mul(sg.y, sg.y);
      arg = f2_fma(f2_bcast(0.5f), vs, tt.y);
    }
    arg = f2_mul(arg, f2_bcast(1.4426950408889634f));
    float a0, a1;
    f2_unpack(arg, a0, a1);
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
    const f2_t e = f2_pack(a0, a1);
    const f2_t half = f2_mul(e, hsa);
    S.lo.template add<FIRST>(f2_sub(c, half));
    S.hi.template add<FIRST>(f2_add(c, half));
    if (la) {
      // Var(centre) + Var(size) / 4, Var(size) = (exp(v) - 1) exp(2 t + v) sa^2
      const f2_t var = f2_mul(saq, f2_fma(f2_mul(f2_expm1(vs), f2_bcast(0.25f)), f2_mul(e, e), vc));
      float v0, v1;
      f2_unpack(var, v0, v1);
      const f2_t sd = f2_pack(f32_sqrt(v0), f32_sqrt(v1));
      S.al = FIRST ? sd : f2_add(S.al, sd);
    }
  }
}

// SINGLE: one sample on both heads (no MC dropout): the state is the sample itself, 64 registers, five CTAs per SM - the
// one-stage items of this case are bound by the per-item latency chain (wait, fold, barrier, arg-max, stores), not by HBM
template <int EPT2, bool SINGLE>
__global__ void __maxnreg__(SINGLE ? 64 : (EPT2 <= 5 ? 112 : 168)) decode_stream_kernel(const StreamParams sp) {
  const DecodeParams& p = sp.d;
  extern __shared__ __align__(128) float smem_f[];
  const int A = p.A, C = p.C, BC = p.BC;
  const int NCT = sp.consumers;
  float* stages = smem_f;
  float* smean = smem_f + sp.stages * sp.stage_floats;  // [2][16 A C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smean + 2 * sp.cls_floats);
  const uint32_t full0 = dm_s32(bars), empty0 = dm_s32(bars + kStreamMaxStages);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < sp.stages; ++s) {
      dm_bar_init(full0 + 8 * s, 1);
      dm_bar_init(empty0 + 8 * s, NCT >> 5);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int Tc = p.Tc, Tb = p.Tb, T = Tc > Tb ? Tc : Tb;

  if (tid >= NCT) {
    // ---- producer warp ----
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < sp.items; item += gridDim.x) {
      const int b = item / sp.tiles, tile = item - b * sp.tiles;
      const int l = find_level(p.tile_off, p.geom.num_levels, tile);
      const int hw = p.geom.h[l] * p.geom.w[l];
      const int p0 = (tile - p.tile_off[l]) * kStreamPx;
      const int npx = min(kStreamPx, hw - p0);
      const size_t c_stride = (size_t)p.batch * hw * A * C, b_stride = (size_t)p.batch * hw * BC;
      const float* csrc = p.cls.p[l] + ((size_t)b * hw + p0) * A * C;
      const float* bsrc = p.box.p[l] + ((size_t)b * hw + p0) * BC;
      const int cn = npx * A * C, bn = npx * BC;
      const bool bulk = ((((uintptr_t)csrc | (uintptr_t)bsrc) & 15) == 0) && (((c_stride | b_stride | (size_t)cn | (size_t)bn) & 3) == 0);
      for (int t = 0; t < T; ++t) {
        dm_bar_wait(empty0 + 8 * s, ph ^ 1u);
        float* dst = stages + (size_t)s * sp.stage_floats;
        const bool wc = t < Tc, wb = t < Tb;
        if (bulk) {
          if (lane == 0) {
            dm_bar_expect_tx(full0 + 8 * s, (wc ? cn * 4 : 0) + (wb ? bn * 4 : 0));
            if (wc) dm_bulk_g2s(dm_s32(dst), csrc + (size_t)t * c_stride, cn * 4, full0 + 8 * s);
            if (wb) dm_bulk_g2s(dm_s32(dst + sp.cls_floats), bsrc + (size_t)t * b_stride, bn * 4, full0 + 8 * s);
          }
        } else {
          if (wc)
            for (int e = lane; e < cn; e += 32) dst[e] = __ldg(csrc + (size_t)t * c_stride + e);
          if (wb)
            for (int e = lane; e < bn; e += 32) dst[sp.cls_floats + e] = __ldg(bsrc + (size_t)t * b_stride + e);
          __syncwarp();
          if (lane == 0) dm_bar_arrive(full0 + 8 * s);
        }
        __syncwarp();
        if (++s == sp.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  // ---- consumers: thread = anchor of the tile (both axes), plus EPT2 strided logit pairs ----
  const int la = p.la;
  const float fTc = (float)Tc, inv_tc = 1.f / fTc, fTb = (float)Tb, inv_tb = 1.f / fTb;
  int s = 0, it = 0;
  uint32_t ph = 0;
  // geometry + anchor of this thread for an item; the next item's are fetched one item ahead (the anchor table is the only
  // global load of the consumers: it must not sit on the critical path of a one-stage item)
  struct Geo {
    int b, npx, count, ii;
    bool live;
    int64_t anchor0;
    float4 anc;
  };
  auto geo_of = [&](int item) {
    Geo g;
    g.b = item / sp.tiles;
    const int tile = item - g.b * sp.tiles;
    const int l = find_level(p.tile_off, p.geom.num_levels, tile);
    const int hw = p.geom.h[l] * p.geom.w[l];
    const int p0 = (tile - p.tile_off[l]) * kStreamPx;
    g.npx = min(kStreamPx, hw - p0);
    g.count = g.npx * A * C;
    g.anchor0 = (int64_t)A * (p.geom.pix_off[l] + p0);
    g.live = tid < g.npx * A;
    g.ii = g.live ? tid : 0;
    g.anc = __ldg(reinterpret_cast<const float4*>(p.anchors) + g.anchor0 + g.ii);
    return g;
  };
  Geo nxt = geo_of(blockIdx.x);
  for (int item = blockIdx.x; item < sp.items; item += gridDim.x, ++it) {
    const Geo g = nxt;
    if (item + (int)gridDim.x < sp.items) nxt = geo_of(item + gridDim.x);
    const int b = g.b, count = g.count, ii = g.ii;
    const bool live = g.live;
    const int64_t anchor0 = g.anchor0, n = anchor0 + ii;
    const float4 anc = g.anc;
    const int box_off = sp.cls_floats + (ii / A) * BC + (ii - (ii / A) * A) * 4;
    const float say = anc.z - anc.x, sax = anc.w - anc.y;
    const f2_t sa = f2_pack(say, sax), hsa = f2_pack(0.5f * say, 0.5f * sax), saq = f2_pack(say * say, sax * sax);
    const f2_t ca = f2_pack(0.5f * (anc.x + anc.z), 0.5f * (anc.y + anc.w));
    StreamState<EPT2, SINGLE> S;
    S.al = 0ull;
    for (int t = 0; t < T; ++t) {
      dm_bar_wait(full0 + 8 * s, ph);
      const float* sc = stages + (size_t)s * sp.stage_floats;
      // (a sample index beyond Tc / Tb exists only on the side that has the longer MC axis; t == 0 is valid for both)
      if (SINGLE || t == 0) stream_consume<EPT2, true, SINGLE>(S, sc, sc + box_off, tid, NCT, true, true, la, 4 * A, sa, hsa, ca, saq);
      else stream_consume<EPT2, false, SINGLE>(S, sc, sc + box_off, tid, NCT, t < Tc, t < Tb, la, 4 * A, sa, hsa, ca, saq);
      __syncwarp();
      if (lane == 0) dm_bar_arrive(empty0 + 8 * s);
      if (++s == sp.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    // ---- finalize the item ----
    float* sm = smean + (it & 1) * sp.cls_floats;
    {
      float* mo = p.out.mean_logits ? p.out.mean_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
      float* so = (p.out.std_logits && p.cls_mc) ? p.out.std_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
      const bool vec = ((((uintptr_t)mo | (uintptr_t)so) & 7) == 0) && ((count & 1) == 0);
#pragma unroll
      for (int j = 0; j < EPT2; ++j) {
        const int e = 2 * (tid + j * NCT);
        if (e < count) {
          float m0, m1, d0 = 0.f, d1 = 0.f;
          S.st[j].mean(fTc, m0, m1);
          if (so) S.st[j].sd(inv_tc, d0, d1);
          *reinterpret_cast<float2*>(sm + e) = make_float2(m0, m1);
          if (vec) {
            if (mo) *reinterpret_cast<float2*>(mo + e) = make_float2(m0, m1);
            if (so) *reinterpret_cast<float2*>(so + e) = make_float2(d0, d1);
          } else {
            if (mo) mo[e] = m0;
            if (so) so[e] = d0;
            if (e + 1 < count) {
              if (mo) mo[e + 1] = m1;
              if (so) so[e + 1] = d1;
            }
          }
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NCT) : "memory");  // the item's mean logits are in shared memory (consumers only)
    if (!live) continue;
    const size_t o = (size_t)b * p.N + n;
    float ymin, xmin, ymax, xmax;
    S.lo.mean(fTb, ymin, xmin);
    S.hi.mean(fTb, ymax, xmax);
    if (p.out.boxes) reinterpret_cast<float4*>(p.out.boxes)[o] = make_float4(ymin, xmin, ymax, xmax);
    if (p.out.mcbox && p.box_mc) {
      float a0, a1, b0, b1;
      S.lo.sd(inv_tb, a0, a1);
      S.hi.sd(inv_tb, b0, b1);
      reinterpret_cast<float4*>(p.out.mcbox)[o] = make_float4(a0, a1, b0, b1);
    }
    if (p.out.albox && la) {
      float a0, a1;
      f2_unpack(S.al, a0, a1);
      a0 *= inv_tb;
      a1 *= inv_tb;
      reinterpret_cast<float4*>(p.out.albox)[o] = make_float4(a0, a1, a0, a1);
    }
    if (p.out.scores || p.out.classes) {
      const float* ml = sm + (size_t)ii * C;
      float best = ml[0];
      int arg = 0;
      for (int c = 1; c < C; ++c) {
        const float v = ml[c];
        if (v > best) {
          best = v;
          arg = c;
        }
      }
      if (p.out.scores) p.out.scores[o] = __frcp_rn(1.f + f32_exp(-best));
      if (p.out.classes) p.out.classes[o] = arg;
    }
  }
}

template <int EPT2, bool SINGLE>
int launch_stream(udal_ctx* ctx, StreamParams& sp) {
  const int threads = sp.consumers + 32;
  int dev = 0, sms = 0;
  UDAL_CUDA(cudaGetDevice(&dev));
  UDAL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // ring depth: as deep as lets the CTAs the register file admits share the SM's shared memory
  cudaFuncAttributes fa;
  UDAL_CUDA(cudaFuncGetAttributes(&fa, decode_stream_kernel<EPT2, SINGLE>));
  int by_regs = 65536 / (fa.numRegs * threads > 0 ? fa.numRegs * ((threads + 31) / 32 * 32) : 1);
  // measured (configs[3], B = 64): two CTAs per SM with a ring of 8 stages are 2 % faster than three with 6 at T >= 10
  // (0.88 / 0.98 / 0.99 of the HBM peak), three are 6 % faster at T = 1 (one-stage items: the per-item statistics dominate)
  const int T = sp.d.Tc > sp.d.Tb ? sp.d.Tc : sp.d.Tb;
  by_regs = std::max(1, std::min(by_regs, SINGLE ? 5 : (T <= 4 ? 3 : 2)));
  const size_t fixed = (size_t)2 * sp.cls_floats * 4 + 2 * kStreamMaxStages * 8;
  const size_t per_cta = (size_t)(227 * 1024) / by_regs - 1024;
  UDAL_REQUIRE(per_cta > fixed + 2 * (size_t)sp.stage_floats * 4, "decode_stream_kernel: stage of %d floats does not fit", sp.stage_floats);
  sp.stages = (int)std::min<size_t>(kStreamMaxStages, (per_cta - fixed) / ((size_t)sp.stage_floats * 4));
  const size_t smem = (size_t)sp.stages * sp.stage_floats * 4 + fixed;
  UDAL_CUDA(cudaFuncSetAttribute(decode_stream_kernel<EPT2, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (without the carve-out hint the driver sizes shared memory for two CTAs of this size, not the three the registers admit)
  UDAL_CUDA(cudaFuncSetAttribute(decode_stream_kernel<EPT2, SINGLE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int nb = 0;
  UDAL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_stream_kernel<EPT2, SINGLE>, threads, smem));
  UDAL_REQUIRE(nb >= 1, "decode_stream_kernel does not fit an SM (%zu bytes of shared memory)", smem);
  const int grid = std::min(sp.items, sms * nb);
  decode_stream_kernel<EPT2, SINGLE><<<grid, threads, smem, ctx->stream>>>(sp);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// logits only (eval variant): mean / std of every (anchor, class) logit
template <int TMAX>
__global__ void __launch_bounds__(kThreads) logit_moments_kernel(const DecodeParams p) {
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
  const int l = find_level(p.tile_off, p.geom.num_levels, tile);
  const int hw = p.geom.h[l] * p.geom.w[l];
  const int p0 = (tile - p.tile_off[l]) * p.tile_px;
  const int npx = min(p.tile_px, hw - p0);
  const int A = p.A, C = p.C;
  const int64_t anchor0 = (int64_t)A * (p.geom.pix_off[l] + p0);
  const int count = npx * A * C;
  const size_t plane = (size_t)hw * A * C;
  const float* base = p.cls.p[l] + ((size_t)b * hw + p0) * A * C;
  float* mo = p.out.mean_logits ? p.out.mean_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
  float* so = (p.out.std_logits && p.cls_mc) ? p.out.std_logits + ((size_t)b * p.N + anchor0) * C : nullptr;
  logits_run<TMAX>(base, (size_t)p.batch * plane, p.Tc, count, mo, so, nullptr, threadIdx.x, kThreads);
}

struct GatherParams {
  udal_level_ptrs box;
  udal_level_geom geom;
  int batch, A, C, BC, Tb, box_mc, la, method, k;
  const float* anchors;
  int64_t N;
  const int32_t* topk_idx;
  const float* topk_val;
  const float* std_logits;  // [B,N,C] or null
  udal_prenms_topk_out out;
};

// decode + moments on the k gathered (anchor, class) rows of each image.  F32: decode_precision fp32 (closed form in fp32);
// otherwise the per-axis fp64 decode of the dense kernel (exp / sqrt through fast_math64.cuh: the same values, a third of
// the instructions of the libm calls this kernel used to make - it was fp64-instruction bound: 189 us for 64 x 5000 rows, T = 10)
template <bool F32>
__global__ void __launch_bounds__(kThreads) decode_gather_kernel(const GatherParams p) {
  __shared__ double smem_tbl[64];
  if (threadIdx.x < 64) smem_tbl[threadIdx.x] = kExp2Table[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.y;
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= p.k) return;
  const size_t o = (size_t)b * p.k + j;
  const int32_t flat = p.topk_idx[o];
  const int n = flat / p.C;
  const int c = flat - n * p.C;
  const int gp = n / p.A;
  const int a = n - gp * p.A;
  const int l = find_level(p.geom.pix_off, p.geom.num_levels, gp);
  const int px = gp - p.geom.pix_off[l];
  const int hw = p.geom.h[l] * p.geom.w[l];
  const float4 anc = __ldg(reinterpret_cast<const float4*>(p.anchors) + n);
  const float* bb = p.box.p[l] + ((size_t)b * hw + px) * p.BC + a * 4;
  const size_t t_stride = (size_t)p.batch * hw * p.BC;
  const int T = p.Tb;
  float sum[4] = {0, 0, 0, 0}, al[4] = {0, 0, 0, 0}, mb[4], ss[4] = {0, 0, 0, 0};
  for (int pass = 0; pass < (p.box_mc ? 2 : 1); ++pass) {
    for (int t = 0; t < T; ++t) {
      const float4 tt = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride));
      float d[4];
      float4 sg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.la) sg = __ldg(reinterpret_cast<const float4*>(bb + (size_t)t * t_stride + 4 * p.A));
      float sd[4] = {0.f, 0.f, 0.f, 0.f};
      if (F32) {
        const float say = anc.z - anc.x, sax = anc.w - anc.y;
        f32_decode_axis(p.method, p.la, say, 0.5f * (anc.x + anc.z), say * say, tt.x, tt.z, sg.x, sg.z, d[0], d[2], sd[0], sd[2]);
        f32_decode_axis(p.method, p.la, sax, 0.5f * (anc.y + anc.w), sax * sax, tt.y, tt.w, sg.y, sg.w, d[1], d[3], sd[1], sd[3]);
      } else if (p.la) {
        decode_axis_la(p.method, smem_tbl, anc.x, anc.z, tt.x, tt.z, sg.x, sg.z, d[0], d[2], sd[0], sd[2]);
        decode_axis_la(p.method, smem_tbl, anc.y, anc.w, tt.y, tt.w, sg.y, sg.w, d[1], d[3], sd[1], sd[3]);
      } else {
        decode_axis_plain(anc.x, anc.z, tt.x, tt.z, d[0], d[2]);
        decode_axis_plain(anc.y, anc.w, tt.y, tt.w, d[1], d[3]);
      }
      if (p.la && pass == 0)
        for (int q = 0; q < 4; ++q) al[q] = t == 0 ? sd[q] : __fadd_rn(al[q], sd[q]);
      for (int q = 0; q < 4; ++q) {
        if (pass == 0) {
          sum[q] = t == 0 ? d[q] : __fadd_rn(sum[q], d[q]);
        } else {
          const float dv = __fsub_rn(d[q], mb[q]);
          ss[q] = t == 0 ? __fmul_rn(dv, dv) : __fadd_rn(ss[q], __fmul_rn(dv, dv));
        }
      }
    }
    if (pass == 0)
      for (int q = 0; q < 4; ++q) mb[q] = p.box_mc ? __fdiv_rn(sum[q], (float)T) : sum[q];
  }
  if (p.out.boxes) reinterpret_cast<float4*>(p.out.boxes)[o] = make_float4(mb[0], mb[1], mb[2], mb[3]);
  if (p.out.albox && p.la) {
    float4 v;
    v.x = p.box_mc ? __fdiv_rn(al[0], (float)T) : al[0];
    v.y = p.box_mc ? __fdiv_rn(al[1], (float)T) : al[1];
    v.z = p.box_mc ? __fdiv_rn(al[2], (float)T) : al[2];
    v.w = p.box_mc ? __fdiv_rn(al[3], (float)T) : al[3];
    reinterpret_cast<float4*>(p.out.albox)[o] = v;
  }
  if (p.out.mcbox && p.box_mc) {
    float4 v;
    v.x = __fsqrt_rn(__fdiv_rn(ss[0], (float)T));
    v.y = __fsqrt_rn(__fdiv_rn(ss[1], (float)T));
    v.z = __fsqrt_rn(__fdiv_rn(ss[2], (float)T));
    v.w = __fsqrt_rn(__fdiv_rn(ss[3], (float)T));
    reinterpret_cast<float4*>(p.out.mcbox)[o] = v;
  }
  if (p.out.scores) p.out.scores[o] = F32 ? __frcp_rn(1.f + f32_exp(-p.topk_val[o])) : sigmoid_ref(p.topk_val[o]);
  if (p.out.classes) p.out.classes[o] = c;
  if (p.out.mcclass && p.std_logits) p.out.mcclass[o] = p.std_logits[(size_t)b * p.N * p.C + flat];
}

int fill_params(udal_ctx* ctx, const float* const* cls, const float* const* box, int batch,
                DecodeParams& p, int* total_tiles, size_t* smem) {
  const udal_config& c = ctx->cfg;
  UDAL_REQUIRE(ctx->anchors_set, "anchors not set (call udal_set_anchors)");
  UDAL_REQUIRE(batch > 0, "batch must be positive");
  p.geom = udal_geom(ctx);
  p.batch = batch;
  p.A = c.anchors_per_loc;
  p.C = c.num_classes;
  p.BC = udal_box_channels(ctx);
  p.cls_mc = c.cls_mc;
  p.box_mc = c.box_mc;
  p.Tc = c.cls_mc ? c.mc_samples : 1;
  p.Tb = c.box_mc ? c.mc_samples : 1;
  p.la = c.loss_attenuation;
  p.method = c.decode_method;
  p.anchors = ctx->anchors;
  p.N = ctx->num_anchors;
  int tile_px = (kThreads / 2) / p.A;  // two threads (y axis, x axis) per anchor
  const int cap = 12288 / (p.A * p.C);
  if (tile_px > cap) tile_px = cap;
  if (tile_px < 1) tile_px = 1;
  p.tile_px = tile_px;
  int off = 0;
  for (int l = 0; l < UDAL_MAX_LEVELS + 1; ++l) {
    p.tile_off[l] = off;
    if (l < c.num_levels) {
      const int hw = c.level_h[l] * c.level_w[l];
      off += (hw + tile_px - 1) / tile_px;
    }
  }
  *total_tiles = off;
  p.stage_off = (tile_px * p.A * p.C + 3) & ~3;
  *smem = (size_t)p.stage_off * sizeof(float);
  for (int l = 0; l < UDAL_MAX_LEVELS; ++l) {
    p.cls.p[l] = (cls && l < c.num_levels) ? cls[l] : nullptr;
    p.box.p[l] = (box && l < c.num_levels) ? box[l] : nullptr;
    if (l < c.num_levels) {
      if (cls) UDAL_REQUIRE(cls[l] != nullptr, "cls level %d is NULL", l);
      if (box) {
        UDAL_REQUIRE(box[l] != nullptr, "box level %d is NULL", l);
        UDAL_REQUIRE(((uintptr_t)box[l] & 15) == 0, "box level %d must be 16-byte aligned", l);
      }
    }
  }
  return UDAL_OK;
}

int pick_tmax(int T) {
  if (T <= 1) return 1;
  if (T <= 4) return 4;
  if (T <= 10) return 10;
  if (T <= 16) return 16;
  if (T <= 32) return 32;
  return 0;
}

}  // namespace

int udal_decode_stream = 1;  // UDAL_DECODE_FP32: the persistent TMA-staged kernel (0: the per-tile kernel; comparison switch)
int udal_decode_chunk = 4;  // samples per load chunk of the 17..32-sample decode path (2 | 4 | 8; tuning switch)

int udal_launch_decode_moments(udal_ctx* ctx, const float* const* cls, const float* const* box,
                               int batch, const udal_prenms_out* out) {
  DecodeParams p;
  int tiles;
  size_t smem;
  UDAL_TRY(fill_params(ctx, cls, box, batch, p, &tiles, &smem));
  p.out = *out;
  const int T = p.Tc > p.Tb ? p.Tc : p.Tb;
  dim3 grid(tiles, batch);
  if (ctx->cfg.decode_precision == UDAL_DECODE_FP32 && udal_decode_stream && p.method != UDAL_DECODE_FALSEDEC && (p.BC & 3) == 0) {
    // persistent TMA-staged kernel: tiles of 16 pixels, thread = anchor
    StreamParams sp;
    sp.d = p;
    sp.d.tile_px = kStreamPx;
    int off = 0;
    for (int l = 0; l < UDAL_MAX_LEVELS + 1; ++l) {
      sp.d.tile_off[l] = off;
      if (l < ctx->cfg.num_levels) off += (ctx->cfg.level_h[l] * ctx->cfg.level_w[l] + kStreamPx - 1) / kStreamPx;
    }
    sp.tiles = off;
    sp.items = off * batch;
    sp.cls_floats = (kStreamPx * p.A * p.C + 3) & ~3;
    sp.stage_floats = sp.cls_floats + kStreamPx * p.BC;
    sp.consumers = (kStreamPx * p.A + 31) / 32 * 32;
    const int pairs = sp.cls_floats / 2;
    const int ept2 = (pairs + sp.consumers - 1) / sp.consumers;
    if (sp.consumers <= 160 && ept2 <= 8 && (size_t)sp.stage_floats * 4 * 2 + (size_t)2 * sp.cls_floats * 4 < 100 * 1024) {
      if (p.Tc == 1 && p.Tb == 1) {
        if (ept2 <= 4) return launch_stream<4, true>(ctx, sp);
        if (ept2 <= 5) return launch_stream<5, true>(ctx, sp);
        return launch_stream<8, true>(ctx, sp);
      }
      if (ept2 <= 4) return launch_stream<4, false>(ctx, sp);
      if (ept2 <= 5) return launch_stream<5, false>(ctx, sp);
      return launch_stream<8, false>(ctx, sp);
    }
  }
  if (ctx->cfg.decode_precision == UDAL_DECODE_FP32) {
    if (T <= 4) decode_moments_f32_kernel<2><<<grid, kThreads, smem, ctx->stream>>>(p);
    else decode_moments_f32_kernel<5><<<grid, kThreads, smem, ctx->stream>>>(p);
    UDAL_CHECK_LAUNCH(ctx);
    return UDAL_OK;
  }
  const bool fast = p.la && p.method == UDAL_DECODE_LNORM && p.box_mc && p.cls_mc && p.Tb == p.Tc;
  const int chunk = udal_decode_chunk == 2 || udal_decode_chunk == 8 ? udal_decode_chunk : 4;
  if (pick_tmax(T) > 16) smem += (size_t)(2 * p.Tb + 4 * chunk) * kThreads * sizeof(float);
  else if (pick_tmax(T) != 0) smem += (size_t)4 * p.Tb * kThreads * sizeof(float);
  UDAL_REQUIRE(smem <= 200 * 1024, "decode_moments: T=%d needs %zu bytes of shared memory", T, smem);
#define LAUNCH_ONE(TM, F, ...)                                                                                           \
  {                                                                                                                      \
    if (smem > 48 * 1024)                                                                                                \
      UDAL_CUDA(cudaFuncSetAttribute(decode_moments_kernel<TM, F, __VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                                        \
    decode_moments_kernel<TM, F, __VA_ARGS__><<<grid, kThreads, smem, ctx->stream>>>(p);                                 \
  }
#define LAUNCH(TM)                              \
  if (fast && T == TM) LAUNCH_ONE(TM, true, 4)  \
  else LAUNCH_ONE(TM, false, 4)                 \
  break;
  switch (pick_tmax(T)) {
    case 1: LAUNCH(1)
    case 4: LAUNCH(4)
    case 10: LAUNCH(10)
    case 16: LAUNCH(16)
    case 32:
      if (chunk == 2) LAUNCH_ONE(32, false, 2)
      else if (chunk == 8) LAUNCH_ONE(32, false, 8)
      else if (T <= 24) LAUNCH_ONE(32, false, 4, true)  // (2 T + 16) KB + the logit tile: three CTAs fit an SM
      else LAUNCH_ONE(32, false, 4)
      break;
    default: LAUNCH_ONE(0, false, 4) break;
  }
#undef LAUNCH_ONE
#undef LAUNCH
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_launch_logit_moments(udal_ctx* ctx, const float* const* cls, int batch, float* mean_logits,
                              float* std_logits) {
  DecodeParams p;
  int tiles;
  size_t smem;
  UDAL_TRY(fill_params(ctx, cls, nullptr, batch, p, &tiles, &smem));
  memset(&p.out, 0, sizeof(p.out));
  p.out.mean_logits = mean_logits;
  p.out.std_logits = std_logits;
  dim3 grid(tiles, batch);
#define LAUNCH(TM)                                                        \
  logit_moments_kernel<TM><<<grid, kThreads, 0, ctx->stream>>>(p);        \
  break;
  switch (pick_tmax(p.Tc)) {
    case 1: LAUNCH(1)
    case 4: LAUNCH(4)
    case 10: LAUNCH(10)
    case 16: LAUNCH(16)
    case 32: LAUNCH(32)
    default: LAUNCH(0)
  }
#undef LAUNCH
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_launch_decode_gather(udal_ctx* ctx, const float* const* box, int batch, int k,
                              const int32_t* topk_idx, const float* topk_val, const float* std_logits,
                              const udal_prenms_topk_out* out) {
  DecodeParams dp;
  int tiles;
  size_t smem;
  UDAL_TRY(fill_params(ctx, nullptr, box, batch, dp, &tiles, &smem));
  GatherParams p;
  p.box = dp.box;
  p.geom = dp.geom;
  p.batch = batch;
  p.A = dp.A;
  p.C = dp.C;
  p.BC = dp.BC;
  p.Tb = dp.Tb;
  p.box_mc = dp.box_mc;
  p.la = dp.la;
  p.method = dp.method;
  p.k = k;
  p.anchors = dp.anchors;
  p.N = dp.N;
  p.topk_idx = topk_idx;
  p.topk_val = topk_val;
  p.std_logits = std_logits;
  p.out = *out;
  dim3 grid((k + kThreads - 1) / kThreads, batch);
  if (ctx->cfg.decode_precision == UDAL_DECODE_FP32) decode_gather_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(p);
  else decode_gather_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
