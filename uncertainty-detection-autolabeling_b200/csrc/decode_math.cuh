// Per-axis anchor decode with exact moment propagation (utils_box.py:105-276) and the reference sigmoid,
// shared by the stand-alone decode kernel (decode_moments.cu) and the fused predict+decode head kernels
// (heads_fused.cu).
#pragma once
#include "fast_math64.cuh"
#include "udal.h"

namespace {

// One axis (y: ty/th, x: tx/tw) of decode_la - the two axes are independent, so the fused kernel
// gives each axis of an anchor its own thread (twice the parallelism, half the registers).
// exp(2 t + v) of utils_box.py:151-152 is taken as exp(t + v/2)^2: both are within 1 ulp(fp64) of
// the true value and round to the same fp32 result except in ~1e-8 of the cases.
__device__ __forceinline__ void decode_axis_la(int method, const double* tbl, float a_lo_f, float a_hi_f,
                                               float t_c_f, float t_s_f, float s_c_f, float s_s_f, float& lo,
                                               float& hi, float& sd_lo, float& sd_hi) {
  const double a_lo = a_lo_f, a_hi = a_hi_f, t_c = t_c_f, t_s = t_s_f;
  const double ca = __dmul_rn(__dadd_rn(a_lo, a_hi), 0.5);
  const double sa = __dsub_rn(a_hi, a_lo);
  const double vc = __dmul_rn((double)s_c_f, (double)s_c_f);
  const double vs = __dmul_rn((double)s_s_f, (double)s_s_f);
  const double c = __dadd_rn(__dmul_rn(t_c, sa), ca);
  if (method == UDAL_DECODE_FALSEDEC) {
    const double half = __dmul_rn(__dmul_rn(exp_fast(t_s, tbl), sa), 0.5);
    lo = (float)__dsub_rn(c, half);
    hi = (float)__dadd_rn(c, half);
    const double dhalf = __dmul_rn(__dmul_rn(exp_fast(vs, tbl), sa), 0.5);
    const double dc = __dadd_rn(__dmul_rn(vc, sa), ca);
    sd_lo = (float)sqrt(fabs(__dsub_rn(dc, dhalf)));
    sd_hi = (float)sqrt(__dadd_rn(dc, dhalf));
    return;
  }
  const double e = exp_fast(__dadd_rn(t_s, __dmul_rn(vs, 0.5)), tbl);
  const double half = __dmul_rn(__dmul_rn(e, sa), 0.5);
  lo = (float)__dsub_rn(c, half);
  hi = (float)__dadd_rn(c, half);
  double var_c, var_s;
  if (method == UDAL_DECODE_NFLOW) {
    const double q = fabs(__dmul_rn(sa, sqrt(vc)));
    var_c = __dmul_rn(q, q);
    const double ss = sqrt(vs);
    const double ss2 = __dmul_rn(ss, ss);
    const double lv = __dmul_rn(__dsub_rn(exp(ss2), 1.0), exp(__dadd_rn(__dmul_rn(2.0, t_s), ss2)));
    const double qs = fabs(__dmul_rn(sa, sqrt(lv)));
    var_s = __dmul_rn(qs, qs);
  } else {
    var_s = __dmul_rn(__dmul_rn(__dsub_rn(exp_fast(vs, tbl), 1.0), __dmul_rn(e, e)), __dmul_rn(sa, sa));
    var_c = __dmul_rn(vc, __dmul_rn(sa, sa));
  }
  sd_lo = sd_hi = (float)sqrt_fast(__dadd_rn(var_c, __dmul_rn(var_s, 0.25)));
}

__device__ __forceinline__ void decode_axis_plain(float a_lo, float a_hi, float t_c, float t_s, float& lo, float& hi) {
  const float ca = __fmul_rn(__fadd_rn(a_lo, a_hi), 0.5f);
  const float sa = __fsub_rn(a_hi, a_lo);
  const float half = __fmul_rn(__fmul_rn((float)exp((double)t_s), sa), 0.5f);
  const float c = __fadd_rn(__fmul_rn(t_c, sa), ca);
  lo = __fsub_rn(c, half);
  hi = __fadd_rn(c, half);
}

__device__ __forceinline__ float sigmoid_ref(float x) {
  // oracle: fp32(1 / (1 + exp(-fp64(x))))
  return (float)(1.0 / (1.0 + exp(-(double)x)));
}

}  // namespace
