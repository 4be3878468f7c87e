// fp64 exp / sqrt tuned for the decode and soft-NMS kernels (shared by decode_moments.cu / nms.cu).
#pragma once
#include <cuda_runtime.h>

// The decode needs fp64 results that round to the same fp32 value as the reference's libm: a
// relative error of ~1e-15 is as good as 0.5 ulp for that purpose (mismatch probability ~
// error / 2^-24).  exp(x) = 2^k * 2^(j/64) * exp(r), |r| <= ln2/128, degree-5 polynomial (truncation
// 3.5e-17), 64-entry table in shared memory; sqrt by rsqrt.approx + two Newton corrections.
// Arguments outside the fast range fall back to the CUDA math library.
static __constant__ double kExp2Table[64] = {
    1.0,
    1.0108892860517005,
    1.0218971486541166,
    1.0330248790212284,
    1.0442737824274138,
    1.0556451783605572,
    1.0671404006768237,
    1.0787607977571199,
    1.0905077326652577,
    1.102382583307841,
    1.1143867425958924,
    1.1265216186082418,
    1.1387886347566916,
    1.1511892299529827,
    1.1637248587775775,
    1.1763969916502812,
    1.189207115002721,
    1.202156731452703,
    1.215247359980469,
    1.22848053610687,
    1.241857812073484,
    1.255380757024691,
    1.2690509571917332,
    1.2828700160787783,
    1.2968395546510096,
    1.3109612115247644,
    1.3252366431597413,
    1.339667524053303,
    1.3542555469368927,
    1.3690024229745905,
    1.383909881963832,
    1.3989796725383112,
    1.4142135623730951,
    1.42961333839197,
    1.4451808069770467,
    1.460917794180647,
    1.4768261459394993,
    1.4929077282912648,
    1.5091644275934228,
    1.5255981507445384,
    1.5422108254079407,
    1.559004400237837,
    1.5759808451078865,
    1.593142151342267,
    1.6104903319492543,
    1.6280274218573478,
    1.645755478153965,
    1.6636765803267364,
    1.681792830507429,
    1.7001063537185235,
    1.718619298122478,
    1.7373338352737062,
    1.7562521603732995,
    1.7753764925265212,
    1.7947090750031072,
    1.8142521755003989,
    1.8340080864093424,
    1.8539791250833855,
    1.8741676341103,
    1.8945759815869656,
    1.9152065613971474,
    1.9360617934922943,
    1.9571441241754002,
    1.978456026387951};

static __device__ __noinline__ double exp_slow(double x) { return exp(x); }
static __device__ __noinline__ double sqrt_slow(double x) { return sqrt(x); }

static __device__ __forceinline__ double exp_fast(double x, const double* tbl) {
  if (!(fabs(x) < 690.0)) return exp_slow(x);
  const double n = rint(x * 92.33248261689366);
  double r = fma(n, -0.010830424696248286, x);
  r = fma(n, -8.59050471673183e-16, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int ni = (int)n;
  const double v = tbl[ni & 63] * p;
  return __longlong_as_double(__double_as_longlong(v) + ((long long)(ni >> 6) << 52));
}

static __device__ __forceinline__ double sqrt_fast(double x) {
  if (!(x > 1e-290 && x < 1e290)) return sqrt_slow(x);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y * 0.5, fma(-x * y, y, 1.0), y);  // Newton step on 1/sqrt(x)
  double s = x * y;
  s = fma(fma(-s, s, x), y * 0.5, s);        // correction of sqrt(x)
  return s;
}


