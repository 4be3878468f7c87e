// decode_uncert(method="sample") - utils_box.py:162-184: draw n_samples regression vectors per anchor from
// N(t, diag(sigma^2)), decode every draw with the plain exp/offset transform, return mean and population
// standard deviation of the decoded corners (tf.nn.moments over the sample axis), everything in float64 and
// rounded to fp32 at the end like the other methods (utils_box.py:125-137, 268-271).
//
// The reference draws from tfp's MultivariateNormalDiag (stateful TF RNG, not reproducible outside TF).  Here the
// standard normals are either INJECTED (parity runs: normals[s][k][i], the oracle consumes the same array) or drawn
// in-kernel: Philox4x32-10, counter = s * n + i, key = seed, the four 32-bit words -> two Box-Muller pairs ->
// (z_y, z_x, z_h, z_w); uncertainty-detection-autolabeling_b200/utils_box.py holds the NumPy twin of that stream.
//
// One thread per anchor, two passes over the samples (mean, then squared deviations): a comparison method of the
// paper, off the serving path - latency, not throughput, matters here.
#include "udal_common.cuh"

namespace {

__device__ __forceinline__ void ds_philox(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  uint32_t c2 = 0u, c3 = 0u;
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

// u in (0, 1): (x >> 8 + 0.5) * 2^-24
__device__ __forceinline__ double ds_unit(uint32_t x) { return ((double)(x >> 8) + 0.5) * (1.0 / 16777216.0); }

__device__ __forceinline__ void ds_normals(uint64_t counter, uint64_t seed, double (&z)[4]) {
  uint32_t r[4];
  ds_philox((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double two_pi = 6.283185307179586476925286766559;
  const double r0 = sqrt(-2.0 * log(ds_unit(r[0]))), a0 = two_pi * ds_unit(r[1]);
  const double r1 = sqrt(-2.0 * log(ds_unit(r[2]))), a1 = two_pi * ds_unit(r[3]);
  z[0] = r0 * cos(a0);
  z[1] = r0 * sin(a0);
  z[2] = r1 * cos(a1);
  z[3] = r1 * sin(a1);
}

__global__ void decode_sample_kernel(const float* __restrict__ pred, const float* __restrict__ sigma,
                                     const float* __restrict__ anchors, const float* __restrict__ normals, int64_t n,
                                     int n_samples, uint64_t seed, float* __restrict__ coords, float* __restrict__ stds) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 af = reinterpret_cast<const float4*>(anchors)[i];
  const float4 tf = reinterpret_cast<const float4*>(pred)[i];
  const float4 sf = reinterpret_cast<const float4*>(sigma)[i];
  const double a0 = af.x, a1 = af.y, a2 = af.z, a3 = af.w;
  const double yca = (a0 + a2) / 2, xca = (a1 + a3) / 2, ha = a2 - a0, wa = a3 - a1;
  const double t[4] = {tf.x, tf.y, tf.z, tf.w};
  // sqrt(square(sigma)) = |sigma|
  const double sd[4] = {sqrt((double)sf.x * (double)sf.x), sqrt((double)sf.y * (double)sf.y),
                        sqrt((double)sf.z * (double)sf.z), sqrt((double)sf.w * (double)sf.w)};
  auto corners = [&](int s, double (&c)[4]) {
    double z[4];
    if (normals) {
#pragma unroll
      for (int k = 0; k < 4; ++k) z[k] = (double)normals[((size_t)s * 4 + k) * (size_t)n + (size_t)i];
    } else {
      ds_normals((uint64_t)s * (uint64_t)n + (uint64_t)i, seed, z);
    }
    const double sy = t[0] + sd[0] * z[0], sx = t[1] + sd[1] * z[1], sh = t[2] + sd[2] * z[2], sw = t[3] + sd[3] * z[3];
    const double w = exp(sw) * wa, h = exp(sh) * ha;
    const double yc = sy * ha + yca, xc = sx * wa + xca;
    c[0] = yc - h / 2.0;
    c[1] = xc - w / 2.0;
    c[2] = yc + h / 2.0;
    c[3] = xc + w / 2.0;
  };
  double mean[4] = {0, 0, 0, 0}, var[4] = {0, 0, 0, 0};
  for (int s = 0; s < n_samples; ++s) {
    double c[4];
    corners(s, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) mean[k] += c[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) mean[k] /= (double)n_samples;
  for (int s = 0; s < n_samples; ++s) {
    double c[4];
    corners(s, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) var[k] += (c[k] - mean[k]) * (c[k] - mean[k]);
  }
  reinterpret_cast<float4*>(coords)[i] = make_float4((float)mean[0], (float)mean[1], (float)mean[2], (float)mean[3]);
  reinterpret_cast<float4*>(stds)[i] =
      make_float4((float)sqrt(var[0] / n_samples), (float)sqrt(var[1] / n_samples), (float)sqrt(var[2] / n_samples),
                  (float)sqrt(var[3] / n_samples));
}

}  // namespace

extern "C" int udal_decode_sample(udal_ctx* ctx, const float* pred, const float* sigma, const float* anchors, int64_t n,
                                  int n_samples, const float* normals, uint64_t seed, float* coords, float* stds) {
  UDAL_REQUIRE(ctx && pred && sigma && anchors && coords && stds, "NULL argument");
  UDAL_REQUIRE(n > 0 && n_samples > 0, "empty input");
  UDAL_REQUIRE((((uintptr_t)pred | (uintptr_t)sigma | (uintptr_t)anchors | (uintptr_t)coords | (uintptr_t)stds) & 15) == 0,
               "udal_decode_sample: [n,4] tensors must be 16-byte aligned");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  UDAL_TRY(udal_join(ctx));
  decode_sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(pred, sigma, anchors, normals, n, n_samples, seed,
                                                                            coords, stds);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
