// Micro-benchmark: fp32 FMA throughput per SM, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 16 independent chains
// per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench tools/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <bool PACKED>
__global__ void __launch_bounds__(512) k(float* out, float x, float y, int iters, long long* cycles) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 a = make_float2(x, x * 1.0001f), b = make_float2(y, y * 0.9999f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        acc[i] = ffma2(acc[i], a, b);
      } else {
        acc[i].x = fmaf(acc[i].x, a.x, b.x);
        acc[i].y = fmaf(acc[i].y, a.y, b.y);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(512) kmufu(float* out, float x, int iters, long long* cycles) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = x + 0.01f * (threadIdx.x + i);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(acc[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(acc[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 2; ++rep) {
      if (packed) k<true><<<148, 512>>>(out, 0.999f, 0.001f, iters, cyc);
      else k<false><<<148, 512>>>(out, 0.999f, 0.001f, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double fma_per_clk = 512.0 * 16.0 * iters / (double)mx;
    printf("%s: %.1f fp32 FMA / clk / SM (16 warps, 16 chains per thread)\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar  ", fma_per_clk);
  }
  for (int op = 0; op < 3; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) kmufu<0><<<148, 512>>>(out, 0.3f, iters, cyc);
      if (op == 1) kmufu<1><<<148, 512>>>(out, -0.3f, iters, cyc);
      if (op == 2) kmufu<2><<<148, 512>>>(out, 1.3f, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s: %.1f / clk / SM\n", op == 0 ? "tanh.approx" : op == 1 ? "ex2.approx " : "rcp.approx ", 512.0 * 8.0 * iters / (double)mx);
  }
  return 0;
}
