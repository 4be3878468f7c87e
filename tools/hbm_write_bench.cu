// Micro-benchmark: HBM write-only, read-only and copy bandwidth (4 GiB buffers, vectorised grid-stride kernels).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_write_bench tools/hbm_write_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_write(uint4* p, size_t n, uint4 v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_read(const uint4* p, size_t n, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}
__global__ void k_copy(const uint4* a, uint4* b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
  const size_t bytes = 4ull << 30, n = bytes / 16;
  uint4 *a, *b;
  unsigned* out;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * 8, block = 512;
  for (int mode = 0; mode < 4; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_write<<<grid, block>>>(a, n, make_uint4(rep, 1, 2, 3));
      if (mode == 1) k_read<<<grid, block>>>(a, n, out);
      if (mode == 2) k_copy<<<grid, block>>>(a, b, n);
      if (mode == 3) cudaMemsetAsync(a, rep, bytes);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    const double gb = (mode == 2 ? 2.0 : 1.0) * bytes / 1e9;
    printf("%-12s %7.3f ms  %7.1f GB/s\n", mode == 0 ? "write" : mode == 1 ? "read" : mode == 2 ? "copy (r+w)" : "cudaMemset", best, gb / (best / 1e3));
  }
  return 0;
}
