// Micro-benchmark (development tool, not part of libudal): cycles per tcgen05.mma (M128, K16, bf16,
// A and B from shared memory) as a function of N and of the operand layout.  Answers: how fast is
// the A-operand fetch when every instruction touches only a 32-byte slice of a 128-byte swizzled
// row (the implicit-GEMM head kernel's case), and would a 32-byte-row layout be faster?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

struct Args {
  int n;          // UMMA N
  int layout;     // descriptor layout code: 2 = SWIZZLE_128B, 6 = SWIZZLE_32B, 4 = SWIZZLE_64B, 0 = none
  int a_sbo, b_sbo, a_lbo, b_lbo;
  int a_step, b_step;  // descriptor start advance (bytes) between consecutive instructions (cycled over 4)
  int a_shift;         // extra start offset of A (bytes)
  int iters;
  int commit_every;  // 0: one commit at the end; n: tcgen05.commit (x commits_per) after every n-th group of 4 MMAs
  int commits_per;
  int pollers;  // 0: none; 1: the other warps spin on mbarrier.try_wait; 2: same with __nanosleep(64) between polls;
                // 3: spin on test_wait; 4: spin on ld.shared of a flag
};

__global__ void __launch_bounds__(320, 1) umma_bench_kernel(Args a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  __shared__ uint64_t bar, bar2, bar3[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar3[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar3[1])));
    stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    // warp-uniform issue loop (descriptors stay in uniform registers), one elected lane issues
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = sb + a.a_shift, b0 = sb + 72 * 1024;
    uint32_t parity = 0;
    long long best = 1ll << 60, best_ns = 0;
    for (int rep = 0; rep < 3; ++rep) {
      unsigned long long g0, g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      const long long t0 = clock64();
      if (a.commit_every == 0) {
        for (int i = 0; i < a.iters; ++i) {
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_desc(a0 + k * a.a_step, a.a_lbo, a.a_sbo, a.layout);
              const uint64_t bd = make_desc(b0 + k * a.b_step, a.b_lbo, a.b_sbo, a.layout);
              mma(tmem, ad, bd, idesc, 1u);
            }
          }
          __syncwarp();
        }
      } else {
        // the head kernel's tile: 9 taps x 4 K slices, shifted A starts, alternating accumulators, commits per tile
        for (int i = 0; i < a.iters / 9; ++i) {
          const uint32_t d = tmem + ((i & 1) ? 128u : 0u);
          const uint32_t in0 = a0 + (uint32_t)(i % 3) * 23552u;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t ad = make_desc(in0 + (uint32_t)(((tap / 3) * 10 + tap % 3) * 128), 16, 1280, 2);
              const uint64_t bd = make_desc(b0 + tap * (uint32_t)a.n * 128u, 16, 1024, 2);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
            }
            for (int c = 0; c < a.commits_per; ++c)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar3[c])) : "memory");
          }
          __syncwarp();
        }
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      __syncwarp();
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s32(&bar)), "r"(parity)
            : "memory");
      }
      parity ^= 1;
      const long long t1 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      if (t1 - t0 < best) {
        best = t1 - t0;
        best_ns = (long long)(g1 - g0);
      }
    }
    if (threadIdx.x == 0) {
      out[blockIdx.x] = best;
      out[gridDim.x + blockIdx.x] = best_ns;
      stop_flag = 1;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar2)) : "memory");
    }
  } else if (threadIdx.x >= 64 && a.pollers) {
    uint32_t done = 0;
    while (!done) {
      if (a.pollers == 4) {
        done = stop_flag;
      } else if (a.pollers == 3) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s32(&bar2)), "r"(0)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s32(&bar2)), "r"(0)
            : "memory");
        if (a.pollers == 2 && !done) __nanosleep(64);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

static void run(const char* name, Args a, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid * 2);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_bench_kernel<<<grid, 320, smem>>>(a, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296] = {0};
  cudaMemcpy(h, d, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost);
  long long mx = 0, ns = 1;
  for (int i = 0; i < grid; ++i)
    if (h[i] > mx) {
      mx = h[i];
      ns = h[grid + i];
    }
  // SM clock during the measurement = cycles / globaltimer ns (the tensor pipe under load may run below the idle clock)
  printf("%-48s N=%3d grid=%3d pollers=%d commit_every=%d x%d iters=%d  %7.1f cycles/MMA  %6.0f MHz  (%s)\n", name, a.n, grid, a.pollers, a.commit_every * 4, a.commits_per, a.iters, (double)mx / (4.0 * a.iters), 1e3 * (double)mx / (double)ns, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 3600;
  for (int n : {64, 80}) {
    run("SW128 pitch-10 halo, fixed operands", Args{n, 2, 1280, 1024, 16, 16, 32, 32, 11 * 128, iters, 0, 1, 0}, 148);
    run("head-kernel tile pattern, 1 commit/tile", Args{n, 2, 1280, 1024, 16, 16, 32, 32, 0, iters, 9, 1, 0}, 148);
    run("head-kernel tile pattern, 2 commits/tile", Args{n, 2, 1280, 1024, 16, 16, 32, 32, 0, iters, 9, 2, 0}, 148);
    run("head-kernel tile pattern, 0 commits/tile", Args{n, 2, 1280, 1024, 16, 16, 32, 32, 0, iters, 9, 0, 0}, 148);
    // sustained: ~25 ms of back-to-back MMAs per repetition, long enough for the power management to settle
    run("head-kernel tile pattern, sustained", Args{n, 2, 1280, 1024, 16, 16, 32, 32, 0, iters * 70, 9, 1, 0}, 148);
  }
  return 0;
}
