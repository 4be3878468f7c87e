// K1 (bf16 tensor-core mode): the class / box(+sigma) head towers on tcgen05 (sm_100a).
//
// Replaces (reference src/): efficientdet_keras.py:353-513 ClassNet, 516-692 BoxNet, 979-1050 MC
// loop - same arithmetic graph as heads_fp32.cu, with the dense contraction (pointwise 1x1 and the
// class / box projections) on the 5th-generation tensor cores:
//
//   per CTA: one 8x16-pixel tile (M = 128 rows of the GEMM) of one pyramid level
//     1. the bf16 input tile + 1-pixel halo is staged in shared memory (zero padded)
//     2. depthwise 3x3 on the CUDA cores (fp32 math, 4 channels x 8 rows per thread, sliding
//        window) writes the A operand [128 x 64] bf16 straight into the canonical K-major
//        128-byte-swizzled UMMA layout
//     3. for each MC sample handled by this CTA: the B operand [N x 64] = folded weights
//        (pointwise * BN scale) * dropout keep-scale of the producer layer, bf16, same layout;
//        one elected thread issues 4 x tcgen05.mma (M128, N = 64|80, K16) accumulating in TMEM;
//        tcgen05.commit -> mbarrier; 8 warps read the accumulator with tcgen05.ld (32x32b),
//        add the folded bias, apply swish and store bf16 activations (or fp32 predictions).
//   SpatialDropout2D is a per-(sample, image, channel) scale; it commutes with the depthwise conv,
//   so it is folded into the rows of B.  Layer 1 therefore computes its depthwise conv ONCE per
//   image and loops the T samples in-kernel with the A tile resident in shared memory; layer 0
//   (sample invariant) runs once per image.
//
// Numerics: bf16 operands / activations, fp32 accumulation; measured tolerance in
// tests/test_gpu_heads.py (the fp32 path of heads_fp32.cu is the parity mode).
#include <cuda_bf16.h>

#include "udal_common.cuh"

namespace {

constexpr int TH = 8, TW = 16, TPX = TH * TW;
constexpr int HALO_W = TW + 2, HALO_H = TH + 2, HALO_PX = HALO_W * HALO_H;
constexpr int KF = 64;            // filters = GEMM K (bf16: one 128-byte swizzle row)
constexpr int kThreads = 256;
constexpr int kMaxN = 80;

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int SM_A = 0;                         // [128][128 B]
constexpr int SM_B = SM_A + TPX * 128;          // 2 x [kMaxN][128 B] (double buffered over the samples)
constexpr int SM_IN = SM_B + 2 * kMaxN * 128;   // [180][64] bf16
constexpr int SM_DW = SM_IN + HALO_PX * KF * 2; // [9][64] fp32
constexpr int SM_FB = SM_DW + 9 * KF * 4;       // [kMaxN] fp32 folded bias
constexpr int kMaxT = 16;                       // samples whose keep-scales are staged in shared memory
constexpr int SM_SC = SM_FB + kMaxN * 4;        // [2][kMaxT][64] fp32: producer-layer scales, this layer's scales
constexpr int SM_MBAR = SM_SC + 2 * kMaxT * KF * 4;  // 2 mbarriers (16 B) + tmem base (4 B)
constexpr int SM_TOTAL = SM_MBAR + 32;
constexpr int SM_ALLOC = SM_TOTAL + 1024;       // slack for the manual 1024-byte alignment

struct TcLayerParams {
  int num_levels;
  int h[UDAL_MAX_LEVELS], w[UDAL_MAX_LEVELS];
  int tiles_x[UDAL_MAX_LEVELS];
  int tile_off[UDAL_MAX_LEVELS + 1];
  const void* in[UDAL_MAX_LEVELS];     // [NB_in,H,W,64] bf16 (or fp32 BiFPN features for layer 0)
  void* out[UDAL_MAX_LEVELS];          // [NB_out,H,W,64] bf16 or [NB_out,H,W,Cout] fp32
  const float* scale[UDAL_MAX_LEVELS]; // [NB_out,64] keep-scale of the producer layer (folded into B), or null
  const float* out_scale[UDAL_MAX_LEVELS]; // [NB_out,64] keep-scale of THIS layer's dropout (applied at store), or null
  const float* wf[UDAL_MAX_LEVELS];    // folded weights [Npad][64] fp32
  const float* fb[UDAL_MAX_LEVELS];    // folded bias [Npad]
  const float* dw;                     // [9][64]
  int Cout, Npad, batch, nt;
  int in_fp32, out_fp32, act, in_per_sample;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::make_umma_desc<Major::K>)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float swish_fast(float x) {
  // x * sigmoid(x), sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5 : one MUFU per element
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads, NPAD <= 64 ? 3 : 2) sepconv_tc_kernel(const TcLayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint8_t* sA = smem + SM_A;
  uint8_t* sB = smem + SM_B;
  uint8_t* sIn = smem + SM_IN;
  float* sDw = reinterpret_cast<float*>(smem + SM_DW);
  float* sFb = reinterpret_cast<float*>(smem + SM_FB);
  float* sSc = reinterpret_cast<float*>(smem + SM_SC);
  const uint32_t mbar = sbase + SM_MBAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_MBAR + 16);
  constexpr uint32_t kTmemCols = NPAD <= 64 ? 128 : 256;  // two accumulators

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int l = 0;
#pragma unroll
  for (int i = 1; i < UDAL_MAX_LEVELS; ++i)
    if (i < p.num_levels && (int)blockIdx.x >= p.tile_off[i]) l = i;
  const int H = p.h[l], W = p.w[l];
  const int tile = blockIdx.x - p.tile_off[l];
  const int ty0 = (tile / p.tiles_x[l]) * TH, tx0 = (tile % p.tiles_x[l]) * TW;
  const int nb0 = blockIdx.y;
  const int in_img = p.in_per_sample ? nb0 : nb0 % p.batch;

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 8, 1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_MBAR + 16),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- stage depthwise weights, folded bias and the input tile (+halo) ----
  for (int e = tid; e < 9 * KF; e += kThreads) sDw[e] = __ldg(p.dw + e);
  if (tid < NPAD) sFb[tid] = __ldg(p.fb[l] + tid);
  const bool sc_smem = p.nt <= kMaxT;  // keep-scales of all samples of this CTA staged once
  if (sc_smem) {
    for (int e = tid; e < p.nt * KF; e += kThreads) {
      const int it = e / KF, k = e - it * KF;
      const size_t row = (size_t)(it * p.batch + nb0) * KF + k;
      sSc[e] = p.scale[l] ? __ldg(p.scale[l] + row) : 1.0f;
      sSc[kMaxT * KF + e] = p.out_scale[l] ? __ldg(p.out_scale[l] + row) : 1.0f;
    }
  }
  if (p.in_fp32) {
    const float* src0 = reinterpret_cast<const float*>(p.in[l]) + (size_t)in_img * H * W * KF;
    for (int e = tid; e < HALO_PX * 8; e += kThreads) {
      const int px = e >> 3, ch = e & 7;
      const int hy = px / HALO_W;
      const int y = ty0 + hy - 1, x = tx0 + (px - hy * HALO_W) - 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float4* src = reinterpret_cast<const float4*>(src0 + ((size_t)y * W + x) * KF + ch * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        v.x = pack_bf16(a.x, a.y);
        v.y = pack_bf16(a.z, a.w);
        v.z = pack_bf16(b.x, b.y);
        v.w = pack_bf16(b.z, b.w);
      }
      *reinterpret_cast<uint4*>(sIn + (size_t)px * 128 + ch * 16) = v;
    }
  } else {
    // bf16 activations: 16-byte cp.async straight into shared memory, zero-filled outside the image
    const __nv_bfloat16* src0 = reinterpret_cast<const __nv_bfloat16*>(p.in[l]) + (size_t)in_img * H * W * KF;
    const uint32_t dst0 = sbase + SM_IN;
#pragma unroll
    for (int i = 0; i < (HALO_PX * 8 + kThreads - 1) / kThreads; ++i) {
      const int e = tid + i * kThreads;
      if (e < HALO_PX * 8) {
        const int px = e >> 3, ch = e & 7;
        const int hy = px / HALO_W;
        const int y = ty0 + hy - 1, x = tx0 + (px - hy * HALO_W) - 1;
        const bool ok = y >= 0 && y < H && x >= 0 && x < W;
        const __nv_bfloat16* src = ok ? src0 + ((size_t)y * W + x) * KF + ch * 8 : src0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)e * 16u), "l"(src),
                     "r"(ok ? 16 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- depthwise 3x3: thread = (channel quad q, column x), 8 output rows, sliding window ----
  {
    const int q = tid & 15, x = tid >> 4;
    float wgt[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w4 = *reinterpret_cast<const float4*>(sDw + t * KF + q * 4);
      wgt[t][0] = w4.x;
      wgt[t][1] = w4.y;
      wgt[t][2] = w4.z;
      wgt[t][3] = w4.w;
    }
    float acc[TH][4];
#pragma unroll
    for (int y = 0; y < TH; ++y)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[y][c] = 0.f;
#pragma unroll
    for (int r = 0; r < HALO_H; ++r) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint2 raw2 = *reinterpret_cast<const uint2*>(sIn + (size_t)(r * HALO_W + x + dx) * 128 + q * 8);
        float v[4];
        v[0] = __uint_as_float(raw2.x << 16);
        v[1] = __uint_as_float(raw2.x & 0xffff0000u);
        v[2] = __uint_as_float(raw2.y << 16);
        v[3] = __uint_as_float(raw2.y & 0xffff0000u);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int y = r - dy;
          if (y >= 0 && y < TH) {
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[y][c] = fmaf(v[c], wgt[dy * 3 + dx][c], acc[y][c]);
          }
        }
      }
      if (r >= 2) {
        const int y = r - 2;
        const int m = y * TW + x;
        uint2 o;
        o.x = pack_bf16(acc[y][0], acc[y][1]);
        o.y = pack_bf16(acc[y][2], acc[y][3]);
        *reinterpret_cast<uint2*>(sA + (size_t)m * 128 + (((q >> 1) ^ (m & 7)) << 4) + (q & 1) * 8) = o;
      }
    }
  }

  // instruction descriptor: D = f32, A = B = bf16, K-major both, N = NPAD, M = 128
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t adesc = umma_desc(sbase + SM_A);
  constexpr int HALF = NPAD / 2;
  const int m = (warp & 3) * 32 + lane;       // accumulator row = pixel of the tile
  const int col0 = (warp >> 2) * HALF;        // this warp's slice of the N columns
  const int oy = ty0 + m / TW, ox = tx0 + m % TW;
  const bool pix_ok = oy < H && ox < W;

  // ---- this thread's slice of the folded weights stays in registers for every sample ----
  constexpr int NCH = (NPAD * 8 + kThreads - 1) / kThreads;  // 16-byte chunks of B per thread
  float wreg[NCH][8];
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    const int e = tid + u * kThreads;
    if (e < NPAD * 8) {
      const float4* wsrc = reinterpret_cast<const float4*>(p.wf[l] + (size_t)(e >> 3) * KF + (e & 7) * 8);
      const float4 a = __ldg(wsrc), b = __ldg(wsrc + 1);
      wreg[u][0] = a.x; wreg[u][1] = a.y; wreg[u][2] = a.z; wreg[u][3] = a.w;
      wreg[u][4] = b.x; wreg[u][5] = b.y; wreg[u][6] = b.z; wreg[u][7] = b.w;
    }
  }
  // B operand of sample `it` -> buffer it & 1: folded weights * keep-scale of the producer layer
  auto prep_b = [&](int it) {
    const float* sc = p.scale[l] ? p.scale[l] + (size_t)(it * p.batch + nb0) * KF : nullptr;
    uint8_t* dstB = sB + (size_t)(it & 1) * (kMaxN * 128);
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int e = tid + u * kThreads;
      if (e < NPAD * 8) {
        const int n = e >> 3, c = e & 7;
        float w8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = wreg[u][i];
        if (sc) {
          float4 s0, s1;
          if (sc_smem) {
            s0 = *reinterpret_cast<const float4*>(sSc + it * KF + c * 8);
            s1 = *reinterpret_cast<const float4*>(sSc + it * KF + c * 8 + 4);
          } else {
            s0 = __ldg(reinterpret_cast<const float4*>(sc + c * 8));
            s1 = __ldg(reinterpret_cast<const float4*>(sc + c * 8) + 1);
          }
          w8[0] *= s0.x; w8[1] *= s0.y; w8[2] *= s0.z; w8[3] *= s0.w;
          w8[4] *= s1.x; w8[5] *= s1.y; w8[6] *= s1.z; w8[7] *= s1.w;
        }
        uint4 v;
        v.x = pack_bf16(w8[0], w8[1]);
        v.y = pack_bf16(w8[2], w8[3]);
        v.z = pack_bf16(w8[4], w8[5]);
        v.w = pack_bf16(w8[6], w8[7]);
        *reinterpret_cast<uint4*>(dstB + (size_t)n * 128 + ((c ^ (n & 7)) << 4)) = v;
      }
    }
  };
  auto issue_mma = [&](int it) {  // one thread: 4 x tcgen05.mma into accumulator it & 1, then commit
    tc_fence_after();
    const uint64_t bd = umma_desc(sbase + SM_B + (uint32_t)(it & 1) * (kMaxN * 128));
    const uint32_t d_tmem = tmem_base + (uint32_t)((it & 1) * NPAD);
#pragma unroll
    for (int k = 0; k < KF / 16; ++k)  // +32 bytes per K=16 step inside the 128-byte swizzle row
      umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     mbar + 8u * (uint32_t)(it & 1))
                 : "memory");
  };

  // software pipeline over the samples: while the tensor core works on sample it+1 the CTA runs the
  // epilogue of sample it; B operands and accumulators are double buffered
  prep_b(0);
  fence_async_smem();   // generic-proxy smem writes (A, B) -> visible to the tensor-core proxy
  tc_fence_before();
  __syncthreads();
  if (tid == 0) issue_mma(0);

  for (int it = 0; it < p.nt; ++it) {
    const int nb = it * p.batch + nb0;
    if (it + 1 < p.nt) {
      // buffer (it+1)&1 was last read by the MMA of sample it-1, whose completion every thread
      // observed in the previous epilogue; accumulator (it+1)&1 was drained there as well
      prep_b(it + 1);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) issue_mma(it + 1);
    }
    mbar_wait(mbar + 8u * (uint32_t)(it & 1), (uint32_t)((it >> 1) & 1));
    tc_fence_after();

    // ---- epilogue: TMEM -> registers -> bias / swish -> global ----
    uint32_t r[HALF / 8][8];
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((it & 1) * NPAD + col0);
#pragma unroll
    for (int j = 0; j < HALF / 8; ++j) tmem_ld8(taddr + j * 8, r[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (pix_ok) {
      const size_t pix = ((size_t)nb * H + oy) * W + ox;
      const float* fb = sFb + col0;
      const float* osc = p.out_scale[l] ? p.out_scale[l] + (size_t)nb * KF + col0 : nullptr;
      if (!p.out_fp32) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out[l]) + pix * KF + col0;
#pragma unroll
        for (int j = 0; j < HALF / 8; ++j) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = __uint_as_float(r[j][i]) + fb[j * 8 + i];
            if (p.act) v[i] = swish_fast(v[i]);
          }
          if (osc) {
            float4 s0, s1;
            if (sc_smem) {
              s0 = *reinterpret_cast<const float4*>(sSc + (kMaxT + it) * KF + col0 + j * 8);
              s1 = *reinterpret_cast<const float4*>(sSc + (kMaxT + it) * KF + col0 + j * 8 + 4);
            } else {
              s0 = __ldg(reinterpret_cast<const float4*>(osc + j * 8));
              s1 = __ldg(reinterpret_cast<const float4*>(osc + j * 8) + 1);
            }
            v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
            v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
          }
          uint4 o;
          o.x = pack_bf16(v[0], v[1]);
          o.y = pack_bf16(v[2], v[3]);
          o.z = pack_bf16(v[4], v[5]);
          o.w = pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(dst + j * 8) = o;
        }
      } else {
        float* dst = reinterpret_cast<float*>(p.out[l]) + pix * p.Cout + col0;
        const bool vec = (p.Cout & 3) == 0;
#pragma unroll
        for (int j = 0; j < HALF / 8; ++j) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = __uint_as_float(r[j][i]) + fb[j * 8 + i];
            if (p.act) v[i] = swish_fast(v[i]);
          }
          const int n = col0 + j * 8;
          if (vec && n + 8 <= p.Cout) {
            *reinterpret_cast<float4*>(dst + j * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + j * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n + i < p.Cout) dst[j * 8 + i] = v[i];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();

  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// folded weights: wf[n][k] = W[k][n] * bn_scale[n];  fb[n] = bias[n] * bn_scale[n] + bn_shift[n]
// columns [n0, n0 + cout) of the [64][ldw] weight matrix (ldw = 0: ldw = cout, n0 = 0)
__global__ void fold_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                    const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int cout,
                                    int npad, float* __restrict__ wf, float* __restrict__ fb, int n0 = 0, int ldw = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (ldw == 0) ldw = cout;
  if (i < npad * KF) {
    const int n = i / KF, k = i % KF;
    float v = 0.f;
    if (n < cout) v = w[(size_t)k * ldw + n0 + n] * (bn_scale ? bn_scale[n0 + n] : 1.f);
    wf[i] = v;
  }
  if (i < npad) {
    float v = 0.f;
    if (i < cout) v = bn_scale ? fmaf(bias[n0 + i], bn_scale[n0 + i], bn_shift[n0 + i]) : bias[n0 + i];
    fb[i] = v;
  }
}

int npad_of(int cout) { return cout <= 64 ? 64 : 80; }

}  // namespace

int udal_heads_ig_rows(int cout, int num_levels);
int udal_heads_ig_build_weights(udal_ctx* ctx, const float* dw, const float* wf, int nrows, void* wimg);
int udal_heads_ig_layer(udal_ctx* ctx, const void* const* in, int NB, const void* wimg, int rows,
                        const float* const* ep_scale, const float* const* ep_bias, int npad, int cout, int predict,
                        const float* const* out_scale, const float* ones, void* const* out, int ch_off = 0, int ch_total = 0);
int udal_heads_l1_layer(udal_ctx* ctx, const void* const* in, int NB, int T, const float* dw, const float* const* wf,
                        const float* const* fb, const float* const* in_scale, const float* const* out_scale,
                        const float* ones, float inv_keep, void* const* out);
int udal_heads_fused_predict(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const void* wimg, int rows,
                             const float* bias, const udal_prenms_out* pre);
int udal_heads_tc_use_ig = 1;  // 0: every layer through the per-tile kernel (debug / comparison; bf16 mode only)
// heads_dw.cu: fp16 mode, tower layers >= 2 and predict layers (depthwise on the CUDA cores, pointwise on tcgen05)
int udal_heads_dw_prepare(udal_ctx* ctx, int head);
int udal_heads_dw_layer(udal_ctx* ctx, int head, int layer, const void* const* in, int NB, const float* const* ep_scale,
                        const float* const* ep_bias, const float* const* out_scale, void* const* out);
int udal_heads_dw_fused_predict(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const udal_prenms_out* pre);
int udal_heads_dw_fused_l2(udal_ctx* ctx, int head, int layer, const void* const* in, int NB, int T, const float* const* ep_scale,
                           const float* const* ep_bias, const float* const* keep, const udal_prenms_out* pre);
// fp16 mode, udal_run: 1 = the last tower layer runs inside the fused predict kernels (heads_dwf_kernel<.., L2>: its output
// never reaches HBM; bit-identical results, tests/test_gpu_heads.py).  Measured on B200 (bench shape): 2.03 / 2.10 ms per head
// against 0.43 + 0.53 / 0.43 + 0.65 ms for the two kernels - the four builder warps (one per scheduler) then carry both
// depthwise passes and the tower epilogue as one dependent chain at ~0.2 IPC while the 12 statistics warps wait (ncu source view,
// DESIGN.md 6).  Off by default until that work is spread over the statistics warps.
int udal_heads_l2_fused = 0;

int udal_heads_l0_prepare(udal_ctx* ctx, int head);  // heads_wide.cu: tower layer 0 of the 64-channel heads
int udal_heads_l0_layer(udal_ctx* ctx, int head, const float* const* feats, int B, void* const* out);
extern int udal_heads_l0_persistent;
int udal_heads_wide_ok(const udal_ctx* ctx);  // heads_wide.cu: 64 < fpn_num_filters <= 128
int udal_heads_wide_prepare(udal_ctx* ctx, int head);
int udal_heads_wide_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                           float* const* box_out);

int udal_heads_tc_prepare(udal_ctx* ctx, int head) {
  const udal_config& c = ctx->cfg;
  udal_head_weights_dev& h = ctx->heads[head];
  if (udal_heads_wide_ok(ctx)) return udal_heads_wide_prepare(ctx, head);
  UDAL_REQUIRE(c.num_filters == KF, "the tensor-core head sampler is built for fpn_num_filters = 64 (D0) and 68..128 "
               "(D1, D2); got %d - use heads_mode fp32", c.num_filters);
  const int R = c.repeats, L = c.num_levels;
  // predict layers with more than 80 channels (C > 8 classes) run as equal chunks of at most 64 channels
  h.pred_chunks = h.cout <= kMaxN ? 1 : (h.cout + KF - 1) / KF;
  h.pred_chunk = h.pred_chunks == 1 ? h.cout : (h.cout + h.pred_chunks - 1) / h.pred_chunks;
  UDAL_REQUIRE(h.pred_chunks == 1 || R >= 2, "predict layers with more than %d channels need box_class_repeats >= 2", kMaxN);
  const int npad_p = h.pred_chunks == 1 ? npad_of(h.cout) : h.pred_chunks * KF;  // rows of the folded predict matrix
  const size_t n_w = (size_t)R * L * KF * KF + (size_t)npad_p * KF;
  const size_t n_b = (size_t)R * L * KF + (size_t)npad_p;
  if (h.pw_bf16) UDAL_CUDA(cudaFree(h.pw_bf16));
  if (h.fold_bias) UDAL_CUDA(cudaFree(h.fold_bias));
  h.pw_bf16 = nullptr;
  h.fold_bias = nullptr;
  UDAL_CUDA(cudaMalloc(&h.pw_bf16, n_w * sizeof(float)));
  UDAL_CUDA(cudaMalloc(&h.fold_bias, n_b * sizeof(float)));
  float* wf = reinterpret_cast<float*>(h.pw_bf16);
  for (int r = 0; r < R; ++r)
    for (int l = 0; l < L; ++l) {
      const size_t o = (size_t)r * L + l;
      fold_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(
          h.pw + (size_t)r * KF * KF, h.bias + (size_t)r * KF, h.bn_scale + o * KF, h.bn_shift + o * KF, KF, KF,
          wf + o * KF * KF, h.fold_bias + o * KF);
      UDAL_CHECK_LAUNCH(ctx);
    }
  if (h.pred_chunks == 1) {
    fold_weights_kernel<<<(npad_p * KF + 255) / 256, 256, 0, ctx->stream>>>(
        h.pwp, h.bp, nullptr, nullptr, h.cout, npad_p, wf + (size_t)R * L * KF * KF, h.fold_bias + (size_t)R * L * KF);
    UDAL_CHECK_LAUNCH(ctx);
  } else {
    for (int q = 0; q < h.pred_chunks; ++q) {
      const int n0 = q * h.pred_chunk, nc = h.cout - n0 < h.pred_chunk ? h.cout - n0 : h.pred_chunk;
      fold_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(
          h.pwp, h.bp, nullptr, nullptr, nc, KF, wf + (size_t)R * L * KF * KF + (size_t)q * KF * KF,
          h.fold_bias + (size_t)R * L * KF + (size_t)q * KF, n0, h.cout);
      UDAL_CHECK_LAUNCH(ctx);
    }
  }
  // implicit-GEMM weight images (level independent: BN scale is applied in the epilogue) for tower
  // layers >= 2 and the predict layer, plus a vector of ones as the predict layer's epilogue scale
  if (R >= 2) {
    const size_t tower_img = (size_t)9 * KF * KF, pred_img = (size_t)9 * npad_p * KF;
    const size_t n_img = (size_t)(R - 2) * tower_img + pred_img;
    if (h.ig_w) UDAL_CUDA(cudaFree(h.ig_w));
    h.ig_w = nullptr;
    UDAL_CUDA(cudaMalloc(&h.ig_w, n_img * 2 + (size_t)kMaxN * sizeof(float)));
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(h.ig_w);
    float* tmp;  // unscaled transposed pointwise weights [64][64] + dummy bias
    UDAL_TRY(udal_scratch_get(ctx, SCR_MISC, ((size_t)KF * KF + KF) * sizeof(float), (void**)&tmp));
    for (int r = 2; r < R; ++r) {
      fold_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(h.pw + (size_t)r * KF * KF, h.bias + (size_t)r * KF,
                                                                        nullptr, nullptr, KF, KF, tmp, tmp + KF * KF);
      UDAL_CHECK_LAUNCH(ctx);
      UDAL_TRY(udal_heads_ig_build_weights(ctx, h.dw + (size_t)r * 9 * KF, tmp, KF, img + (size_t)(r - 2) * tower_img));
    }
    if (h.pred_chunks == 1) {
      h.ig_rows = udal_heads_ig_rows(h.cout, L);  // rows per tap of the predict image (<= npad_p)
      UDAL_TRY(udal_heads_ig_build_weights(ctx, h.dwp, wf + (size_t)R * L * KF * KF, h.ig_rows,
                                           img + (size_t)(R - 2) * tower_img));
    } else {
      h.ig_rows = KF;
      for (int q = 0; q < h.pred_chunks; ++q)
        UDAL_TRY(udal_heads_ig_build_weights(ctx, h.dwp, wf + (size_t)R * L * KF * KF + (size_t)q * KF * KF, KF,
                                             img + (size_t)(R - 2) * tower_img + (size_t)q * tower_img));
    }
    // predict layer image of the fused predict + K2 kernels: 72 rows per tap (zero rows past cout) and an 80-entry bias,
    // or 96 / 96 for the 90 logits of a 10-class head
    if (h.fused_w) UDAL_CUDA(cudaFree(h.fused_w));
    h.fused_w = nullptr;
    h.fused_rows = 0;
    if (h.cout <= 96) {
      const int frows = h.cout <= 72 ? 72 : 96, fpad = h.cout <= 72 ? kMaxN : 96;
      const size_t img_bytes = (size_t)9 * frows * KF * 2;
      UDAL_CUDA(cudaMalloc(&h.fused_w, img_bytes + fpad * sizeof(float)));
      float* tmpw;
      UDAL_TRY(udal_scratch_get(ctx, SCR_MISC, ((size_t)fpad * KF + fpad) * sizeof(float), (void**)&tmpw));
      fold_weights_kernel<<<(fpad * KF + 255) / 256, 256, 0, ctx->stream>>>(
          h.pwp, h.bp, nullptr, nullptr, h.cout, fpad, tmpw, reinterpret_cast<float*>(reinterpret_cast<char*>(h.fused_w) + img_bytes));
      UDAL_CHECK_LAUNCH(ctx);
      UDAL_TRY(udal_heads_ig_build_weights(ctx, h.dwp, tmpw, frows, h.fused_w));
      UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
      h.fused_rows = frows;
    }
    std::vector<float> ones(kMaxN, 1.0f);
    UDAL_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h.ig_w) + n_img * 2, ones.data(), kMaxN * sizeof(float),
                              cudaMemcpyHostToDevice, ctx->stream));
    UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  UDAL_TRY(udal_heads_l0_prepare(ctx, head));
  if (c.heads_mode == UDAL_HEADS_FP16_TC) UDAL_TRY(udal_heads_dw_prepare(ctx, head));
  return UDAL_OK;
}

static int run_tower_tc(udal_ctx* ctx, int head, const float* const* feats, int batch, const float* scale_all,
                        float* const* outs, const udal_prenms_out* fused_pre) {
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  const int R = c.repeats, L = c.num_levels, T = c.mc_samples, B = batch;
  const bool mc = head == UDAL_HEAD_CLASS ? c.cls_mc != 0 : c.box_mc != 0;
  const int NBt = mc ? T * B : B;
  const size_t P = (size_t)ctx->num_pixels;
  __nv_bfloat16 *a0, *pp;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_A, (size_t)B * P * KF * 2, (void**)&a0));
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_B, 2 * (size_t)NBt * P * KF * 2, (void**)&pp));
  const float* wf_all = reinterpret_cast<const float*>(h.pw_bf16);
  TcLayerParams p;
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
  p.batch = B;
  UDAL_CUDA(cudaFuncSetAttribute(sepconv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  UDAL_CUDA(cudaFuncSetAttribute(sepconv_tc_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC));
  const int npad_p = h.pred_chunks == 1 ? npad_of(h.cout) : h.pred_chunks * KF;
  auto mark = [&]() {
    if (!ctx->profile_layers) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      cudaEventRecord(e, ctx->stream);
      ctx->layer_events.push_back(e);
    }
  };
  const void* l2_in[UDAL_MAX_LEVELS];
  const float *l2_scale[UDAL_MAX_LEVELS], *l2_bias[UDAL_MAX_LEVELS], *l2_keep[UDAL_MAX_LEVELS];
  bool l2_pending = false;
  for (int layer = 0; layer <= R; ++layer) {
    const bool predict = layer == R;
    mark();
    p.dw = predict ? h.dwp : h.dw + (size_t)layer * 9 * KF;
    p.Cout = predict ? h.cout : KF;
    p.Npad = predict ? npad_p : KF;
    p.act = predict ? 0 : 1;
    p.out_fp32 = predict ? 1 : 0;
    p.in_fp32 = layer == 0 ? 1 : 0;
    p.in_per_sample = layer >= 2 ? 1 : 0;
    // layer 1 reads the sample-invariant layer-0 output: one CTA per (tile, image) loops the samples
    p.nt = (layer == 1 && mc) ? T : 1;
    const int grid_y = layer <= 1 ? B : NBt;
    const bool fp16 = c.heads_mode == UDAL_HEADS_FP16_TC;
    const bool use_ig = (udal_heads_tc_use_ig || fp16) && layer >= 2;
    for (int l = 0; l < L; ++l) {
      const size_t lvl = (size_t)ctx->level_pix_off[l] * KF;
      if (layer == 0) p.in[l] = feats[l];
      else if (layer == 1) p.in[l] = a0 + (size_t)B * lvl;
      else p.in[l] = pp + (size_t)((layer - 1) & 1) * NBt * P * KF + (size_t)NBt * lvl;
      if (predict) p.out[l] = outs ? outs[l] : nullptr;
      else if (layer == 0) p.out[l] = a0 + (size_t)B * lvl;
      else p.out[l] = pp + (size_t)(layer & 1) * NBt * P * KF + (size_t)NBt * lvl;
      // SpatialDropout2D bookkeeping: layer 0 stores its (sample invariant) output WITHOUT dropout and
      // layer 1 folds that dropout (mask r = 0) into its B operand; every tower layer >= 1 applies its
      // own dropout (mask r = layer) when it stores, so later consumers read ready-made inputs.
      auto mask = [&](int r) { return scale_all + (((size_t)head * L + l) * R + r) * (size_t)NBt * KF; };
      p.scale[l] = (mc && layer == 1) ? mask(0) : nullptr;
      p.out_scale[l] = (mc && !predict && layer >= 1) ? mask(layer) : nullptr;
      if (predict) {
        p.wf[l] = wf_all + (size_t)R * L * KF * KF;
        p.fb[l] = h.fold_bias + (size_t)R * L * KF;
      } else {
        p.wf[l] = wf_all + ((size_t)layer * L + l) * KF * KF;
        p.fb[l] = h.fold_bias + ((size_t)layer * L + l) * KF;
      }
    }
    if (fp16 && layer >= 2) {
      // fp16 mode: depthwise on the CUDA cores (packed fp16) + one K = 64 GEMM per tile on tcgen05 (heads_dw.cu)
      if (fused_pre && udal_heads_l2_fused && mc && layer == R - 1) {
        // the last tower layer runs inside the fused predict kernel: remember its input and tables, launch nothing
        for (int l = 0; l < L; ++l) {
          l2_in[l] = p.in[l];
          l2_scale[l] = h.bn_scale + ((size_t)layer * L + l) * KF;
          l2_bias[l] = p.fb[l];
          l2_keep[l] = p.out_scale[l];
        }
        l2_pending = true;
        mark();   // (a zero-length layer in the profile)
        continue;
      }
      if (predict && fused_pre) {
        UDAL_REQUIRE(mc, "fused predict kernels: MC dropout on both heads");
        if (l2_pending) UDAL_TRY(udal_heads_dw_fused_l2(ctx, head, R - 1, l2_in, B, T, l2_scale, l2_bias, l2_keep, fused_pre));
        else UDAL_TRY(udal_heads_dw_fused_predict(ctx, head, p.in, B, T, fused_pre));
      } else {
        const float* ep_scale[UDAL_MAX_LEVELS];
        for (int l = 0; l < L; ++l) ep_scale[l] = predict ? nullptr : h.bn_scale + ((size_t)layer * L + l) * KF;
        UDAL_TRY(udal_heads_dw_layer(ctx, head, layer, p.in, NBt, ep_scale, p.fb, mc && !predict ? p.out_scale : nullptr, p.out));
      }
      mark();
      continue;
    }
    if ((fp16 || (udal_heads_tc_use_ig && udal_heads_l0_persistent)) && layer == 0) {
      // fp32 BiFPN features -> bf16 layer-0 output: persistent kernel, depthwise on the CUDA cores, pointwise on tcgen05
      UDAL_TRY(udal_heads_l0_layer(ctx, head, feats, B, p.out));
      mark();
      continue;
    }
    if ((udal_heads_tc_use_ig || fp16) && layer == 1) {
      // sample-invariant input: persistent kernel, one depthwise pass per image, T samples back to back
      const size_t tower_img = (size_t)9 * KF * KF, pred_img = (size_t)9 * npad_p * KF;
      const float* ones = reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(h.ig_w) +
                                                         (size_t)(R - 2) * tower_img + pred_img);
      UDAL_TRY(udal_heads_l1_layer(ctx, p.in, B, mc ? T : 1, p.dw, p.wf, p.fb, mc ? p.scale : nullptr,
                                   mc ? p.out_scale : nullptr, ones,
                                   head == UDAL_HEAD_CLASS ? c.inv_keep_class : c.inv_keep_box, p.out));
      mark();
      continue;
    }
    if (predict && fused_pre) {
      // predict layer + MC moments / decode in one kernel: the [T,...] head outputs never reach HBM
      UDAL_REQUIRE(udal_heads_tc_use_ig && mc && R >= 2 && h.fused_w, "fused predict kernels: configuration not covered");
      const float* fbias = reinterpret_cast<const float*>(reinterpret_cast<const char*>(h.fused_w) + (size_t)9 * h.fused_rows * KF * 2);
      UDAL_TRY(udal_heads_fused_predict(ctx, head, p.in, B, T, h.fused_w, h.fused_rows, fbias, fused_pre));
      mark();
      continue;
    }
    if (use_ig) {
      const size_t tower_img = (size_t)9 * KF * KF, pred_img = (size_t)9 * npad_p * KF;
      const __nv_bfloat16* img0 = reinterpret_cast<const __nv_bfloat16*>(h.ig_w);
      const __nv_bfloat16* img = img0 + (predict ? (size_t)(R - 2) * tower_img : (size_t)(layer - 2) * tower_img);
      const float* ones = reinterpret_cast<const float*>(img0 + (size_t)(R - 2) * tower_img + pred_img);
      const float* ep_scale[UDAL_MAX_LEVELS];
      for (int l = 0; l < L; ++l) ep_scale[l] = predict ? ones : h.bn_scale + ((size_t)layer * L + l) * KF;
      if (predict && h.pred_chunks > 1) {
        for (int q = 0; q < h.pred_chunks; ++q) {
          const int n0 = q * h.pred_chunk, nc = h.cout - n0 < h.pred_chunk ? h.cout - n0 : h.pred_chunk;
          const float* fbq[UDAL_MAX_LEVELS];
          for (int l = 0; l < L; ++l) fbq[l] = p.fb[l] + (size_t)q * KF;
          UDAL_TRY(udal_heads_ig_layer(ctx, p.in, NBt, img + (size_t)q * tower_img, KF, ep_scale, fbq, KF, nc, 1, nullptr, ones,
                                       p.out, n0, h.cout));
        }
        mark();
        continue;
      }
      UDAL_TRY(udal_heads_ig_layer(ctx, p.in, NBt, img, predict ? h.ig_rows : KF, ep_scale, p.fb, p.Npad, p.Cout,
                                   predict ? 1 : 0, mc && !predict ? p.out_scale : nullptr, ones, p.out));
      mark();
      continue;
    }
    if (!udal_heads_tc_use_ig && layer >= 2 && mc) {
      // per-tile kernel for every layer: inputs already carry their dropout, nothing to fold
      for (int l = 0; l < L; ++l) p.scale[l] = nullptr;
    }
    UDAL_REQUIRE(p.Npad <= kMaxN, "predict layer with %d channels: the per-tile kernel handles at most %d", h.cout, kMaxN);
    dim3 grid(total_tiles, grid_y);
    if (p.Npad == 64) sepconv_tc_kernel<64><<<grid, kThreads, SM_ALLOC, ctx->stream>>>(p);
    else sepconv_tc_kernel<80><<<grid, kThreads, SM_ALLOC, ctx->stream>>>(p);
    UDAL_CHECK_LAUNCH(ctx);
    mark();
  }
  return UDAL_OK;
}

int udal_heads_tc_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                         float* const* box_out, const udal_prenms_out* fused_pre) {
  if (udal_heads_wide_ok(ctx)) {
    UDAL_REQUIRE(!fused_pre, "the fused predict + decode kernels are built for fpn_num_filters = 64");
    return udal_heads_wide_sample(ctx, feats, batch, scale, cls_out, box_out);
  }
  for (int l = 0; l < ctx->cfg.num_levels; ++l)
    UDAL_REQUIRE(((uintptr_t)feats[l] & 15) == 0 &&
                     (fused_pre || (((uintptr_t)cls_out[l] & 15) == 0 && ((uintptr_t)box_out[l] & 15) == 0)),
                 "level %d: feature / output pointers must be 16-byte aligned", l);
  UDAL_TRY(run_tower_tc(ctx, UDAL_HEAD_CLASS, feats, batch, scale, fused_pre ? nullptr : cls_out, fused_pre));
  if (fused_pre && ctx->between_heads) UDAL_TRY(ctx->between_heads(ctx, ctx->between_heads_arg));
  UDAL_TRY(run_tower_tc(ctx, UDAL_HEAD_BOX, feats, batch, scale, fused_pre ? nullptr : box_out, fused_pre));
  return UDAL_OK;
}
