// K1, tower layers >= 2 and the predict layers (bf16 tensor-core mode): the separable 3x3 conv as ONE
// implicit GEMM on tcgen05 - the depthwise stage is folded into the dense contraction,
//
//     out[m][n] = sum_tap sum_k  in[pix(m) + tap][k] * ( dw[tap][k] * pw[k][n] * bn_scale[n] )
//
// so the CUDA cores do no convolution arithmetic at all.  Persistent, warp specialised, one CTA
// per SM:
//   warp 0  producer : TMA (cp.async.bulk.tensor.4d, 128B swizzle, zero OOB fill = SAME padding)
//                      loads the 18x16-pixel halo tile of the next work item into a 2-stage ring;
//                      loads the 9 x [N x 64] weight image once (cp.async.bulk), resident after that
//   warp 1  MMA      : one lane issues 36 x tcgen05.mma (M128, N = 64|80, K16) per tile; the A
//                      operand of tap (dy,dx) is the SAME shared-memory tile addressed through a
//                      UMMA descriptor whose start is shifted by (dy*16 + dx) pixel rows (row pitch
//                      16 px = 2048 B, so every 8-row group keeps the same swizzle phase);
//                      accumulators double buffered in TMEM; tcgen05.commit releases the smem
//                      stage and publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld -> folded bias -> swish (tanh.approx) -> this layer's
//                      SpatialDropout2D keep-scale -> bf16 store (or fp32 predictions)
// Work item = (16x8-pixel tile, (sample,image)); items are strided over the CTAs.
//
// Reference arithmetic replaced: efficientdet_keras.py:448-483 / 628-664 (_conv_bn_act and the
// predict SeparableConv2D) for repeats >= 2; numerics as heads_tc.cu (bf16 operands, fp32 accum).
#include <cuda.h>
#include <cuda_bf16.h>

#include "udal_common.cuh"

namespace {

constexpr int IG_TH = 16, IG_TW = 8;               // output tile: 16 rows x 8 px = 128 GEMM rows
constexpr int IG_ROWS = IG_TH + 2, IG_PITCH = 16;  // staged halo tile: 18 rows, row pitch 16 px
constexpr int IG_STAGE_BYTES = IG_ROWS * IG_PITCH * 128;  // 36 864 (smem footprint of a stage)
constexpr int IG_BOXW = IG_TW + 2;                        // pixels actually loaded per row
constexpr int KF = 64;
constexpr int kIgStages = 3;          // TMA ring depth
constexpr int kIgThreads = 320;       // producer warp, MMA warp, 2 x 4 epilogue warps

struct IgParams {
  int num_levels, NB, items;             // items = sum_l tiles[l] * NB
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];     // prefix of tiles[l] * NB (items are level major)
  void* out[UDAL_MAX_LEVELS];            // [NB,H,W,64] bf16 or [NB,H,W,Cout] fp32
  const float* out_scale[UDAL_MAX_LEVELS];  // [NB,64] keep-scale of THIS layer's dropout, or null
  const float* ep_scale[UDAL_MAX_LEVELS];   // [NPAD] per-level BN scale (1 for the predict layer)
  const float* ep_bias[UDAL_MAX_LEVELS];    // [NPAD] folded bias
  const void* wimg;                      // bf16 [9][NPAD][64] pre-swizzled smem image (level independent)
  int Cout, act, out_fp32;
};

struct IgMaps {
  CUtensorMap m[UDAL_MAX_LEVELS];
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t ig_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void ig_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void ig_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ig_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint32_t ig_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ig_swish(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);  // x * sigmoid(x) = h * tanh(h) + h
}

struct IgItem {
  int l, nb, ty0, tx0;
};
__device__ __forceinline__ IgItem ig_item(const IgParams& p, int item) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < UDAL_MAX_LEVELS; ++i)
    if (i < p.num_levels && item >= p.item_off[i]) l = i;
  const int r = item - p.item_off[l];
  const int nb = r / p.tiles[l], tile = r - nb * p.tiles[l];
  IgItem it;
  it.l = l;
  it.nb = nb;
  it.ty0 = (tile / p.tiles_x[l]) * IG_TH;
  it.tx0 = (tile % p.tiles_x[l]) * IG_TW;
  return it;
}

template <int NPAD>
__global__ void __launch_bounds__(kIgThreads, 1) heads_ig_kernel(const __grid_constant__ IgMaps maps,
                                                                 const IgParams p) {
  constexpr int B_BYTES = 9 * NPAD * 128;
  constexpr int SM_B = 0;
  constexpr int SM_IN = SM_B + B_BYTES;                 // multiple of 1024 for NPAD = 64 | 80
  constexpr int SM_BAR = SM_IN + kIgStages * IG_STAGE_BYTES;  // barriers + tmem slot
  constexpr int SM_FBS = SM_BAR + 128;                  // per level: BN scale [NPAD] then folded bias [NPAD], fp32
  constexpr uint32_t kTmemCols = NPAD <= 64 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: full[3] @0  empty[3] @24  tfull[2] @48  tempty[2] @64  bfull @80  tmem slot @88
  const uint32_t bar0 = sb + SM_BAR;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 24, bar_tfull = bar0 + 48, bar_tempty = bar0 + 64, bar_b = bar0 + 80;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 88);
  float* sFb = reinterpret_cast<float*>(smem + SM_FBS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kIgStages; ++i) {
      bar_init(bar_full + 8 * i, 1);
      bar_init(bar_empty + 8 * i, 1);
    }
    bar_init(bar_tfull, 1);
    bar_init(bar_tfull + 8, 1);
    bar_init(bar_tempty, 128);
    bar_init(bar_tempty + 8, 128);
    bar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + SM_BAR + 88),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = threadIdx.x; e < p.num_levels * NPAD; e += kIgThreads) {
    const int l = e / NPAD, n = e - l * NPAD;
    sFb[(2 * l) * NPAD + n] = __ldg(p.ep_scale[l] + n);
    sFb[(2 * l + 1) * NPAD + n] = __ldg(p.ep_bias[l] + n);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int G = gridDim.x;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      bar_expect_tx(bar_b, B_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       sb + SM_B),
                   "l"(p.wimg), "r"(B_BYTES), "r"(bar_b)
                   : "memory");
      int it = 0, s = 0, ph = 0;
      for (int item = blockIdx.x; item < p.items; item += G, ++it) {
        bar_wait(bar_empty + 8 * s, ph ^ 1);  // stage free (first round passes immediately)
        const IgItem w = ig_item(p, item);
        // 18 row boxes of 10 pixels (the halo tile) into row slots of pitch 16 pixels: the 2 KB pitch
        // keeps every 8-row UMMA group on the same swizzle phase, the 10-pixel boxes keep L2 traffic
        // at 1.4x (instead of 2.25x) of the useful bytes
        bar_expect_tx(bar_full + 8 * s, IG_ROWS * IG_BOXW * 128);
        const uint32_t dst = sb + SM_IN + s * IG_STAGE_BYTES;
#pragma unroll 1
        for (int r = 0; r < IG_ROWS; ++r)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(dst + r * IG_PITCH * 128), "l"(&maps.m[w.l]), "r"(bar_full + 8 * s), "r"(0), "r"(w.tx0 - 1),
              "r"(w.ty0 - 1 + r), "r"(w.nb)
              : "memory");
        if (++s == kIgStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
      bar_wait(bar_b, 0);  // weights resident
      int it = 0, s = 0, ph = 0;
      for (int item = blockIdx.x; item < p.items; item += G, ++it) {
        const int a = it & 1;
        bar_wait(bar_tempty + 8 * a, ((it >> 1) & 1) ^ 1);  // accumulator drained by its epilogue group
        bar_wait(bar_full + 8 * s, ph);                      // halo tile landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t in0 = sb + SM_IN + s * IG_STAGE_BYTES;
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * NPAD);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap % 3;
          // the swizzle XOR is a function of the absolute shared-memory address bits, so a start
          // shifted by whole 128-byte rows needs no base offset (verified on hardware)
          const uint64_t adesc = ig_desc(in0 + (uint32_t)((dy * IG_PITCH + dx) * 128), IG_PITCH * 128, 0);
          const uint64_t bdesc = ig_desc(sb + SM_B + tap * NPAD * 128, 1024, 0);
#pragma unroll
          for (int k = 0; k < KF / 16; ++k)
            ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
        }
        ig_commit(bar_empty + 8 * s);   // smem stage reusable once these MMAs retire
        ig_commit(bar_tfull + 8 * a);   // accumulator ready
        if (++s == kIgStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: 2 groups x 4 warps, group g drains accumulator g =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // GEMM row = pixel (m / 8, m % 8) of the tile
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += G, ++it) {
      if ((it & 1) != g) continue;
      const int a = g;
      const IgItem w = ig_item(p, item);
      const int nb = w.nb, H = p.H[w.l], W = p.W[w.l];
      const int oy = w.ty0 + (m >> 3), ox = w.tx0 + (m & 7);
      const bool ok = oy < H && ox < W;
      const size_t pix = ((size_t)nb * H + oy) * W + ox;
      const float* ep_s = sFb + (2 * w.l) * NPAD;
      const float* ep_b = ep_s + NPAD;
      const float* osc = p.out_scale[w.l];
      void* outp = p.out[w.l];
      // this item's dropout keep-scales: fetched while the MMAs are still running
      float4 scv[KF / 4];
      const bool has_sc = osc != nullptr && !p.out_fp32;
      if (has_sc) {
        const float4* sc = reinterpret_cast<const float4*>(osc + (size_t)nb * KF);
#pragma unroll
        for (int i = 0; i < KF / 4; ++i) scv[i] = __ldg(sc + i);
      }
      bar_wait(bar_tfull + 8 * a, (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NPAD);
      uint32_t r[NPAD / 8][8];
#pragma unroll
      for (int j = 0; j < NPAD / 8; ++j) ig_ld8(taddr + j * 8, r[j]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      bar_arrive(bar_tempty + 8 * a);  // accumulator may be overwritten
      if (ok) {
#pragma unroll
        for (int j = 0; j < NPAD / 8; ++j) {
          const int n0 = j * 8;
          float v[8];
          const float4 f0 = *reinterpret_cast<const float4*>(ep_b + n0);
          const float4 f1 = *reinterpret_cast<const float4*>(ep_b + n0 + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(ep_s + n0);
          const float4 g1 = *reinterpret_cast<const float4*>(ep_s + n0 + 4);
          const float fbv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          const float gsv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = fmaf(__uint_as_float(r[j][i]), gsv[i], fbv[i]);  // BN scale + folded bias
            if (p.act) v[i] = ig_swish(v[i]);
          }
          if (!p.out_fp32) {
            if (j < KF / 8) {
              if (has_sc) {
                const float4 s0 = scv[2 * (j < KF / 8 ? j : 0)], s1 = scv[2 * (j < KF / 8 ? j : 0) + 1];
                v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
                v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
              }
              uint4 o;
              o.x = ig_pack(v[0], v[1]);
              o.y = ig_pack(v[2], v[3]);
              o.z = ig_pack(v[4], v[5]);
              o.w = ig_pack(v[6], v[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(outp) + pix * KF + n0) = o;
            }
          } else {
            float* dst = reinterpret_cast<float*>(outp) + pix * p.Cout + n0;
            if ((p.Cout & 3) == 0 && n0 + 8 <= p.Cout) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (n0 + i < p.Cout) dst[i] = v[i];
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// weight image: wimg[tap][n][k] = bf16( dw[tap][k] * wf[n][k] ) in the swizzled shared-memory layout
__global__ void build_ig_weights_kernel(const float* __restrict__ dw, const float* __restrict__ wf, int npad,
                                        __nv_bfloat16* __restrict__ wimg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * npad * KF) return;
  const int k = i % KF, n = (i / KF) % npad, tap = i / (KF * npad);
  const float v = dw[tap * KF + k] * wf[(size_t)n * KF + k];
  const size_t byte = (size_t)tap * npad * 128 + (size_t)n * 128 + (size_t)((((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
  wimg[byte / 2] = __float2bfloat16_rn(v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

// builds the swizzled weight image of one (layer, level): out must hold 9*npad*64 bf16
int udal_heads_ig_build_weights(udal_ctx* ctx, const float* dw, const float* wf, int npad, void* wimg) {
  const int total = 9 * npad * KF;
  build_ig_weights_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(dw, wf, npad, reinterpret_cast<__nv_bfloat16*>(wimg));
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// one tower (>= 2) or predict layer over ALL pyramid levels: in[l] [NB,H_l,W_l,64] bf16 (dropout already applied)
int udal_heads_ig_layer(udal_ctx* ctx, const void* const* in, int NB, const void* wimg, const float* const* ep_scale,
                        const float* const* ep_bias, int npad, int cout, int act, int out_fp32,
                        const float* const* out_scale, void* const* out) {
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const udal_config& c = ctx->cfg;
  IgMaps maps;
  IgParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.num_levels = c.num_levels;
  p.NB = NB;
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    const cuuint64_t gdim[4] = {(cuuint64_t)KF, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    const cuuint64_t gstr[3] = {(cuuint64_t)KF * 2, (cuuint64_t)W * KF * 2, (cuuint64_t)H * W * KF * 2};
    const cuuint32_t box[4] = {(cuuint32_t)KF, (cuuint32_t)IG_BOXW, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = encode(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in[l]), gdim, gstr, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UDAL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,64]", (int)r, NB, H, W);
    p.H[l] = H;
    p.W[l] = W;
    p.tiles_x[l] = (W + IG_TW - 1) / IG_TW;
    p.tiles[l] = p.tiles_x[l] * ((H + IG_TH - 1) / IG_TH);
    p.item_off[l] = off;
    off += p.tiles[l] * NB;
    p.out[l] = out[l];
    p.out_scale[l] = out_scale ? out_scale[l] : nullptr;
    p.ep_scale[l] = ep_scale[l];
    p.ep_bias[l] = ep_bias[l];
  }
  for (int l = c.num_levels; l <= UDAL_MAX_LEVELS; ++l) p.item_off[l] = off;
  p.items = off;
  p.Cout = cout;
  p.act = act;
  p.out_fp32 = out_fp32;
  p.wimg = wimg;
  const int grid = p.items < UDAL_NUM_SMS ? p.items : UDAL_NUM_SMS;
  if (npad == 64) {
    constexpr int smem = 9 * 64 * 128 + kIgStages * IG_STAGE_BYTES + 128 + UDAL_MAX_LEVELS * 2 * 64 * 4 + 1024;
    UDAL_CUDA(cudaFuncSetAttribute(heads_ig_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    heads_ig_kernel<64><<<grid, kIgThreads, smem, ctx->stream>>>(maps, p);
  } else {
    constexpr int smem = 9 * 80 * 128 + kIgStages * IG_STAGE_BYTES + 128 + UDAL_MAX_LEVELS * 2 * 80 * 4 + 1024;
    UDAL_CUDA(cudaFuncSetAttribute(heads_ig_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    heads_ig_kernel<80><<<grid, kIgThreads, smem, ctx->stream>>>(maps, p);
  }
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
