// K1 for head widths other than 64 (bf16 tensor-core mode): fpn_num_filters in (64, 128] - EfficientDet-D1 (88) and
// D2 (112, BASELINE configs[4]) - with the channels zero-padded to 128.
//
// The implicit-GEMM kernels of heads_ig.cu keep a 9-tap weight image resident (9 x N x K bf16); at K = N = 128
// that image is 288 KB and no longer fits an SM, so this kernel splits the separable conv the classic way:
//
//   warp 0      producer : TMA box (18 x 10 px halo tile x 128 ch bf16, zero OOB fill = SAME padding), 2-stage ring;
//                          loads the [128 n][128 k] pointwise weight image once (32 KB, resident)
//   warps 2-9   builders : depthwise 3x3 on the CUDA cores (fp32 accumulate; thread = 4 channels x 1 tile column,
//                          sliding down the rows), times the SpatialDropout2D keep-scale of the producer layer (it
//                          commutes with the depthwise conv) -> A operand [128 px][128 ch] bf16 in the K-major
//                          128B-swizzled UMMA layout (two 64-channel atoms), double buffered
//   warp 1      MMA      : 8 x tcgen05.mma (M128 N128 K16) per tile, accumulators double buffered in TMEM
//   warps 10-17 epilogue : lane quarter x column half: tcgen05.ld -> BN scale / folded bias -> swish -> bf16 staging
//                          tile -> 2 TMA tensor stores (tower layers), or + bias -> fp32 staging -> coalesced rows of
//                          32 channels (predict layers, any channel count; > 128 channels run as chunks)
// Work item = (16x8-pixel tile, (sample, image)), strided over one CTA per SM.
//
// Reference arithmetic replaced: efficientdet_keras.py:448-483 / 628-664 (_conv_bn_act, the predict SeparableConv2D)
// and the MC loop 979-1050; numerics as the 64-channel path: bf16 operands (depthwise output rounded once), fp32
// accumulation, BN applied to the fp32 accumulator.
#include <cuda.h>
#include <cuda_bf16.h>

#include "udal_common.cuh"
#include "heads_umma.cuh"

namespace {

constexpr int kWidePad = 128;                             // padded channel count of the wide heads
constexpr int WF = kWidePad;                              // (host code and helper kernels; the head kernel has its own CH)
constexpr int kWdBuilderWarps = 8, kWdEpiWarps = 8;
constexpr int kWdThreads = 64 + 32 * (kWdBuilderWarps + kWdEpiWarps);
constexpr int WD_STG_STRIDE = 34;                         // floats per pixel row of the fp32 staging tile (32 + pad, even)
// compile-time shape of one kernel variant: CH channels (64 | 128, in = out), input bf16 or fp32 (layer 0 of the
// 64-channel towers reads the BiFPN features directly).  Shared memory map, offsets from a 1024-byte aligned base.
// X3 (fp32-accurate mode, 64 channels): both GEMM operands as fp16 pairs hi + lo (x = hi + lo to ~2^-22), the product
// as three tensor-core passes into one fp32 accumulator: A_hi B_hi + A_hi B_lo + A_lo B_hi (the dropped lo x lo term is
// ~2^-22 relative); activations stay fp32 in HBM.
template <int CH_, bool F32IN_, bool X3_ = false>
struct WdShape {
  static constexpr int CH = CH_, ATOMS = CH_ / 64;
  static constexpr bool F32IN = F32IN_, X3 = X3_;
  static constexpr int PX_BYTES = CH * (F32IN ? 4 : 2);
  static constexpr int STAGE = IG_ROWS * IG_BOXW * PX_BYTES;   // halo tile, linear [18][10][CH]
  static constexpr int A_TILE = ATOMS * 16384;                 // ATOMS x [128 px][128 B] depthwise output (UMMA A)
  static constexpr int A_BYTES = A_TILE * (X3 ? 2 : 1);        // X3: hi tile, lo tile
  static constexpr int B_IMG = ATOMS * CH * 128;               // ATOMS x [CH n][128 B] pointwise weights (UMMA B)
  static constexpr int B_BYTES = B_IMG * (X3 ? 2 : 1);         // X3: hi image, lo image
  static constexpr int A = 0;                                  // 2 x A_BYTES (double buffered)
  static constexpr int B = A + 2 * A_BYTES;
  static constexpr int OUT = B + B_BYTES;                      // staging: bf16 ATOMS x [128][128 B], or fp32 2 x [128][34]
  static constexpr int OUT_BYTES = 2 * 128 * WD_STG_STRIDE * 4;
  static constexpr int STAGES = (CH == 64 && !X3) ? 3 : 2;     // halo tile ring (HBM latency; 128 channels / X3: no room for a third)
  static_assert(!X3 || (CH == 64 && F32IN), "the fp32-accurate mode: 64 channels, fp32 activations");
  static constexpr int IN = OUT + OUT_BYTES;                   // STAGES x halo tile
  static constexpr int BAR = IN + STAGES * STAGE;              // mbarriers + tmem slot
  static constexpr int EP = BAR + 256;                         // [2][CH] fp32 epilogue scale | bias of the current item's level
  static constexpr int QUEUE = EP + 2 * CH * 4;                // item-index ring (IG_QRING ints)
  static constexpr int SMEM = QUEUE + IG_QRING * 4 + 1024;
  static_assert(CH == 64 || CH == 128, "channel count");
  static_assert(STAGE % 128 == 0 && IN % 1024 == 0 && B % 1024 == 0 && OUT % 1024 == 0, "alignment");   // (the halo tile is linear: no swizzle atom)
  static_assert(A_TILE <= OUT_BYTES, "bf16 staging tile");
  static_assert(SMEM <= kIgSmemLimit, "shared-memory budget");
};

struct WdParams {
  int num_levels, NB, items;             // NB = (sample, image) pairs written; items = sum_l tiles[l] * NB (level major)
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];
  int in_nb;                             // images of the input tensor: input image of pair nb = nb % in_nb
  int F;                                 // real channel count (row stride of in_scale)
  const float* in_scale[UDAL_MAX_LEVELS];   // [NB][F] keep-scale of the producer layer's dropout, or null
  const float* ep[UDAL_MAX_LEVELS];      // [2][128]: epilogue scale (BN scale | 1), folded bias
  const float* dw;                       // [9][128] depthwise weights (zero past F)
  const void* wimg;                      // bf16 2 atoms x [128 n][64 k], 128B swizzle
  void* out[UDAL_MAX_LEVELS];            // tower: [NB,H,W,128] bf16 (through the tensor maps); predict: [NB,H,W,ch_total] fp32
  int predict, Cout, ch_off, ch_total;   // predict: this launch writes channels [ch_off, ch_off + Cout)
  int act;                               // fp32-out path (predict = 1): 1 = epilogue scale, bias and an accurate swish (X3 tower layers); 2 = scale and bias (BiFPN)
  int debug;                             // timing experiments only (wrong results): 1 = no depthwise math, 2 = no epilogue math / stores
  int* counter;                          // zeroed work-item counter of this launch (dynamic claiming, heads_umma.cuh)
};

struct WdMaps {
  CUtensorMap in[UDAL_MAX_LEVELS];    // [in_nb,H,W,128] bf16, box {128,10,18,1}, no swizzle
  CUtensorMap out[UDAL_MAX_LEVELS];   // [NB,H,W,128] bf16, box {64,8,16,1}, 128B swizzle
};

__device__ __forceinline__ void wd_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void wd_half_sync(int hc) { asm volatile("bar.sync %0, 128;" ::"r"(2 + hc) : "memory"); }

__device__ __forceinline__ float wd_swish_accurate(float x) {  // x * sigmoid(x), ~4e-7 relative (ex2 / rcp approximations)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

template <int CH, bool F32IN, bool FP16, bool X3 = false>
__global__ void __launch_bounds__(kWdThreads, 1) heads_wide_kernel(const __grid_constant__ WdMaps maps, const WdParams p) {
  using S = WdShape<CH, F32IN, X3>;
  constexpr int WF = CH, WD_STAGE = S::STAGE, WD_A = S::A, WD_B = S::B, WD_OUT = S::OUT, WD_IN = S::IN, WD_BAR = S::BAR,
                WD_EP = S::EP, A_BYTES = S::A_BYTES, B_BYTES = S::B_BYTES, PX_BYTES = S::PX_BYTES;
  constexpr uint32_t kTmemCols = 2 * CH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: in_full[4] @0  in_empty[4] @32  a_full[2] @64  a_empty[2] @80  tfull[2] @96  tempty[2] @112  bfull @128  slot @136
  constexpr int STAGES = S::STAGES;
  const uint32_t bar0 = sb + WD_BAR;
  const uint32_t in_full = bar0, in_empty = bar0 + 32, a_full = bar0 + 64, a_empty = bar0 + 80, tfull = bar0 + 96,
                 tempty = bar0 + 112, bar_b = bar0 + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + WD_BAR + 136);
  volatile int* sQ = reinterpret_cast<volatile int*>(smem + S::QUEUE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(in_full + 8 * i, 1);
      bar_init(in_empty + 8 * i, kWdBuilderWarps);  // one arrival per builder warp
    }
    for (int i = 0; i < 2; ++i) {
      bar_init(a_full + 8 * i, kWdBuilderWarps);
      bar_init(a_empty + 8 * i, 1);
      bar_init(tfull + 8 * i, 1);
      bar_init(tempty + 8 * i, kWdEpiWarps);        // one arrival per epilogue warp
    }
    bar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + WD_BAR + 136), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer (warp-uniform loop, one elected lane issues) =====================
    if (ig_elect_one()) {
      bar_expect_tx(bar_b, B_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + WD_B),
                   "l"(p.wimg), "r"(B_BYTES), "r"(bar_b)
                   : "memory");
    }
    __syncwarp();
    int s = 0, ph = 0;
    for (int i = 0;; ++i) {
      const int item = ig_claim(p.counter, p.items, lane);
      const IgItem w = ig_item(p, item < 0 ? 0 : item);
      const int nb_in = w.nb % p.in_nb;
      if (ig_elect_one()) {
        bar_wait(in_empty + 8 * s, ph ^ 1);
        sQ[i & (IG_QRING - 1)] = item;  // published by the arrival on the stage's full barrier; -1 ends the stream
        if (item < 0) {
          bar_arrive(in_full + 8 * s);
        } else {
          bar_expect_tx(in_full + 8 * s, WD_STAGE);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(sb + WD_IN + s * WD_STAGE), "l"(&maps.in[w.l]), "r"(in_full + 8 * s), "r"(0), "r"(w.tx0 - 1), "r"(w.ty0 - 1),
              "r"(nb_in)
              : "memory");
        }
      }
      __syncwarp();
      if (item < 0) break;
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ig_idesc<FP16>(WF);
    if (lane == 0) bar_wait(bar_b, 0);  // weights resident
    __syncwarp();
    for (int i = 0;; ++i) {
      const int ab = i & 1;
      const uint32_t d_tmem = tmem_base + (uint32_t)(ab * WF);
      if (lane == 0) bar_wait(a_full + 8 * ab, (i >> 1) & 1);  // depthwise output of this item in place (or the end marker)
      __syncwarp();
      if (ig_queue_read(sQ, i) < 0) {
        if (ig_elect_one()) {  // wake the epilogue: its next accumulator "arrives" empty
          bar_wait(tempty + 8 * ab, ((i >> 1) & 1) ^ 1);
          bar_arrive(tfull + 8 * ab);
        }
        __syncwarp();
        break;
      }
      if (ig_elect_one()) {
        bar_wait(tempty + 8 * ab, ((i >> 1) & 1) ^ 1);    // accumulator drained
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (X3) {
          // A_hi B_hi + A_hi B_lo + A_lo B_hi, smallest terms last
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = ig_desc(sb + WD_A + ab * A_BYTES + (uint32_t)((g == 2 ? S::A_TILE : 0) + k * 32), 1024, 0);
              const uint64_t bdesc = ig_desc(sb + WD_B + (uint32_t)((g == 1 ? S::B_IMG : 0) + k * 32), 1024, 0);
              ig_mma(d_tmem, adesc, bdesc, idesc, (g | k) ? 1u : 0u);
            }
        } else {
#pragma unroll
          for (int k = 0; k < WF / 16; ++k) {
            const uint64_t adesc = ig_desc(sb + WD_A + ab * A_BYTES + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32), 1024, 0);
            const uint64_t bdesc = ig_desc(sb + WD_B + (uint32_t)((k >> 2) * (CH * 128) + (k & 3) * 32), 1024, 0);
            ig_mma(d_tmem, adesc, bdesc, idesc, k ? 1u : 0u);
          }
        }
        ig_commit(a_empty + 8 * ab);   // A buffer reusable once these MMAs retire
        ig_commit(tfull + 8 * ab);     // accumulator ready
      }
      __syncwarp();
    }
  } else if (warp < 2 + kWdBuilderWarps) {
    // ===================== builders: depthwise 3x3 -> A operand =====================
    const int tid = threadIdx.x - 64;
    // thread = channel quad (4 q4 .. 4 q4 + 3) x tile column; 128 channels: both row halves of the tile in turn,
    // 64 channels: the two thread halves take one row half each
    constexpr int QUADS = CH / 4;
    const int q4 = tid & (QUADS - 1), x = (tid / QUADS) & 7;
    const int half_lo = CH == 128 ? 0 : tid >> 7, half_hi = CH == 128 ? 2 : half_lo + 1;
    float2 wgt[9][2];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.dw + tp * WF) + q4);
      wgt[tp][0] = make_float2(w4.x, w4.y);
      wgt[tp][1] = make_float2(w4.z, w4.w);
    }
    const bool real = 4 * q4 < p.F;  // F % 4 == 0: a quad is entirely real or entirely padding
    int s = 0, ph = 0;
    for (int i = 0;; ++i) {
      const int ab = i & 1;
      if (lane == 0) {
        bar_wait(in_full + 8 * s, ph);                    // halo tile landed
        bar_wait(a_empty + 8 * ab, ((i >> 1) & 1) ^ 1);   // the MMAs of item i-2 are done with this A buffer
      }
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) {  // end of the stream: pass it on to the MMA warp through the A barrier
        if (lane == 0) bar_arrive(a_full + 8 * ab);
        break;
      }
      const IgItem w = ig_item(p, item);
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.in_scale[w.l] && real) sc = __ldg(reinterpret_cast<const float4*>(p.in_scale[w.l] + (size_t)w.nb * p.F) + q4);
      const uint8_t* sIn = smem + WD_IN + s * WD_STAGE;
      uint8_t* sA = smem + WD_A + ab * A_BYTES + (q4 >> 4) * 16384;
      const uint32_t chunk = (uint32_t)((q4 & 15) >> 1), sub = (uint32_t)(q4 & 1) * 8;
#pragma unroll 1
      for (int half = (p.debug & 1) ? half_hi : half_lo; half < half_hi; ++half) {  // tile rows 8 half .. 8 half + 7 (halo rows 8 half .. 8 half + 9)
        float2 acc[8][2];  // packed pairs: the 9 x 4 FMAs per pixel issue as 18 FFMA2
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[y][0] = acc[y][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint8_t* px = sIn + (size_t)((8 * half + r) * IG_BOXW + x + dx) * PX_BYTES;
            float2 v01, v23;
            if constexpr (F32IN) {
              // BiFPN features enter the fp32 depthwise accumulation unrounded (the A operand is rounded to bf16 once)
              const float4 f = *reinterpret_cast<const float4*>(px + q4 * 16);
              v01 = make_float2(f.x, f.y);
              v23 = make_float2(f.z, f.w);
            } else {
              const uint2 raw2 = *reinterpret_cast<const uint2*>(px + q4 * 8);
              v01 = ig_unpack16<FP16>(raw2.x);
              v23 = ig_unpack16<FP16>(raw2.y);
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int y = r - dy;
              if (y >= 0 && y < 8) {
                ig_ffma2(acc[y][0], v01, wgt[dy * 3 + dx][0]);
                ig_ffma2(acc[y][1], v23, wgt[dy * 3 + dx][1]);
              }
            }
          }
          if (r >= 2) {
            const int y = r - 2;
            const int m = (8 * half + y) * IG_TW + x;
            const float v0 = acc[y][0].x * sc.x, v1 = acc[y][0].y * sc.y, v2 = acc[y][1].x * sc.z, v3 = acc[y][1].y * sc.w;
            uint2 o;
            o.x = ig_pack16<FP16>(v0, v1);
            o.y = ig_pack16<FP16>(v2, v3);
            uint8_t* const dstA = sA + (size_t)m * 128 + ((chunk ^ (uint32_t)(m & 7)) << 4) + sub;
            *reinterpret_cast<uint2*>(dstA) = o;
            if constexpr (X3) {  // the part of the value the 16-bit operand lost, as a second operand tile
              const float2 h01 = ig_unpack16<FP16>(o.x), h23 = ig_unpack16<FP16>(o.y);
              uint2 lo;
              lo.x = ig_pack16<FP16>(v0 - h01.x, v1 - h01.y);
              lo.y = ig_pack16<FP16>(v2 - h23.x, v3 - h23.y);
              *reinterpret_cast<uint2*>(dstA + S::A_TILE) = lo;
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A tile -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        bar_arrive(a_full + 8 * ab);
        bar_arrive(in_empty + 8 * s);
      }
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ===================== epilogue: 4 lane quarters x 2 column halves =====================
    const int ew = warp - 2 - kWdBuilderWarps;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    constexpr int HC = CH / 2, PASSES = HC / 32;  // columns per warp, in passes of 32
    const int hc = ew >> 2;                 // column half: accumulator columns HC hc .. HC hc + HC - 1
    const int m = q * 32 + lane;            // GEMM row = pixel (m / 8, m % 8) of the tile
    const bool elected = ew == 0 && lane == 0;
    const uint32_t swz = (uint32_t)(m & 7);
    uint8_t* const ob = smem + WD_OUT;
    int ep_level = -1;  // level whose epilogue table is in shared memory
    for (int i = 0;; ++i) {
      const int ab = i & 1;
      if (lane == 0) bar_wait(tfull + 8 * ab, (i >> 1) & 1);
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) break;
      const IgItem w = ig_item(p, item);
      // this item's epilogue table -> shared memory (tower: halved, x*sigmoid(x) = h*tanh(h) + h with h = x/2); the
      // previous item's readers passed its last group barrier, the first barrier below publishes the table
      float* const sEp = reinterpret_cast<float*>(smem + WD_EP);
      if (w.l != ep_level) {  // (items arrive level by level: a handful of reloads per CTA)
        ep_level = w.l;
        const int et = threadIdx.x - 64 - 32 * kWdBuilderWarps;
        if (!p.predict) {
          if (et < 2 * CH) sEp[et] = 0.5f * __ldg(p.ep[w.l] + et);
        } else if ((et & 127) < HC) {  // predict: each column-half group (own barrier) loads the biases it reads
          const int n = WF + hc * HC + (et & 127);
          sEp[n] = __ldg(p.ep[w.l] + n);
          if (p.act) sEp[n - WF] = __ldg(p.ep[w.l] + n - WF);
        }
      }
      const float4* eps = reinterpret_cast<const float4*>(sEp + hc * HC);
      const float4* epb = reinterpret_cast<const float4*>(sEp + WF + hc * HC);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * WF + hc * HC);
      if (p.debug & 2) {
        uint32_t r8[8];
        ig_ld8(taddr, r8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(tempty + 8 * ab);
      } else if (!p.predict) {
        // ---- tower layer: BN scale + folded bias -> swish -> bf16 staging tile (atom hc) -> TMA store ----
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free again
        __syncwarp();
        wd_epi_sync();
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
          uint32_t r[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) ig_ld8(taddr + pass * 32 + u * 8, r[u]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (pass == PASSES - 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty + 8 * ab);  // accumulator may be overwritten
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = pass * 4 + u;                 // 8-channel group of this warp's columns (epilogue table index)
            const int col = hc * HC + j * 8;            // first output channel of the group
            const float4 g0 = eps[2 * j], g1 = eps[2 * j + 1];
            const float4 f0 = epb[2 * j], f1 = epb[2 * j + 1];
            const float2 gs[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
            const float2 fb[4] = {make_float2(f0.x, f0.y), make_float2(f0.z, f0.w), make_float2(f1.x, f1.y), make_float2(f1.z, f1.w)};
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // packed pairs: 2 FFMA2 + 2 MUFU per 2 outputs
              const float2 h = ig_fma2(make_float2(__uint_as_float(r[u][2 * e]), __uint_as_float(r[u][2 * e + 1])), gs[e], fb[e]);
              const float2 sw = ig_fma2(h, make_float2(ig_tanh(h.x), ig_tanh(h.y)), h);
              v[2 * e] = sw.x;
              v[2 * e + 1] = sw.y;
            }
            uint4 o;
            o.x = ig_pack16<FP16>(v[0], v[1]);
            o.y = ig_pack16<FP16>(v[2], v[3]);
            o.z = ig_pack16<FP16>(v[4], v[5]);
            o.w = ig_pack16<FP16>(v[6], v[7]);
            *reinterpret_cast<uint4*>(ob + (col >> 6) * 16384 + m * 128 + (((uint32_t)((col & 63) >> 3) ^ swz) << 4)) = o;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to TMA
        wd_epi_sync();
        if (elected) {
          ig_tma_store(&maps.out[w.l], s32(ob), 0, w.tx0, w.ty0, w.nb);
          if (CH == 128) ig_tma_store(&maps.out[w.l], s32(ob) + 16384, 64, w.tx0, w.ty0, w.nb);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      } else {
        // ---- predict layer: + bias -> fp32 staging [128 px][34] of this column half -> rows of 32 channels ----
        const int H = p.H[w.l], W = p.W[w.l];
        float* const stg = reinterpret_cast<float*>(ob) + hc * (128 * WD_STG_STRIDE);
        const int ht = threadIdx.x - 64 - 32 * kWdBuilderWarps - hc * 128;  // thread of the column-half group
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
          const int c0 = hc * HC + pass * 32;  // first accumulator column of the pass
          uint32_t r[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) ig_ld8(taddr + pass * 32 + u * 8, r[u]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (pass == PASSES - 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty + 8 * ab);
          }
          wd_half_sync(hc);  // the previous pass's copy loop is done with the staging tile
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 f0 = epb[pass * 8 + 2 * u], f1 = epb[pass * 8 + 2 * u + 1];
            const float fb[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            if (p.act) {  // tower layer of the fp32-accurate mode: BN scale, folded bias, swish
              const float4 g0 = eps[pass * 8 + 2 * u], g1 = eps[pass * 8 + 2 * u + 1];
              const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const float a0 = fmaf(__uint_as_float(r[u][e]), gs[e], fb[e]), a1 = fmaf(__uint_as_float(r[u][e + 1]), gs[e + 1], fb[e + 1]);
                *reinterpret_cast<float2*>(stg + m * WD_STG_STRIDE + u * 8 + e) =
                    p.act == 1 ? make_float2(wd_swish_accurate(a0), wd_swish_accurate(a1)) : make_float2(a0, a1);  // 2: BN only
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; e += 2)
                *reinterpret_cast<float2*>(stg + m * WD_STG_STRIDE + u * 8 + e) =
                    make_float2(__fadd_rn(__uint_as_float(r[u][e]), fb[e]), __fadd_rn(__uint_as_float(r[u][e + 1]), fb[e + 1]));
            }
          }
          wd_half_sync(hc);
          const int nch = min(32, p.Cout - c0);  // channels of this pass that exist
          if (nch > 0) {
            float* const dst0 = reinterpret_cast<float*>(p.out[w.l]) + (((size_t)w.nb * H + w.ty0) * W + w.tx0) * p.ch_total + p.ch_off + c0;
            const size_t row_stride = (size_t)W * p.ch_total;
            const int wg = ht >> 5;
            if (((p.ch_total | p.ch_off) & 1) == 0) {
              // 8-byte stores: half-warp = one pixel (16 channel pairs), thread keeps its tile column, loop over the rows
              const int col = wg * 2 + (lane >> 4), cp = 2 * (lane & 15);
              if (w.tx0 + col < W && cp < nch) {
                const float* src = stg + col * WD_STG_STRIDE + cp;
                float* dst = dst0 + (size_t)col * p.ch_total + cp;
                const int rows = min(IG_TH, H - w.ty0);
                if (cp + 1 < nch) {
                  for (int row = 0; row < rows; ++row)
                    *reinterpret_cast<float2*>(dst + row * row_stride) = *reinterpret_cast<const float2*>(src + row * (IG_TW * WD_STG_STRIDE));
                } else {
                  for (int row = 0; row < rows; ++row) dst[row * row_stride] = src[row * (IG_TW * WD_STG_STRIDE)];
                }
              }
            } else {
              for (int pp = wg; pp < 128; pp += 4) {  // warp = pixel, lane = channel
                const int oy = w.ty0 + (pp >> 3), ox = w.tx0 + (pp & 7);
                if (oy < H && ox < W && lane < nch) dst0[(size_t)(pp >> 3) * row_stride + (size_t)(pp & 7) * p.ch_total + lane] = stg[pp * WD_STG_STRIDE + lane];
              }
            }
          }
        }
      }
    }
    if (!p.predict && elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// fp32 [n_px][F] features -> bf16 [n_px][128], zero past F
template <bool FP16>
__global__ void wide_convert_kernel(const float* __restrict__ in, size_t n_px, int F, __nv_bfloat16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 output channels
  if (i >= n_px * (WF / 4)) return;
  const size_t px = i / (WF / 4);
  const int c = (int)(i % (WF / 4)) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < F) v = __ldg(reinterpret_cast<const float4*>(in + px * F + c));
  uint2 o;
  o.x = ig_pack16<FP16>(v.x, v.y);
  o.y = ig_pack16<FP16>(v.z, v.w);
  *reinterpret_cast<uint2*>(out + px * WF + c) = o;
}

// weight image of one pointwise matrix: wimg[atom = k / 64][n][k % 64] = bf16(w[k][n0 + n]) (n < cout, k < F), 128B swizzle
template <bool FP16>
__global__ void wide_weights_kernel(const float* __restrict__ w, int F, int ldw, int n0, int cout, __nv_bfloat16* __restrict__ wimg,
                                    int ch = WF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ch * ch) return;
  const int k = i % ch, n = i / ch;
  float v = 0.f;
  if (k < F && n < cout) v = w[(size_t)k * ldw + n0 + n];
  const int kk = k & 63;
  const size_t byte = (size_t)(k >> 6) * ((size_t)ch * 128) + (size_t)n * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2);
  if constexpr (FP16) reinterpret_cast<__half*>(wimg)[byte / 2] = __float2half_rn(v);
  else wimg[byte / 2] = __float2bfloat16_rn(v);
}

// dst[9][128] = src[9][F] zero padded
__global__ void wide_dw_kernel(const float* __restrict__ src, int F, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * WF) return;
  const int c = i % WF, tp = i / WF;
  dst[i] = c < F ? src[tp * F + c] : 0.f;
}

// ep[0][n] = scale, ep[1][n] = bias*scale + shift (tower) or ep = (1, bias) (predict), zero padded
__global__ void wide_ep_kernel(const float* __restrict__ bias, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                               int n0, int cout, float* __restrict__ ep, int ch = WF) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= ch) return;
  float s = 0.f, b = 0.f;
  if (n < cout) {
    s = bn_scale ? bn_scale[n0 + n] : 1.f;
    b = bn_scale ? fmaf(bias[n0 + n], bn_scale[n0 + n], bn_shift[n0 + n]) : bias[n0 + n];
  }
  ep[n] = s;
  ep[ch + n] = b;
}

}  // namespace

int udal_wide_debug = 0;  // timing experiments (WdParams::debug)

int udal_heads_wide_ok(const udal_ctx* ctx) { return ctx->cfg.num_filters > KF && ctx->cfg.num_filters <= WF; }

// builds the padded weight tables of one head: images [R + chunks][2][128][64] bf16, dw [R + 1][9][128],
// ep [R][L][2][128] + [chunks][2][128]
int udal_heads_wide_prepare(udal_ctx* ctx, int head) {
  const udal_config& c = ctx->cfg;
  const bool fp16 = c.heads_mode == UDAL_HEADS_FP16_TC;
  udal_head_weights_dev& h = ctx->heads[head];
  const int F = c.num_filters, R = c.repeats, L = c.num_levels;
  UDAL_REQUIRE(udal_heads_wide_ok(ctx) && F % 4 == 0, "wide tensor-core heads: fpn_num_filters %d not in (64, 128]", F);
  const int chunks = (h.cout + WF - 1) / WF;
  h.wide_chunks = chunks;
  if (h.wide_w) UDAL_CUDA(cudaFree(h.wide_w));
  if (h.wide_f) UDAL_CUDA(cudaFree(h.wide_f));
  h.wide_w = nullptr;
  h.wide_f = nullptr;
  const size_t n_img = (size_t)(R + chunks) * WF * WF;
  const size_t n_f = (size_t)(R + 1) * 9 * WF + ((size_t)R * L + chunks) * 2 * WF;
  UDAL_CUDA(cudaMalloc(&h.wide_w, n_img * 2));
  UDAL_CUDA(cudaMalloc(&h.wide_f, n_f * sizeof(float)));
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(h.wide_w);
  float* dwp = h.wide_f;
  float* ep = h.wide_f + (size_t)(R + 1) * 9 * WF;
  const int tb = 256;
  for (int r = 0; r < R; ++r) {
    if (fp16) wide_weights_kernel<true><<<(WF * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pw + (size_t)r * F * F, F, F, 0, F, img + (size_t)r * WF * WF);
    else wide_weights_kernel<false><<<(WF * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pw + (size_t)r * F * F, F, F, 0, F, img + (size_t)r * WF * WF);
    UDAL_CHECK_LAUNCH(ctx);
    wide_dw_kernel<<<(9 * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.dw + (size_t)r * 9 * F, F, dwp + (size_t)r * 9 * WF);
    UDAL_CHECK_LAUNCH(ctx);
    for (int l = 0; l < L; ++l) {
      const size_t o = (size_t)r * L + l;
      wide_ep_kernel<<<1, WF, 0, ctx->stream>>>(h.bias + (size_t)r * F, h.bn_scale + o * F, h.bn_shift + o * F, 0, F, ep + o * 2 * WF);
      UDAL_CHECK_LAUNCH(ctx);
    }
  }
  wide_dw_kernel<<<(9 * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.dwp, F, dwp + (size_t)R * 9 * WF);
  UDAL_CHECK_LAUNCH(ctx);
  for (int q = 0; q < chunks; ++q) {
    const int n0 = q * WF, nc = h.cout - n0 < WF ? h.cout - n0 : WF;
    if (fp16) wide_weights_kernel<true><<<(WF * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pwp, F, h.cout, n0, nc, img + (size_t)(R + q) * WF * WF);
    else wide_weights_kernel<false><<<(WF * WF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pwp, F, h.cout, n0, nc, img + (size_t)(R + q) * WF * WF);
    UDAL_CHECK_LAUNCH(ctx);
    wide_ep_kernel<<<1, WF, 0, ctx->stream>>>(h.bp, nullptr, nullptr, n0, nc, ep + ((size_t)R * L + q) * 2 * WF);
    UDAL_CHECK_LAUNCH(ctx);
  }
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  return UDAL_OK;
}

// (num_levels, Hs, Ws, F): the maps this launch covers - the context's pyramid for the heads, one map for a BiFPN node
template <int CH, bool F32IN, bool FP16, bool X3 = false>
static int launch_wide_geom(udal_ctx* ctx, int num_levels, const int* Hs, const int* Ws, int F, const void* const* in, int in_nb, int NB,
                            const float* const* in_scale, const float* dw, const void* wimg, const float* const* ep, int predict,
                            int cout, int ch_off, int ch_total, void* const* out, int act = 0) {
  using S = WdShape<CH, F32IN, X3>;
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  struct {
    int num_levels;
    const int *level_h, *level_w;
  } c = {num_levels, Hs, Ws};
  WdMaps maps;
  WdParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.num_levels = c.num_levels;
  p.NB = NB;
  p.in_nb = in_nb;
  p.F = F;
  p.dw = dw;
  p.wimg = wimg;
  p.predict = predict;
  p.Cout = cout;
  p.ch_off = ch_off;
  p.ch_total = ch_total;
  p.act = act;
  p.debug = udal_wide_debug;
  UDAL_TRY(udal_work_counter(ctx, &p.counter));
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    constexpr CUtensorMapDataType dt16 = FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    UDAL_TRY(encode_nhwc(encode, &maps.in[l], F32IN ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, F32IN ? 4 : 2,
                         in[l], in_nb, H, W, CH, CH, IG_BOXW, IG_ROWS, false));
    if (!predict)
      UDAL_TRY(encode_nhwc(encode, &maps.out[l], dt16, 2, out[l], NB, H, W, CH, 64, IG_TW, IG_TH, true));
    p.H[l] = H;
    p.W[l] = W;
    p.tiles_x[l] = (W + IG_TW - 1) / IG_TW;
    p.tiles[l] = p.tiles_x[l] * ((H + IG_TH - 1) / IG_TH);
    p.tiles_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles[l] - 1) / (uint64_t)p.tiles[l]);
    p.tiles_x_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles_x[l] - 1) / (uint64_t)p.tiles_x[l]);
    UDAL_REQUIRE((int64_t)p.tiles[l] * NB * p.tiles[l] < (1ll << 32), "level %d: too many work items for the item decode", l);
    p.item_off[l] = off;
    off += p.tiles[l] * NB;
    p.in_scale[l] = in_scale ? in_scale[l] : nullptr;
    p.ep[l] = ep[l];
    p.out[l] = out[l];
  }
  for (int l = c.num_levels; l <= UDAL_MAX_LEVELS; ++l) p.item_off[l] = off;
  p.items = off;
  const int grid = udal_persistent_grid(ctx, p.items);
  UDAL_CUDA(cudaFuncSetAttribute(heads_wide_kernel<CH, F32IN, FP16, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
  heads_wide_kernel<CH, F32IN, FP16, X3><<<grid, kWdThreads, S::SMEM, ctx->stream>>>(maps, p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

template <int CH, bool F32IN, bool FP16, bool X3 = false>
static int launch_wide_t(udal_ctx* ctx, const void* const* in, int in_nb, int NB, const float* const* in_scale, const float* dw,
                         const void* wimg, const float* const* ep, int predict, int cout, int ch_off, int ch_total,
                         void* const* out, int act = 0) {
  const udal_config& c = ctx->cfg;
  return launch_wide_geom<CH, F32IN, FP16, X3>(ctx, c.num_levels, c.level_h, c.level_w, c.num_filters, in, in_nb, NB, in_scale, dw,
                                               wimg, ep, predict, cout, ch_off, ch_total, out, act);
}

// 16-bit format by the context's heads mode
template <int CH, bool F32IN>
static int launch_wide(udal_ctx* ctx, const void* const* in, int in_nb, int NB, const float* const* in_scale, const float* dw,
                       const void* wimg, const float* const* ep, int predict, int cout, int ch_off, int ch_total,
                       void* const* out) {
  if (ctx->cfg.heads_mode == UDAL_HEADS_FP16_TC)
    return launch_wide_t<CH, F32IN, true>(ctx, in, in_nb, NB, in_scale, dw, wimg, ep, predict, cout, ch_off, ch_total, out);
  return launch_wide_t<CH, F32IN, false>(ctx, in, in_nb, NB, in_scale, dw, wimg, ep, predict, cout, ch_off, ch_total, out);
}

// ---- layer 0 of the 64-channel towers through the same kernel (CH = 64, fp32 BiFPN features in) ----
int udal_heads_l0_persistent = 1;  // 0: layer 0 through the per-tile kernel of heads_tc.cu (debug / comparison)

// pointwise image [64 n][64 k] bf16 and, per level, the epilogue table (BN scale | folded bias) of tower layer 0
int udal_heads_l0_prepare(udal_ctx* ctx, int head) {
  const udal_config& c = ctx->cfg;
  udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(c.num_filters == KF, "layer-0 kernel: 64 channels");
  const int L = c.num_levels;
  if (h.l0_w) UDAL_CUDA(cudaFree(h.l0_w));
  if (h.l0_ep) UDAL_CUDA(cudaFree(h.l0_ep));
  h.l0_w = nullptr;
  h.l0_ep = nullptr;
  UDAL_CUDA(cudaMalloc(&h.l0_w, (size_t)KF * KF * 2));
  UDAL_CUDA(cudaMalloc(&h.l0_ep, (size_t)L * 2 * KF * sizeof(float)));
  if (c.heads_mode == UDAL_HEADS_FP16_TC)
    wide_weights_kernel<true><<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(h.pw, KF, KF, 0, KF, reinterpret_cast<__nv_bfloat16*>(h.l0_w), KF);
  else
    wide_weights_kernel<false><<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(h.pw, KF, KF, 0, KF, reinterpret_cast<__nv_bfloat16*>(h.l0_w), KF);
  UDAL_CHECK_LAUNCH(ctx);
  for (int l = 0; l < L; ++l) {
    wide_ep_kernel<<<1, KF, 0, ctx->stream>>>(h.bias, h.bn_scale + (size_t)l * KF, h.bn_shift + (size_t)l * KF, 0, KF,
                                              h.l0_ep + (size_t)l * 2 * KF, KF);
    UDAL_CHECK_LAUNCH(ctx);
  }
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  return UDAL_OK;
}

// tower layer 0: feats[l] fp32 [B,H_l,W_l,64] -> out[l] bf16 [B,H_l,W_l,64] = swish(BN(sepconv(bf16(feats)))), no dropout applied.
// With udal_set_feature_format(ctx, UDAL_FEAT_F16) the feature maps are fp16 [B,H_l,W_l,64] (what the reference's GPU exports
// produce under mixed_float16, efficientdet_keras.py / hparams_config.py `mixed_precision`): half the bytes over PCIe and HBM.
int udal_heads_l0_layer(udal_ctx* ctx, int head, const float* const* feats, int B, void* const* out) {
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(h.l0_w && h.l0_ep, "layer-0 tables not built");
  const void* in[UDAL_MAX_LEVELS];
  const float* ep[UDAL_MAX_LEVELS];
  for (int l = 0; l < c.num_levels; ++l) {
    in[l] = feats[l];
    ep[l] = h.l0_ep + (size_t)l * 2 * KF;
  }
  if (ctx->feat_f16) {
    UDAL_REQUIRE(c.heads_mode == UDAL_HEADS_FP16_TC, "fp16 feature maps need heads_mode fp16");
    return launch_wide_t<64, false, true>(ctx, in, B, B, nullptr, h.dw, h.l0_w, ep, 0, KF, 0, KF, out);
  }
  return launch_wide<64, true>(ctx, in, B, B, nullptr, h.dw, h.l0_w, ep, 0, KF, 0, KF, out);
}

// one head: R tower layers + the predict layer(s); feats16[l] = bf16 [B,H_l,W_l,128] features (wide_convert_kernel),
// scale_all = the keep-scale table [2][L][R][T*B][F] of heads_fp32.cu, outs[l] fp32 [NBt,H_l,W_l,cout]
static int run_tower_wide(udal_ctx* ctx, int head, const __nv_bfloat16* const* feats16, int batch, const float* scale_all,
                          float* const* outs) {
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  const int F = c.num_filters, R = c.repeats, L = c.num_levels, T = c.mc_samples, B = batch;
  const bool mc = head == UDAL_HEAD_CLASS ? c.cls_mc != 0 : c.box_mc != 0;
  const int NBt = mc ? T * B : B;
  const size_t P = (size_t)ctx->num_pixels;
  __nv_bfloat16 *a0, *pp;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_A, (size_t)B * P * WF * 2, (void**)&a0));
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_B, 2 * (size_t)NBt * P * WF * 2, (void**)&pp));
  const __nv_bfloat16* img = reinterpret_cast<const __nv_bfloat16*>(h.wide_w);
  const float* dwp = h.wide_f;
  const float* ep_all = h.wide_f + (size_t)(R + 1) * 9 * WF;
  auto mark = [&]() {
    if (!ctx->profile_layers) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      cudaEventRecord(e, ctx->stream);
      ctx->layer_events.push_back(e);
    }
  };
  for (int layer = 0; layer <= R; ++layer) {
    const bool predict = layer == R;
    const void* in[UDAL_MAX_LEVELS];
    void* out[UDAL_MAX_LEVELS];
    const float* in_scale[UDAL_MAX_LEVELS];
    const float* ep[UDAL_MAX_LEVELS];
    const int nb_out = layer == 0 ? B : NBt;
    const int in_nb = layer <= 1 ? B : NBt;
    for (int l = 0; l < L; ++l) {
      const size_t lvl = (size_t)ctx->level_pix_off[l] * WF;
      if (layer == 0) in[l] = feats16[l];
      else if (layer == 1) in[l] = a0 + (size_t)B * lvl;
      else in[l] = pp + (size_t)((layer - 1) & 1) * NBt * P * WF + (size_t)NBt * lvl;
      if (predict) out[l] = outs[l];
      else if (layer == 0) out[l] = a0 + (size_t)B * lvl;
      else out[l] = pp + (size_t)(layer & 1) * NBt * P * WF + (size_t)NBt * lvl;
      // SpatialDropout2D of the producer layer (layer - 1), applied to the depthwise output (it commutes)
      in_scale[l] = (mc && layer >= 1) ? scale_all + (((size_t)head * L + l) * R + (layer - 1)) * (size_t)NBt * F : nullptr;
      ep[l] = predict ? nullptr : ep_all + ((size_t)layer * L + l) * 2 * WF;
    }
    mark();
    if (!predict) {
      UDAL_TRY((launch_wide<kWidePad, false>)(ctx, in, in_nb, nb_out, (mc && layer >= 1) ? in_scale : nullptr, dwp + (size_t)layer * 9 * WF,
                           img + (size_t)layer * WF * WF, ep, 0, WF, 0, WF, out));
    } else {
      for (int q = 0; q < h.wide_chunks; ++q) {
        const int n0 = q * WF, nc = h.cout - n0 < WF ? h.cout - n0 : WF;
        for (int l = 0; l < L; ++l) ep[l] = ep_all + ((size_t)R * L + q) * 2 * WF;
        UDAL_TRY((launch_wide<kWidePad, false>)(ctx, in, in_nb, nb_out, (mc && layer >= 1) ? in_scale : nullptr, dwp + (size_t)R * 9 * WF,
                             img + (size_t)(R + q) * WF * WF, ep, 1, nc, n0, h.cout, out));
      }
    }
    mark();
  }
  return UDAL_OK;
}

int udal_heads_wide_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                           float* const* box_out) {
  const udal_config& c = ctx->cfg;
  const int L = c.num_levels, F = c.num_filters;
  const size_t P = (size_t)ctx->num_pixels;
  __nv_bfloat16* f16;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_D, (size_t)batch * P * WF * 2, (void**)&f16));
  const __nv_bfloat16* feats16[UDAL_MAX_LEVELS];
  for (int l = 0; l < L; ++l) {
    UDAL_REQUIRE(((uintptr_t)feats[l] & 15) == 0, "level %d: feature pointers must be 16-byte aligned", l);
    __nv_bfloat16* dst = f16 + (size_t)batch * ctx->level_pix_off[l] * WF;
    const size_t n_px = (size_t)batch * c.level_h[l] * c.level_w[l];
    const size_t nthr = n_px * (WF / 4);
    if (c.heads_mode == UDAL_HEADS_FP16_TC) wide_convert_kernel<true><<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(feats[l], n_px, F, dst);
    else wide_convert_kernel<false><<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(feats[l], n_px, F, dst);
    UDAL_CHECK_LAUNCH(ctx);
    feats16[l] = dst;
  }
  UDAL_TRY(run_tower_wide(ctx, UDAL_HEAD_CLASS, feats16, batch, scale, cls_out));
  UDAL_TRY(run_tower_wide(ctx, UDAL_HEAD_BOX, feats16, batch, scale, box_out));
  return UDAL_OK;
}

// =====================================================================================================================
// fp32-accurate tensor-core mode (UDAL_HEADS_FP32X3_TC; 64-channel towers): every layer through heads_wide_kernel<64, fp32
// in, fp16 operands, X3> - depthwise in fp32 on the CUDA cores, the pointwise GEMM as three fp16 tensor-core passes over
// (hi, lo) operand pairs, fp32 accumulation, accurate swish, fp32 activations in HBM.  The 1e-4 parity contract of the
// CUDA-core fp32 towers (heads_fp32.cu) at a multiple of their speed; predict layers write the [T,...] outputs, K2 is the
// stand-alone decode_moments kernel.
// =====================================================================================================================
namespace {

// hi / lo fp16 images of one pointwise matrix: [2][64 n][64 k], 128B swizzle; lo = fp16(w - fp32(fp16(w)))
__global__ void x3_weights_kernel(const float* __restrict__ w, int ldw, int n0, int cout, __half* __restrict__ wimg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KF * KF) return;
  const int k = i % KF, n = i / KF;
  const float v = n < cout ? w[(size_t)k * ldw + n0 + n] : 0.f;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const size_t e = ((size_t)n * 128 + (size_t)((((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2)) / 2;
  wimg[e] = hi;
  wimg[(size_t)KF * KF + e] = lo;
}

}  // namespace

int udal_heads_x3_prepare(udal_ctx* ctx, int head) {
  const udal_config& c = ctx->cfg;
  udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(c.num_filters == KF, "heads_mode fp32x3: fpn_num_filters 64 (EfficientDet-D0) only, got %d", c.num_filters);
  const int R = c.repeats, L = c.num_levels;
  const int chunks = (h.cout + KF - 1) / KF;
  h.x3_chunks = chunks;
  if (h.x3_w) UDAL_CUDA(cudaFree(h.x3_w));
  if (h.x3_f) UDAL_CUDA(cudaFree(h.x3_f));
  h.x3_w = nullptr;
  h.x3_f = nullptr;
  UDAL_CUDA(cudaMalloc(&h.x3_w, (size_t)(R + chunks) * 2 * KF * KF * sizeof(__half)));
  UDAL_CUDA(cudaMalloc(&h.x3_f, ((size_t)R * L + chunks) * 2 * KF * sizeof(float)));
  __half* img = reinterpret_cast<__half*>(h.x3_w);
  for (int r = 0; r < R; ++r) {
    x3_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(h.pw + (size_t)r * KF * KF, KF, 0, KF, img + (size_t)r * 2 * KF * KF);
    UDAL_CHECK_LAUNCH(ctx);
    for (int l = 0; l < L; ++l) {
      const size_t o = (size_t)r * L + l;
      wide_ep_kernel<<<1, KF, 0, ctx->stream>>>(h.bias + (size_t)r * KF, h.bn_scale + o * KF, h.bn_shift + o * KF, 0, KF, h.x3_f + o * 2 * KF, KF);
      UDAL_CHECK_LAUNCH(ctx);
    }
  }
  for (int q = 0; q < chunks; ++q) {
    const int n0 = q * KF, nc = h.cout - n0 < KF ? h.cout - n0 : KF;
    x3_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(h.pwp, h.cout, n0, nc, img + (size_t)(R + q) * 2 * KF * KF);
    UDAL_CHECK_LAUNCH(ctx);
    wide_ep_kernel<<<1, KF, 0, ctx->stream>>>(h.bp, nullptr, nullptr, n0, nc, h.x3_f + ((size_t)R * L + q) * 2 * KF, KF);
    UDAL_CHECK_LAUNCH(ctx);
  }
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  return UDAL_OK;
}

static int run_tower_x3(udal_ctx* ctx, int head, const float* const* feats, int batch, const float* scale_all, float* const* outs) {
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(h.x3_w && h.x3_f, "fp32x3 tables not built");
  const int R = c.repeats, L = c.num_levels, T = c.mc_samples, B = batch;
  const bool mc = head == UDAL_HEAD_CLASS ? c.cls_mc != 0 : c.box_mc != 0;
  const int NBt = mc ? T * B : B;
  const size_t P = (size_t)ctx->num_pixels;
  float *a0, *pp;
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_A, (size_t)B * P * KF * 4, (void**)&a0));
  UDAL_TRY(udal_scratch_get(ctx, SCR_HEADS_B, 2 * (size_t)NBt * P * KF * 4, (void**)&pp));
  const __half* img = reinterpret_cast<const __half*>(h.x3_w);
  auto mark = [&]() {
    if (!ctx->profile_layers) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      cudaEventRecord(e, ctx->stream);
      ctx->layer_events.push_back(e);
    }
  };
  for (int layer = 0; layer <= R; ++layer) {
    const bool predict = layer == R;
    const void* in[UDAL_MAX_LEVELS];
    void* out[UDAL_MAX_LEVELS];
    const float* in_scale[UDAL_MAX_LEVELS];
    const float* ep[UDAL_MAX_LEVELS];
    const int nb_out = layer == 0 ? B : NBt;
    const int in_nb = layer <= 1 ? B : NBt;
    for (int l = 0; l < L; ++l) {
      const size_t lvl = (size_t)ctx->level_pix_off[l] * KF;
      if (layer == 0) in[l] = feats[l];
      else if (layer == 1) in[l] = a0 + (size_t)B * lvl;
      else in[l] = pp + (size_t)((layer - 1) & 1) * NBt * P * KF + (size_t)NBt * lvl;
      if (predict) out[l] = outs[l];
      else if (layer == 0) out[l] = a0 + (size_t)B * lvl;
      else out[l] = pp + (size_t)(layer & 1) * NBt * P * KF + (size_t)NBt * lvl;
      // SpatialDropout2D of the producer layer (layer - 1), applied to the depthwise output (it commutes)
      in_scale[l] = (mc && layer >= 1) ? scale_all + (((size_t)head * L + l) * R + (layer - 1)) * (size_t)NBt * KF : nullptr;
      ep[l] = predict ? nullptr : h.x3_f + ((size_t)layer * L + l) * 2 * KF;
    }
    mark();
    const float* const* sc = (mc && layer >= 1) ? in_scale : nullptr;
    if (!predict) {
      // fp32 out through the staged row path (predict = 1) with BN + swish (act = 1)
      UDAL_TRY((launch_wide_t<64, true, true, true>)(ctx, in, in_nb, nb_out, sc, h.dw + (size_t)layer * 9 * KF,
                                                     img + (size_t)layer * 2 * KF * KF, ep, 1, KF, 0, KF, out, 1));
    } else {
      for (int q = 0; q < h.x3_chunks; ++q) {
        const int n0 = q * KF, nc = h.cout - n0 < KF ? h.cout - n0 : KF;
        for (int l = 0; l < L; ++l) ep[l] = h.x3_f + ((size_t)R * L + q) * 2 * KF;
        UDAL_TRY((launch_wide_t<64, true, true, true>)(ctx, in, in_nb, nb_out, sc, h.dwp, img + (size_t)(R + q) * 2 * KF * KF, ep, 1, nc,
                                                       n0, h.cout, out, 0));
      }
    }
    mark();
  }
  return UDAL_OK;
}

int udal_heads_x3_sample(udal_ctx* ctx, const float* const* feats, int batch, const float* scale, float* const* cls_out,
                         float* const* box_out) {
  for (int l = 0; l < ctx->cfg.num_levels; ++l)
    UDAL_REQUIRE(((uintptr_t)feats[l] & 15) == 0, "level %d: feature pointers must be 16-byte aligned", l);
  UDAL_TRY(run_tower_x3(ctx, UDAL_HEAD_CLASS, feats, batch, scale, cls_out));
  UDAL_TRY(run_tower_x3(ctx, UDAL_HEAD_BOX, feats, batch, scale, box_out));
  return UDAL_OK;
}

// =====================================================================================================================
// BiFPN: OpAfterCombine's separable conv + BN of a 64-channel node on the tensor cores, fp32 accurate -
// the fp32x3 tower kernel on ONE map: fp32 depthwise on the CUDA cores, the pointwise GEMM as three fp16 tcgen05 passes over
// (hi, lo) operand pairs, epilogue = BN scale | folded bias (act 2), fp32 in and out.  The fp32 CUDA-core kernel it replaces
// took 3.2 ms of an 8.4 ms FPNCells call (D0 1280x384, batch 64).
// =====================================================================================================================
extern "C" int udal_sepconv_tc_prepare(udal_ctx* ctx, const float* pw, const float* bias, const float* bn_scale, const float* bn_shift,
                                       void** table) {
  UDAL_REQUIRE(ctx && pw && bias && table, "NULL argument");
  UDAL_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "udal_sepconv_tc_prepare: BN scale and shift go together");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  UDAL_TRY(udal_join(ctx));
  // [2][64 n][64 k] fp16 (hi, lo) images, then [2][64] fp32 (scale | folded bias)
  void* t = nullptr;
  UDAL_CUDA(cudaMalloc(&t, (size_t)2 * KF * KF * sizeof(__half) + 2 * KF * sizeof(float)));
  ctx->user_allocs.push_back(t);
  x3_weights_kernel<<<(KF * KF + 255) / 256, 256, 0, ctx->stream>>>(pw, KF, 0, KF, reinterpret_cast<__half*>(t));
  UDAL_CHECK_LAUNCH(ctx);
  float* ep = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(t) + (size_t)2 * KF * KF * sizeof(__half));
  wide_ep_kernel<<<1, KF, 0, ctx->stream>>>(bias, bn_scale, bn_shift, 0, KF, ep, KF);
  UDAL_CHECK_LAUNCH(ctx);
  *table = t;
  return UDAL_OK;
}

extern "C" int udal_sepconv_tc(udal_ctx* ctx, const float* in, int NB, int H, int W, const float* dw, const void* table, int act,
                               float* out) {
  UDAL_REQUIRE(ctx && in && dw && table && out, "NULL argument");
  UDAL_REQUIRE(act == UDAL_ACT_BN || act == UDAL_ACT_BN_SWISH, "udal_sepconv_tc: act must be UDAL_ACT_BN or UDAL_ACT_BN_SWISH");
  UDAL_REQUIRE(NB > 0 && H > 0 && W > 0, "udal_sepconv_tc: bad sizes");
  UDAL_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "udal_sepconv_tc: maps must be 16-byte aligned");
  UDAL_CUDA(cudaSetDevice(ctx->cfg.device));
  UDAL_TRY(udal_join(ctx));
  UDAL_TRY(udal_work_counters_reset(ctx));   // (stream ordered: behind the previous launch that used the counters)
  const void* ins[UDAL_MAX_LEVELS] = {in};
  void* outs[UDAL_MAX_LEVELS] = {out};
  const float* ep[UDAL_MAX_LEVELS] = {reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(table) + (size_t)2 * KF * KF * sizeof(__half))};
  return launch_wide_geom<64, true, true, true>(ctx, 1, &H, &W, KF, ins, NB, NB, nullptr, dw, table, ep, 1, KF, 0, KF, outs,
                                                act == UDAL_ACT_BN_SWISH ? 1 : 2);
}
