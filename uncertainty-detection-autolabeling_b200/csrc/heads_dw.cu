// K1 in the fp16 tensor-core mode, tower layers >= 2 and the predict layers: the separable 3x3 conv split the way the
// arithmetic asks for it -
//
//     depthwise 3x3  : CUDA cores, packed fp16 (HFMA2: 2 FMAs per lane and issue slot), 9-fold tap reuse in registers
//     pointwise 1x1  : tcgen05 tensor cores, ONE K = 64 GEMM per tile (4 x tcgen05.mma M128 N K16), fp32 accumulate in TMEM
//
// instead of the implicit GEMM of heads_ig.cu / heads_fused.cu, which folds the depthwise stage into the contraction and
// so executes 9x the algorithmic MMAs (36 per tile) and is bound by the tensor core's shared-memory operand fetch
// (0.072 of the tensor peak).  Persistent, warp specialised, one CTA per SM, dynamic work-item claiming
// (heads_umma.cuh):
//
//   warp 0        producer : one TMA box per (tile, sample): the 18x10-pixel halo tile x 64 channels fp16, linear (no swizzle),
//                            zero OOB fill = SAME padding; 3-4 stage ring
//   warp 1        MMA      : 4 x tcgen05.mma per (tile, sample); A = the builders' tile, B = the resident [N][64] fp16 image of
//                            the pointwise weights; accumulators double buffered in TMEM
//   warps 2-5     builders : dw_build_tile (heads_umma.cuh): halo tile -> A operand [128 px][64 ch] fp16, K-major, 128B swizzle,
//                            double buffered
//   warps 6-...   epilogue : tower layer: BN scale + folded bias -> swish (tanh.approx) -> this layer's SpatialDropout2D
//                            keep-scale -> fp16 staging tile -> TMA tensor store (2 groups x 4 warps);
//                            predict layer: + bias -> fp32 [T,B,H,W,Cout] (stand-alone head sampler), or
//                            fused with K2 (udal_run): the T samples of a (tile, image) run back to back and 12 epilogue warps
//                            keep the Monte-Carlo statistics of their anchors in registers - logit mean / std / argmax /
//                            sigmoid score (class head), per-sample closed-form decode + box moments + mean aleatoric std
//                            (box head) - so the [T,...] head outputs never reach HBM
//
// Reference arithmetic replaced: efficientdet_keras.py:448-483 / 628-664 (_conv_bn_act, the predict SeparableConv2D) x the
// MC loop 979-1050; utils_extra.py:220-244, utils_box.py:125-160, postprocess.py:123-135, 284, 297-331 in the fused epilogues.
// Numerics: fp16 activations and weights (11-bit significand), fp16 depthwise accumulation, fp32 GEMM accumulation and fp32
// epilogues; measured against the oracle in tests/test_gpu_bench_parity.py.
#include <type_traits>

#include "udal_common.cuh"
#include "heads_umma.cuh"
#include "decode_math.cuh"

namespace {

constexpr int DW_IN_BYTES = IG_ROWS * IG_BOXW * 128;                 // 23 040: linear halo tile
constexpr int DW_IN_STRIDE = (DW_IN_BYTES + 1023) / 1024 * 1024;     // 23 552
constexpr int kDwBuilderWarps = kDwBuilderThreads / 32;              // 4
// (first epilogue warp of heads_dw_kernel = 2 + GROUPS x 4 = 6 or 10: (warp & 3) = 2,3,0,1 - every TMEM lane quarter per 4 warps)

// weight image of one pointwise matrix: wimg[n][k] = fp16(w[k][n0 + n]) (n < cout, zero rows past it), K-major, 128B swizzle
__global__ void dw_weights_kernel(const float* __restrict__ w, int ldw, int n0, int cout, int rows, __half* __restrict__ wimg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * KF) return;
  const int k = i % KF, n = i / KF;
  const float v = n < cout ? w[(size_t)k * ldw + n0 + n] : 0.f;
  const size_t byte = (size_t)n * 128 + (size_t)((((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
  wimg[byte / 2] = __float2half_rn(v);
}

__global__ void dw_bias_kernel(const float* __restrict__ bias, int n0, int cout, int npad, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) out[i] = i < cout ? bias[n0 + i] : 0.f;
}

// =====================================================================================================================
// tower layers (fp16 out) and stand-alone predict layers (fp32 out)
// =====================================================================================================================
// threads: producer, MMA, GROUPS x 4 builder warps, 2 x 4 epilogue warps.  Tower layers: GROUPS = 2, two tiles are built
// concurrently (the builders' latency chain is the period of the pipeline); the stand-alone predict layers keep
// one group (their 64-80 KB of fp32 staging leave no room for more A buffers / stages, and at 576 threads x 96 registers
// their epilogue spills: measured 0.50 -> 0.64 ms with two)

template <int NPAD_, bool PREDICT_>
struct DwShape {
  // (the stand-alone predict shapes carry 64-80 KB of fp32 staging: 3 stages and one A buffer per builder group there)
  static constexpr int NPAD = NPAD_, NROWS = NPAD_, STAGES = PREDICT_ ? 3 : 4, ABUF = PREDICT_ ? 2 : 4;
  static constexpr int GROUPS = PREDICT_ ? 1 : 2;                          // builder groups
  static constexpr int THREADS = 64 + GROUPS * kDwBuilderThreads + 256;
  static constexpr int FIRST_EPI_WARP = 2 + GROUPS * kDwBuilderWarps;
  static constexpr bool PREDICT = PREDICT_;
  static constexpr int B_BYTES = (NROWS * 128 + 1023) / 1024 * 1024;
  // staging tile of one epilogue group: fp16 [128][64] (16 KB), or fp32 32-channel regions [128][32] (16 KB each, 128B
  // swizzle) plus a dense remainder region [128][NPAD - 64]
  static constexpr int OUT_BYTES = !PREDICT ? 16384 : (NPAD == 64 ? 32768 : 32768 + 128 * (NPAD - 64) * 4);
  static constexpr int SM_B = 0;
  static constexpr int SM_A = SM_B + B_BYTES;                 // ABUF x [128][128 B]
  static constexpr int SM_OUT = SM_A + ABUF * 16384;
  static constexpr int SM_IN = SM_OUT + 2 * OUT_BYTES;
  static constexpr int SM_BAR = SM_IN + STAGES * DW_IN_STRIDE;   // barriers (192 B) + 2 x 64 keep-scales (512 B)
  static constexpr int SM_FBS = SM_BAR + 768;
  static constexpr int smem(int levels) { return SM_FBS + levels * 2 * NPAD * 4 + IG_QRING * 4 + 1024; }
  static_assert(OUT_BYTES % 1024 == 0, "swizzled regions must stay 1 KB aligned");
};

struct DwParams {
  int num_levels, NB, items;             // items = sum_l tiles[l] * NB (level major)
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];
  void* out[UDAL_MAX_LEVELS];            // [NB,H,W,64] fp16 or [NB,H,W,ch_total] fp32
  const float* out_scale[UDAL_MAX_LEVELS];  // [NB,64] keep-scale of THIS layer's dropout (sc_stride 64) or ones (0)
  int sc_stride;
  const float* ep_scale[UDAL_MAX_LEVELS];   // [NPAD] per-level BN scale (ones for the predict layer)
  const float* ep_bias[UDAL_MAX_LEVELS];    // [NPAD] folded bias
  const float* dw;                       // [9][64] fp32 depthwise weights
  const void* wimg;                      // fp16 [NPAD][64] swizzled image of the pointwise weights
  int Cout, ch_off, ch_total;            // predictions: this launch writes channels [ch_off, ch_off + Cout) of ch_total
  int tma_store;                         // predictions through the staging tile + TMA store (Cout % 4 == 0, one chunk)
  int* counter;
};

struct DwMaps {
  CUtensorMap m[UDAL_MAX_LEVELS];   // input halo boxes {64, 10, 18, 1}, fp16, no swizzle
  CUtensorMap o[UDAL_MAX_LEVELS];   // output: fp16 [64 ch, 8, 16] box, or fp32 32-channel regions
  CUtensorMap o2[UDAL_MAX_LEVELS];  // output: fp32 remainder region (Cout % 32 channels)
};

template <class S>
__global__ void __launch_bounds__(S::THREADS, 1) heads_dw_kernel(const __grid_constant__ DwMaps maps, const DwParams p) {
  constexpr int NPAD = S::NPAD, STAGES = S::STAGES;
  constexpr uint32_t kTmemCols = NPAD <= 64 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: in_full[4] @0  in_empty[4] @32  a_full[4] @64  tfull[2] @96  tempty[2] @112  bfull @128  slot @136  a_empty[4] @144
  const uint32_t bar0 = sb + S::SM_BAR;
  const uint32_t in_full = bar0, in_empty = bar0 + 32, a_full = bar0 + 64, a_empty = bar0 + 144, bar_tfull = bar0 + 96,
                 bar_tempty = bar0 + 112, bar_b = bar0 + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::SM_BAR + 136);
  float* sFb = reinterpret_cast<float*>(smem + S::SM_FBS);
  volatile int* sQ = reinterpret_cast<volatile int*>(smem + S::SM_FBS + p.num_levels * 2 * NPAD * 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(in_full + 8 * i, 1);
      bar_init(in_empty + 8 * i, kDwBuilderWarps);   // one arrival per builder warp
    }
    for (int i = 0; i < S::ABUF; ++i) {
      bar_init(a_full + 8 * i, kDwBuilderWarps);
      bar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      bar_init(bar_tfull + 8 * i, 1);
      bar_init(bar_tempty + 8 * i, 4);               // one arrival per epilogue warp of the group
    }
    bar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + IG_QRING) sQ[threadIdx.x - 64] = -1;  // (slots are peeked ahead of their time)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + S::SM_BAR + 136), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue constants; with swish the 0.5 of x*sigmoid(x) = h*tanh(h) + h, h = x/2, is folded in
  const float half = S::PREDICT ? 1.0f : 0.5f;
  for (int e = threadIdx.x; e < p.num_levels * NPAD; e += S::THREADS) {
    const int l = e / NPAD, n = e - l * NPAD;
    sFb[(2 * l) * NPAD + n] = half * __ldg(p.ep_scale[l] + n);
    sFb[(2 * l + 1) * NPAD + n] = half * __ldg(p.ep_bias[l] + n);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer (warp-uniform loop, one elected lane issues) =====================
    if (ig_elect_one()) {
      bar_expect_tx(bar_b, S::NROWS * 128);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + S::SM_B),
                   "l"(p.wimg), "r"(S::NROWS * 128), "r"(bar_b)
                   : "memory");
    }
    __syncwarp();
    int s = 0, ph = 0;
    for (int k = 0;; ++k) {
      const int item = ig_claim(p.counter, p.items, lane);
      const IgItem w = ig_item(p, item < 0 ? 0 : item);
      if (ig_elect_one()) {
        bar_wait(in_empty + 8 * s, ph ^ 1);   // stage free (first round passes immediately)
        sQ[k & (IG_QRING - 1)] = item;        // published by the arrival on the stage's full barrier
        if (item < 0) {
          sQ[(k + 1) & (IG_QRING - 1)] = -1;  // end of the stream, for both builder / epilogue groups
          bar_arrive(in_full + 8 * s);
          if constexpr (S::GROUPS == 2) {
            const int s1 = s + 1 == STAGES ? 0 : s + 1;   // the other builder group waits for the next stage
            bar_wait(in_empty + 8 * s1, (s1 == 0 ? ph ^ 1 : ph) ^ 1);
            bar_arrive(in_full + 8 * s1);
          }
        } else {
          bar_expect_tx(in_full + 8 * s, DW_IN_BYTES);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(sb + S::SM_IN + s * DW_IN_STRIDE), "l"(&maps.m[w.l]), "r"(in_full + 8 * s), "r"(0), "r"(w.tx0 - 1),
              "r"(w.ty0 - 1), "r"(w.nb)
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
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    constexpr uint32_t idesc = ig_idesc<true>(NPAD);
    if (lane == 0) bar_wait(bar_b, 0);  // weights resident
    __syncwarp();
    for (int it = 0;; ++it) {
      const int a = it & 1, ab = it % S::ABUF;
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * NPAD);
      if (lane == 0) bar_wait(a_full + 8 * ab, (it / S::ABUF) & 1);  // the builders' A tile of this item (or the end marker)
      __syncwarp();
      if (ig_queue_read(sQ, it) < 0) {
        // end of the stream: wake both epilogue groups (their next accumulator "arrives" empty)
        if (ig_elect_one()) {
          bar_wait(bar_tempty + 8 * a, ((it >> 1) & 1) ^ 1);
          bar_arrive(bar_tfull + 8 * a);
          bar_wait(bar_tempty + 8 * (a ^ 1), (((it + 1) >> 1) & 1) ^ 1);
          bar_arrive(bar_tfull + 8 * (a ^ 1));
        }
        __syncwarp();
        break;
      }
      if (ig_elect_one()) {
        bar_wait(bar_tempty + 8 * a, ((it >> 1) & 1) ^ 1);  // accumulator drained by its epilogue group
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = ig_desc(sb + S::SM_A + ab * 16384, 1024, 0);
        const uint64_t bdesc = ig_desc(sb + S::SM_B, 1024, 0);
#pragma unroll
        for (int k = 0; k < KF / 16; ++k) ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        ig_commit(a_empty + 8 * ab);    // A buffer reusable once these MMAs retire
        ig_commit(bar_tfull + 8 * a);   // accumulator ready
      }
      __syncwarp();
    }
  } else if (warp < S::FIRST_EPI_WARP) {
    // ===================== builders: depthwise 3x3 (packed fp16) -> A operand =====================
    // group gb builds the items i = gb, gb + 2, ..: stage i % STAGES, A buffer i % ABUF
    const int gb = (warp - 2) >> 2;
    const int btid = (threadIdx.x - 64) & (kDwBuilderThreads - 1);
    DwWeights W;
    dw_load_weights(p.dw, btid & 7, W);
    for (int i = gb;; i += S::GROUPS) {
      const int ab = i % S::ABUF, s = i % STAGES, ph = (i / STAGES) & 1;
      if (lane == 0) {
        bar_wait(in_full + 8 * s, ph);                              // halo tile landed
        bar_wait(a_empty + 8 * ab, ((i / S::ABUF) & 1) ^ 1);        // the MMAs of item i - ABUF are done with this A buffer
      }
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) {  // end of the stream: pass it on to the MMA warp through the A barrier
        if (lane == 0) bar_arrive(a_full + 8 * ab);
        break;
      }
      const IgItem w = ig_item(p, item);
      dw_build_tile(smem + S::SM_IN + s * DW_IN_STRIDE, smem + S::SM_A + ab * 16384, W, btid, min(IG_TH, p.H[w.l] - w.ty0));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A tile -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        bar_arrive(a_full + 8 * ab);
        bar_arrive(in_empty + 8 * s);
      }
    }
  } else {
    // ===================== epilogue: 2 groups x 4 warps, group g drains accumulator g =====================
    const int ew = warp - S::FIRST_EPI_WARP;
    const int g = ew >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // GEMM row = pixel (m / 8, m % 8) of the tile
    const bool elected = (ew & 3) == 0 && lane == 0;
    const bool staged = !S::PREDICT || p.tma_store;
    uint8_t* const ob = smem + S::SM_OUT + g * S::OUT_BYTES;
    const uint32_t swz = (uint32_t)(m & 7);
    int pre_item = -1;  // item whose keep-scales were fetched ahead
    float4 pre_sc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = g;; it += 2) {  // group g drains the items with it % 2 == g
      const int a = g;
      if (lane == 0) bar_wait(bar_tfull + 8 * a, (it >> 1) & 1);  // one poller per warp
      __syncwarp();
      const int item = ig_queue_read(sQ, it);
      if (item < 0) break;
      const IgItem w = ig_item(p, item);
      const int nb = w.nb;
      const float* ep_s = sFb + (2 * w.l) * NPAD;
      const float* ep_b = ep_s + NPAD;
      // this item's dropout keep-scales -> this group's slot in shared memory (the previous tile's math of the group ended
      // before its second group barrier; the first barrier below publishes the slot)
      float* const sSc = reinterpret_cast<float*>(smem + S::SM_BAR + 192) + g * KF;
      if constexpr (!S::PREDICT) {
        if ((ew & 3) == 1 && lane < KF / 4) {  // (predicated, no divergence)
          const float4* sc = reinterpret_cast<const float4*>(p.out_scale[w.l] + (size_t)nb * p.sc_stride);
          reinterpret_cast<float4*>(sSc)[lane] = item == pre_item ? pre_sc : __ldg(sc + lane);
          // the group's next item (it + 2) is usually in the ring already: fetch its keep-scales now, verified against the ring
          // when the item is really due (a stale slot only costs a wasted load)
          pre_item = ig_queue_read(sQ, it + 2);
          if (pre_item >= 0 && pre_item < p.items) {
            const IgItem w2 = ig_item(p, pre_item);
            pre_sc = __ldg(reinterpret_cast<const float4*>(p.out_scale[w2.l] + (size_t)w2.nb * p.sc_stride) + lane);
          } else {
            pre_item = -1;
          }
        }
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NPAD);
      uint32_t r[NPAD / 8][8];
#pragma unroll
      for (int j = 0; j < NPAD / 8; ++j) ig_ld8(taddr + j * 8, r[j]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(bar_tempty + 8 * a);  // accumulator may be overwritten
      if (staged) {
        // the TMA store of this group's previous tile must have finished reading the staging tile
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        ig_group_sync(g);
      }
      if constexpr (!S::PREDICT) {
        // ---- tower layer: BN scale + folded bias (both halved) -> swish -> dropout keep-scale -> fp16 ----
#pragma unroll
        for (int j = 0; j < KF / 8; ++j) {
          const float4 f0 = *reinterpret_cast<const float4*>(ep_b + j * 8);
          const float4 f1 = *reinterpret_cast<const float4*>(ep_b + j * 8 + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(ep_s + j * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(ep_s + j * 8 + 4);
          const float fbv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          const float gsv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float4 s0 = *reinterpret_cast<const float4*>(sSc + j * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(sSc + j * 8 + 4);
          const float scl[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {  // packed pairs (FFMA2 / FMUL2): the same IEEE operations, half the issue slots
            const float2 h = ig_fma2(make_float2(__uint_as_float(r[j][i]), __uint_as_float(r[j][i + 1])), make_float2(gsv[i], gsv[i + 1]),
                                     make_float2(fbv[i], fbv[i + 1]));
            const float2 sw = ig_mul2(ig_fma2(h, make_float2(ig_tanh(h.x), ig_tanh(h.y)), h), make_float2(scl[i], scl[i + 1]));
            v[i] = sw.x;
            v[i + 1] = sw.y;
          }
          uint4 o;
          o.x = ig_pack16<true>(v[0], v[1]);
          o.y = ig_pack16<true>(v[2], v[3]);
          o.z = ig_pack16<true>(v[4], v[5]);
          o.w = ig_pack16<true>(v[6], v[7]);
          *reinterpret_cast<uint4*>(ob + m * 128 + (((uint32_t)j ^ swz) << 4)) = o;
        }
      } else if (staged) {
        // ---- predictions through the staging tile: 32-channel swizzled regions, then the dense remainder ----
        const int nfull = p.Cout >> 5, rem = p.Cout - (nfull << 5);
#pragma unroll
        for (int j = 0; j < NPAD / 8; ++j) {
          const int n0 = j * 8;
          const float4 f0 = *reinterpret_cast<const float4*>(ep_b + n0);
          const float4 f1 = *reinterpret_cast<const float4*>(ep_b + n0 + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(ep_s + n0);
          const float4 g1 = *reinterpret_cast<const float4*>(ep_s + n0 + 4);
          const float4 lo = make_float4(fmaf(__uint_as_float(r[j][0]), g0.x, f0.x), fmaf(__uint_as_float(r[j][1]), g0.y, f0.y),
                                        fmaf(__uint_as_float(r[j][2]), g0.z, f0.z), fmaf(__uint_as_float(r[j][3]), g0.w, f0.w));
          const float4 hi = make_float4(fmaf(__uint_as_float(r[j][4]), g1.x, f1.x), fmaf(__uint_as_float(r[j][5]), g1.y, f1.y),
                                        fmaf(__uint_as_float(r[j][6]), g1.z, f1.z), fmaf(__uint_as_float(r[j][7]), g1.w, f1.w));
          const int rg = n0 >> 5;
          if (rg < nfull) {
            const uint32_t c = (uint32_t)(n0 & 31) >> 2;  // 16-byte chunk inside the 128-byte region row
            uint8_t* row = ob + rg * 16384 + m * 128;
            *reinterpret_cast<float4*>(row + ((c ^ swz) << 4)) = lo;
            *reinterpret_cast<float4*>(row + (((c + 1) ^ swz) << 4)) = hi;
          } else {
            const int c0 = n0 - (nfull << 5);
            float* row = reinterpret_cast<float*>(ob + nfull * 16384) + m * rem;
            if (c0 + 4 <= rem) *reinterpret_cast<float4*>(row + c0) = lo;
            if (c0 + 8 <= rem) *reinterpret_cast<float4*>(row + c0 + 4) = hi;
          }
        }
      } else {
        // ---- predictions whose channel count breaks TMA's 16-byte stride rule (C = 7: 63 channels) or that are one chunk
        //      of a wider layer: dense staging tile [128 px][Cout], then coalesced 4-byte stores ----
        const int H = p.H[w.l], W = p.W[w.l], Cout = p.Cout;
        float* stg = reinterpret_cast<float*>(ob);
        ig_group_sync(g);  // the previous tile's copy loop of this group is done with the staging tile
#pragma unroll
        for (int j = 0; j < NPAD / 8; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (j * 8 + i < Cout) stg[m * Cout + j * 8 + i] = fmaf(__uint_as_float(r[j][i]), ep_s[j * 8 + i], ep_b[j * 8 + i]);
        ig_group_sync(g);
        const int gt = (ew & 3) * 32 + lane;                 // thread of the group
        const int npx = min(IG_TW, W - w.tx0), rows = min(IG_TH, H - w.ty0);
        float* dst0 = reinterpret_cast<float*>(p.out[w.l]) + (((size_t)nb * H + w.ty0) * W + w.tx0) * p.ch_total + p.ch_off;
        if (p.ch_total == Cout) {
          const int run = npx * Cout;                        // floats per tile row inside the image
          for (int row = 0; row < rows; ++row) {
            const float* src = stg + row * IG_TW * Cout;
            float* dst = dst0 + (size_t)row * W * Cout;
            for (int e = gt; e < run; e += 128) dst[e] = src[e];
          }
        } else {
          // channel chunk of a wider prediction: one run of Cout floats per pixel; warp = pixel, lane = channel
          for (int pp = gt >> 5; pp < rows * IG_TW; pp += 4) {
            const int row = pp >> 3, px = pp & 7;
            if (px < npx) {
              const float* src = stg + pp * Cout;
              float* dst = dst0 + ((size_t)row * W + px) * p.ch_total;
              for (int cc = gt & 31; cc < Cout; cc += 32) dst[cc] = src[cc];
            }
          }
        }
      }
      if (staged) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to TMA
        ig_group_sync(g);
        if (elected) {
          const uint32_t src = s32(ob);
          if constexpr (!S::PREDICT) {
            ig_tma_store(&maps.o[w.l], src, 0, w.tx0, w.ty0, nb);
          } else {
            const int nfull = p.Cout >> 5, rem = p.Cout - (nfull << 5);
            for (int rg = 0; rg < nfull; ++rg) ig_tma_store(&maps.o[w.l], src + rg * 16384, rg * 32, w.tx0, w.ty0, nb);
            if (rem) ig_tma_store(&maps.o2[w.l], src + nfull * 16384, nfull * 32, w.tx0, w.ty0, nb);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();  // lanes 1..31 must not start the next item while lane 0 still issues the store
      }
    }
    if (staged && elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

using DwTower = DwShape<64, false>;
using DwPred64 = DwShape<64, true>;
using DwPred80 = DwShape<80, true>;
static_assert(DwTower::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit && DwPred64::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit &&
              DwPred80::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit, "shared-memory budget");

template <class S>
int launch_dw(udal_ctx* ctx, const DwMaps& maps, const DwParams& p, int grid) {
  const int smem = S::smem(p.num_levels);
  UDAL_CUDA(cudaFuncSetAttribute(heads_dw_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  heads_dw_kernel<S><<<grid, S::THREADS, smem, ctx->stream>>>(maps, p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// fills the tile bookkeeping shared by both parameter structs
template <class P>
int dw_fill_geometry(udal_ctx* ctx, P& p, int NB) {
  const udal_config& c = ctx->cfg;
  p.num_levels = c.num_levels;
  p.NB = NB;
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    p.H[l] = H;
    p.W[l] = W;
    p.tiles_x[l] = (W + IG_TW - 1) / IG_TW;
    p.tiles[l] = p.tiles_x[l] * ((H + IG_TH - 1) / IG_TH);
    p.tiles_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles[l] - 1) / (uint64_t)p.tiles[l]);
    p.tiles_x_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles_x[l] - 1) / (uint64_t)p.tiles_x[l]);
    UDAL_REQUIRE((int64_t)p.tiles[l] * NB * p.tiles[l] < (1ll << 32), "level %d: too many work items for the item decode", l);
    p.item_off[l] = off;
    off += p.tiles[l] * NB;
  }
  for (int l = c.num_levels; l <= UDAL_MAX_LEVELS; ++l) p.item_off[l] = off;
  p.items = off;
  return UDAL_OK;
}

// =====================================================================================================================
// predict layer fused with K2 (udal_run, serving configuration)
// =====================================================================================================================
// 20 warps = 5 warpgroups with one role each, so that setmaxnreg can move registers from the two issue-only warps to the
// epilogue (whose per-anchor Monte-Carlo accumulators want ~120 registers per thread):
//   warpgroup 0 : warp 0 producer, warp 1 MMA issuer, warps 2-3 idle    96 -> 40 registers per thread
//   warpgroup 1 : warps 4-7 builders                                     96 -> 104
//   warpgroups 2-4 : warps 8-19 epilogue                                 96 -> 112
constexpr int kDfThreads = 640;
constexpr int kDfFirstBuilderWarp = 4, kDfFirstEpiWarp = 8;
constexpr int kDfRegsIssue = 40, kDfRegsBuild = 104, kDfRegsEpi = 112;  // 128 x (40 + 104 + 3 x 112) = 640 x 96: the launch allocation
constexpr int kDfRegsBuild2 = 104;   // (setmaxnreg can only redistribute the CTA's launch allocation, 640 x 96: a larger request would wait forever)
template <int N>
__device__ __forceinline__ void df_reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void df_reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// compile-time shape of one kernel variant (offsets from a 1024-byte aligned base).  L2 = the last tower layer runs inside
// the kernel as well: its input halo tile is 20 x 12 pixels, its output never leaves shared memory.
constexpr int DF2_ROWS = IG_ROWS + 2, DF2_BOXW = IG_BOXW + 2;      // 20 x 12 pixel halo tile of the fused tower layer
constexpr int DF2_IN_BYTES = DF2_ROWS * DF2_BOXW * 128;            // 30 720
constexpr int DF2_MID_PX = IG_ROWS * IG_BOXW;                      // 180 pixels of tower-layer output feed one predict tile
template <int NPAD_, int STAGES_, int OUT_BYTES_, int OUT2_BYTES_, bool L2_ = false>
struct DfShape {
  static constexpr int NPAD = NPAD_, STAGES = L2_ ? 2 : STAGES_;
  static constexpr bool L2 = L2_;
  static constexpr int B = 0;
  static constexpr int B_BYTES = (NPAD * 128 + 1023) / 1024 * 1024;
  static constexpr int B2 = B + B_BYTES;               // L2: [64][128 B] tower pointwise image
  static constexpr int A = B2 + (L2 ? 8192 : 0);       // 2 x [128][128 B]
  static constexpr int A2 = A + 2 * 16384;             // L2: [256][128 B] depthwise output of the tower layer (180 rows used)
  static constexpr int MID = A2 + (L2 ? 32768 : 0);    // L2: [18][10][128 B] tower-layer output = the predict layer's halo tile
  static constexpr int IN = MID + (L2 ? DW_IN_STRIDE : 0);
  static constexpr int IN_STRIDE = L2 ? DF2_IN_BYTES : DW_IN_STRIDE;
  static constexpr int IN_BYTES = L2 ? DF2_IN_BYTES : DW_IN_BYTES;
  static constexpr int OUT = IN + STAGES * IN_STRIDE;  // class: [128 px][9 NC] fp32; box: boxes | albox, 2 x [128 px][36]
  static constexpr int OUT2 = OUT + OUT_BYTES_;        // class: scores [128][9] fp32, classes [128][9] i32; box: mcbox [128][36]
  static constexpr int BAR = OUT2 + OUT2_BYTES_;       // barriers + tmem slot (256 B)
  static constexpr int BIAS = BAR + 256;               // [NPAD] fp32 predict bias
  static constexpr int QUEUE = BIAS + NPAD * 4;        // item-index ring (IG_QRING ints)
  static constexpr int QUEUEW = QUEUE + IG_QRING * 4;  // decoded items (IG_QRING int4)
  static constexpr int EP2 = QUEUEW + IG_QRING * 16;   // L2: [2][64] halved BN scale | folded bias of the current level, then [64] keep-scales
  static constexpr int DWW = EP2 + (L2 ? 3 * KF * 4 : 0);   // L2: depthwise weights [9][64] fp16 of the tower layer
  static constexpr int SMEM = DWW + (L2 ? 2 * 9 * KF * 2 : 0) + 1024;   // (tower layer's, then the predict layer's)
  static_assert(STAGES <= 4 && 2 * NPAD <= 256, "layout");
  static_assert(DF2_IN_BYTES % 1024 == 0 && A % 1024 == 0 && A2 % 1024 == 0 && IN % 128 == 0, "alignment");
  static_assert(SMEM <= kIgSmemLimit, "shared-memory budget");
};
// A = 9 anchors: 63 / 72 logits (7 / 8 classes) and the 72 box + sigma channels share one shape; 10 classes (BDD100K) = 90 logits
template <bool L2>
using DfShape72 = DfShape<80, 4, 128 * 72 * 4, 128 * 36 * 4, L2>;
template <bool L2>
using DfShape96 = DfShape<96, 4, 128 * 90 * 4, 128 * 9 * 8, L2>;
template <bool BOX, int NC, bool L2 = false>
using DfShapeOf = typename std::conditional<(BOX || NC <= 8), DfShape72<L2>, DfShape96<L2>>::type;

struct DfParams {
  int num_levels, NB, T, items;          // NB = images; items = sum_l tiles[l] * NB (level major)
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];
  int pix_off[UDAL_MAX_LEVELS + 1];      // prefix of H_l * W_l
  const float* dw;                       // [9][64] fp32 depthwise weights of the predict layer
  const void* wimg;                      // fp16 [NPAD][64] swizzled image of the predict pointwise weights
  const float* bias;                     // [NPAD]
  const float* anchors;                  // [N,4]
  int* counter;
  long long N;                           // anchors per image
  float* mean_logits;                    // class head outputs [NB,N,NC]
  float* std_logits;
  float* scores;                         // [NB,N]
  int32_t* classes;
  float* boxes;                          // box head outputs [NB,N,4]
  float* albox;
  float* mcbox;
  // L2 variant: the last tower layer (its input = maps.m, box {64,12,20,1})
  const float* dw2;                      // [9][64] fp32 depthwise weights of the tower layer
  const void* wimg2;                     // fp16 [64][64] swizzled image of its pointwise weights
  const float* ep2_scale[UDAL_MAX_LEVELS];  // [64] per-level BN scale
  const float* ep2_bias[UDAL_MAX_LEVELS];   // [64] folded bias
  const float* keep2[UDAL_MAX_LEVELS];      // [T*NB][64] keep-scale of the tower layer's own SpatialDropout2D
};

struct DfMaps {
  CUtensorMap m[UDAL_MAX_LEVELS];        // [T*NB,H,W,64] fp16 last tower layer output, box {64,10,18,1}, no swizzle
  CUtensorMap o[3][UDAL_MAX_LEVELS];     // class: mean_logits, std_logits (ch = 9 NC); box: boxes, albox, mcbox (ch = 36)
};

__device__ __forceinline__ void df_epi_sync() {  // the 384 epilogue threads
  asm volatile("bar.sync 1, 384;" ::: "memory");
}
__device__ __forceinline__ float df_exp(float x) {  // ex2.approx: ~2^-22 relative
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float df_expm1(float v) {  // v = sigma^2 >= 0
  if (v < 0.25f) {
    float q = fmaf(v, 1.f / 720.f, 1.f / 120.f);
    q = fmaf(q, v, 1.f / 24.f);
    q = fmaf(q, v, 1.f / 6.f);
    q = fmaf(q, v, 0.5f);
    q = fmaf(q, v, 1.f);
    return q * v;
  }
  return df_exp(v) - 1.f;
}
// x / T, correctly rounded (= __fdiv_rn for finite x: Markstein's sequence with the correctly rounded reciprocal of the small
// integer T; checked against IEEE division on 2.7e7 values, T = 1..64, 100, 128, 1000) in 3 instructions instead of ~20
__device__ __forceinline__ float df_div_t(float x, float fT, float rT) {
  const float q = x * rT;
  return fmaf(fmaf(-q, fT, x), rT, q);
}
// sigmoid in fp32 (ex2.approx + IEEE reciprocal, ~2 ulp): the scores of the 16-bit heads carry 1e-3 of rounding already
__device__ __forceinline__ float df_sigmoid(float x) { return __frcp_rn(1.f + df_exp(-x)); }
__device__ __forceinline__ float df_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// one axis of utils_box.py:125-160 (l-norm), fp32: sa = anchor size, ca = anchor centre, sa2 = sa * sa
__device__ __forceinline__ void df_decode_axis(float sa, float ca, float sa2, float t_c, float t_s, float s_c, float s_s,
                                               float& lo, float& hi, float& sd) {
  const float vs = s_s * s_s, vc = s_c * s_c;
  const float c = fmaf(t_c, sa, ca);
  const float e = df_exp(fmaf(0.5f, vs, t_s));
  const float half = 0.5f * e * sa;
  lo = c - half;
  hi = c + half;
  // Var(centre) + Var(size) / 4 with Var(size) = (exp(v) - 1) exp(2 t + v) sa^2
  sd = df_sqrt(sa2 * fmaf(0.25f * df_expm1(vs), e * e, vc));
}
__device__ __forceinline__ void df_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

// ---- L2 variant helpers: the last tower layer inside the fused predict kernel -------------------------------------------
__device__ __forceinline__ void df2_builder_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }  // the 4 builder warps

// depthwise 3x3 of the tower layer over the 18 x 10 positions that feed one predict tile: halo tile sIn [20][12] px x 64 ch
// fp16 (linear) -> A operand sA2 [256][128 B] (row m = y * 10 + x, rows 180.. unused), K-major, 128B swizzle.  45 units of
// 2 x 2 outputs x 8 channel groups over 128 threads = 3 passes; a unit reads a 4 x 4 pixel window row by row.
__device__ __forceinline__ void df2_build_tile(const uint8_t* __restrict__ sIn, uint8_t* __restrict__ sA2,
                                               const uint8_t* __restrict__ sW /* [9][64] fp16 */, int btid) {
  const int cg = btid & 7;
  __half2 W[9][4];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const uint4 w4 = *reinterpret_cast<const uint4*>(sW + tp * 128 + cg * 16);
    W[tp][0] = *reinterpret_cast<const __half2*>(&w4.x);
    W[tp][1] = *reinterpret_cast<const __half2*>(&w4.y);
    W[tp][2] = *reinterpret_cast<const __half2*>(&w4.z);
    W[tp][3] = *reinterpret_cast<const __half2*>(&w4.w);
  }
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int u = pass * 16 + (btid >> 3);
    if (u >= 45) continue;
    const int y0 = 2 * (u / 5), x0 = 2 * (u % 5);
    __half2 acc[2][2][4];
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[oy][ox][q] = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint4 in[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        in[c] = *reinterpret_cast<const uint4*>(sIn + (size_t)((y0 + r) * DF2_BOXW + x0 + c) * 128 + cg * 16);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int oy = r - dy;
        if (oy >= 0 && oy < 2) {
#pragma unroll
          for (int ox = 0; ox < 2; ++ox)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const __half2* v = reinterpret_cast<const __half2*>(&in[ox + dx]);
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[oy][ox][q] = __hfma2(v[q], W[dy * 3 + dx][q], acc[oy][ox][q]);
            }
        }
      }
      if (r >= 2) {
        const int oy = r - 2;
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          const int m = (y0 + oy) * IG_BOXW + x0 + ox;
          uint4 o;
          o.x = *reinterpret_cast<const uint32_t*>(&acc[oy][ox][0]);
          o.y = *reinterpret_cast<const uint32_t*>(&acc[oy][ox][1]);
          o.z = *reinterpret_cast<const uint32_t*>(&acc[oy][ox][2]);
          o.w = *reinterpret_cast<const uint32_t*>(&acc[oy][ox][3]);
          *reinterpret_cast<uint4*>(sA2 + (size_t)m * 128 + (((uint32_t)cg ^ (uint32_t)(m & 7)) << 4)) = o;
        }
      }
    }
  }
}

// NC = classes per anchor of the class head (7: KITTI map of the reference YAMLs, 8, 10: BDD100K); the box head ignores it.
// L2: the last tower layer runs inside the kernel - per (tile, sample): TMA 20 x 12 halo tile of the layer's INPUT ->
// builders: depthwise over 18 x 10 (df2_build_tile) -> 2 x 4 MMAs (M = 128 each) into TMEM columns 256.. -> builders: BN +
// swish + keep-scale -> fp16 tile [18][10] in shared memory, zero outside the image (= the SAME padding the predict layer's
// depthwise conv sees) -> dw_build_tile -> the predict GEMM -> statistics.  The tower layer's output never reaches HBM.
template <bool BOX, int NC, bool L2 = false>
__global__ void __launch_bounds__(kDfThreads, 1) heads_dwf_kernel(const __grid_constant__ DfMaps maps, const DfParams p) {
  using S = DfShapeOf<BOX, NC, L2>;
  constexpr int NPAD = S::NPAD, STAGES = S::STAGES;
  constexpr uint32_t kD2Col = 256;   // TMEM columns of the tower layer's accumulators (L2)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: in_full[4] @0  in_empty[4] @32  a_full[2] @64  a_empty[2] @80  tfull[2] @96  tempty[2] @112  bfull @128  slot @136
  //           L2: a2_full @144  a2_empty @152  d2_full @160  d2_empty @168
  const uint32_t bar0 = sb + S::BAR;
  const uint32_t in_full = bar0, in_empty = bar0 + 32, a_full = bar0 + 64, a_empty = bar0 + 80, bar_tfull = bar0 + 96,
                 bar_tempty = bar0 + 112, bar_b = bar0 + 128, a2_full = bar0 + 144, a2_empty = bar0 + 152, d2_full = bar0 + 160,
                 d2_empty = bar0 + 168;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::BAR + 136);
  float* sBias = reinterpret_cast<float*>(smem + S::BIAS);
  volatile int* sQ = reinterpret_cast<volatile int*>(smem + S::QUEUE);
  int4* sQW = reinterpret_cast<int4*>(smem + S::QUEUEW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(in_full + 8 * i, 1);
      bar_init(in_empty + 8 * i, kDwBuilderWarps);
    }
    for (int i = 0; i < 2; ++i) {
      bar_init(a_full + 8 * i, kDwBuilderWarps);
      bar_init(a_empty + 8 * i, 1);
      bar_init(bar_tfull + 8 * i, 1);
      bar_init(bar_tempty + 8 * i, 12);   // one arrival per epilogue warp
    }
    bar_init(bar_b, 1);
    if constexpr (L2) {
      bar_init(a2_full, kDwBuilderWarps);
      bar_init(a2_empty, 1);
      bar_init(d2_full, 1);
      bar_init(d2_empty, kDwBuilderWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (L2)
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sb + S::BAR + 136) : "memory");
    else
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sb + S::BAR + 136) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < NPAD) sBias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  if constexpr (L2) {
    // tower-layer depthwise weights as fp16 (the BN tables are loaded per level by the builders)
    __half* sW2 = reinterpret_cast<__half*>(smem + S::DWW);
    for (int e = threadIdx.x; e < 9 * KF; e += kDfThreads) {
      sW2[e] = __float2half_rn(__ldg(p.dw2 + e));
      sW2[9 * KF + e] = __float2half_rn(__ldg(p.dw + e));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // (each role's setmaxnreg is the first instruction of its own branch: ptxas sizes the branch by the setmaxnreg that
  // dominates it)
  if (warp == 0) {
    // ===================== producer =====================
    df_reg_dec<kDfRegsIssue>();
    if (ig_elect_one()) {
      bar_expect_tx(bar_b, NPAD * 128 + (L2 ? 8192 : 0));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + S::B),
                   "l"(p.wimg), "r"(NPAD * 128), "r"(bar_b)
                   : "memory");
      if constexpr (L2)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + S::B2),
                     "l"(p.wimg2), "r"(8192), "r"(bar_b)
                     : "memory");
    }
    __syncwarp();
    constexpr int kHalo = L2 ? 2 : 1;
    int s = 0, ph = 0;
    for (int i = 0;; ++i) {
      const int item = ig_claim(p.counter, p.items, lane);
      if (item < 0) {  // end of the stream: an empty "sample" carries it to the builders
        if (ig_elect_one()) {
          bar_wait(in_empty + 8 * s, ph ^ 1);
          sQ[i & (IG_QRING - 1)] = -1;
          bar_arrive(in_full + 8 * s);
        }
        __syncwarp();
        break;
      }
      const IgItem w = ig_item(p, item);
      for (int t = 0; t < T; ++t) {
        if (ig_elect_one()) {
          bar_wait(in_empty + 8 * s, ph ^ 1);
          if (t == 0) ig_queue_put(sQ, sQW, i, item, w);  // published by the arrival on the first sample's full barrier
          bar_expect_tx(in_full + 8 * s, S::IN_BYTES);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(sb + S::IN + s * S::IN_STRIDE), "l"(&maps.m[w.l]), "r"(in_full + 8 * s), "r"(0), "r"(w.tx0 - kHalo),
              "r"(w.ty0 - kHalo), "r"(t * p.NB + w.nb)
              : "memory");
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    df_reg_dec<kDfRegsIssue>();
    constexpr uint32_t idesc = ig_idesc<true>(NPAD);
    if (lane == 0) bar_wait(bar_b, 0);
    __syncwarp();
    // the predict GEMM of sample j (A buffer / accumulator j & 1)
    auto predict_mma = [&](int j) {
      const int a = j & 1;
      if (lane == 0) bar_wait(a_full + 8 * a, (j >> 1) & 1);
      __syncwarp();
      if (ig_elect_one()) {
        bar_wait(bar_tempty + 8 * a, ((j >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = ig_desc(sb + S::A + a * 16384, 1024, 0);
        const uint64_t bdesc = ig_desc(sb + S::B, 1024, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * NPAD);
#pragma unroll
        for (int k = 0; k < KF / 16; ++k) ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        ig_commit(a_empty + 8 * a);
        ig_commit(bar_tfull + 8 * a);
      }
      __syncwarp();
    };
    int j = 0;
    bool done = false;
    for (int i = 0; !done; ++i) {
      for (int t = 0; t < T; ++t, ++j) {
        if constexpr (L2) {
          // the tower layer's GEMM of sample j, then the predict GEMM of sample j - 1 (the builders produce in this order)
          if (lane == 0) bar_wait(a2_full, j & 1);
          __syncwarp();
          if (t == 0 && ig_queue_read(sQ, i) < 0) {
            done = true;
            break;
          }
          if (ig_elect_one()) {
            bar_wait(d2_empty, (j & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            constexpr uint32_t idesc2 = ig_idesc<true>(KF);
            const uint64_t bdesc = ig_desc(sb + S::B2, 1024, 0);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint64_t adesc = ig_desc(sb + S::A2 + half * 16384, 1024, 0);
#pragma unroll
              for (int k = 0; k < KF / 16; ++k)
                ig_mma(tmem_base + kD2Col + (uint32_t)(half * KF), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc2, k ? 1u : 0u);
            }
            ig_commit(a2_empty);
            ig_commit(d2_full);
          }
          __syncwarp();
          if (j > 0) predict_mma(j - 1);
        } else {
          const int a = j & 1;
          if (lane == 0) bar_wait(a_full + 8 * a, (j >> 1) & 1);  // the builders' A tile of this sample (or the end marker)
          __syncwarp();
          if (t == 0 && ig_queue_read(sQ, i) < 0) {
            if (ig_elect_one()) {  // wake the epilogue: its next accumulator "arrives" empty
              bar_wait(bar_tempty + 8 * a, ((j >> 1) & 1) ^ 1);
              bar_arrive(bar_tfull + 8 * a);
            }
            __syncwarp();
            done = true;
            break;
          }
          if (ig_elect_one()) {
            bar_wait(bar_tempty + 8 * a, ((j >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = ig_desc(sb + S::A + a * 16384, 1024, 0);
            const uint64_t bdesc = ig_desc(sb + S::B, 1024, 0);
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * NPAD);
#pragma unroll
            for (int k = 0; k < KF / 16; ++k) ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
            ig_commit(a_empty + 8 * a);
            ig_commit(bar_tfull + 8 * a);
          }
          __syncwarp();
        }
      }
    }
    if constexpr (L2) {
      // j = the number of samples issued to the tower GEMM: the last predict GEMM is still due, then the end marker
      if (j > 0) predict_mma(j - 1);
      const int a = j & 1;
      if (ig_elect_one()) {
        bar_wait(bar_tempty + 8 * a, ((j >> 1) & 1) ^ 1);
        bar_arrive(bar_tfull + 8 * a);
      }
      __syncwarp();
    }
  } else if (warp < kDfFirstBuilderWarp) {
    // warps 2-3: no role (they fill the issue-only warpgroup and donate their registers)
    df_reg_dec<kDfRegsIssue>();
  } else if (warp < kDfFirstEpiWarp) {
    // ===================== builders =====================
    if constexpr (L2) df_reg_inc<kDfRegsBuild2>();
    else df_reg_inc<kDfRegsBuild>();
    const int btid = threadIdx.x - 32 * kDfFirstBuilderWarp;
    if constexpr (L2) {
      const int q = warp & 3;              // TMEM lane quarter (builder warps 4..7 -> 0..3)
      float* const sEp2 = reinterpret_cast<float*>(smem + S::EP2);
      float* const sSc = sEp2 + 2 * KF;
      uint8_t* const sMid = smem + S::MID;
      int ep_level = -1;   // level whose BN tables are in shared memory (items arrive level by level)
      // tower-layer epilogue of sample jj (item w, sample t) into the mid tile, then the predict layer's depthwise pass
      auto mid_and_predict = [&](int jj, const IgItem& w, int t) {
        const int H = p.H[w.l], Wd = p.W[w.l];
        // keep-scales of this (sample, image); the barrier also orders the previous predict pass's reads of the mid tile
        if (btid < KF / 4)
          reinterpret_cast<float4*>(sSc)[btid] =
              __ldg(reinterpret_cast<const float4*>(p.keep2[w.l] + (size_t)(t * p.NB + w.nb) * KF) + btid);
        if (w.l != ep_level) {   // halved: x * sigmoid(x) = h * tanh(h) + h with h = x / 2
          ep_level = w.l;
          sEp2[btid] = 0.5f * __ldg((btid < KF ? p.ep2_scale[w.l] : p.ep2_bias[w.l]) + (btid & (KF - 1)));
        }
        if (lane == 0) bar_wait(d2_full, jj & 1);
        __syncwarp();
        df2_builder_sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* ep_s = sEp2;
        const float* ep_b = ep_s + KF;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int m = half * 128 + q * 32 + lane;              // row of the tower GEMM = pixel (m / 10, m % 10) of the mid tile
          if (half * 128 + q * 32 >= DF2_MID_PX) break;          // warp-uniform: rows 192.. do not exist
          uint32_t r[KF / 8][8];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kD2Col + (uint32_t)(half * KF);
#pragma unroll
          for (int jc = 0; jc < KF / 8; ++jc) ig_ld8(taddr + jc * 8, r[jc]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (m < DF2_MID_PX) {
            const int py = m / IG_BOXW, px = m - py * IG_BOXW;
            const int gy = w.ty0 - 1 + py, gx = w.tx0 - 1 + px;
            const bool inside = gy >= 0 && gy < H && gx >= 0 && gx < Wd;
            uint4* dst = reinterpret_cast<uint4*>(sMid + (size_t)m * 128);
#pragma unroll
            for (int jc = 0; jc < KF / 8; ++jc) {
              uint4 o = make_uint4(0u, 0u, 0u, 0u);
              if (inside) {
                const float4 f0 = *reinterpret_cast<const float4*>(ep_b + jc * 8), f1 = *reinterpret_cast<const float4*>(ep_b + jc * 8 + 4);
                const float4 g0 = *reinterpret_cast<const float4*>(ep_s + jc * 8), g1 = *reinterpret_cast<const float4*>(ep_s + jc * 8 + 4);
                const float4 s0 = *reinterpret_cast<const float4*>(sSc + jc * 8), s1 = *reinterpret_cast<const float4*>(sSc + jc * 8 + 4);
                const float fbv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                const float gsv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float scl[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  const float2 h = ig_fma2(make_float2(__uint_as_float(r[jc][e]), __uint_as_float(r[jc][e + 1])), make_float2(gsv[e], gsv[e + 1]),
                                           make_float2(fbv[e], fbv[e + 1]));
                  const float2 sw = ig_mul2(ig_fma2(h, make_float2(ig_tanh(h.x), ig_tanh(h.y)), h), make_float2(scl[e], scl[e + 1]));
                  v[e] = sw.x;
                  v[e + 1] = sw.y;
                }
                o.x = ig_pack16<true>(v[0], v[1]);
                o.y = ig_pack16<true>(v[2], v[3]);
                o.z = ig_pack16<true>(v[4], v[5]);
                o.w = ig_pack16<true>(v[6], v[7]);
              }
              dst[jc] = o;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(d2_empty);     // the tower accumulators may be overwritten
        df2_builder_sync();                      // mid tile complete
        const int ab = jj & 1;
        if (lane == 0) bar_wait(a_empty + 8 * ab, ((jj >> 1) & 1) ^ 1);
        __syncwarp();
        {
          DwWeights W3;   // (from shared memory, so that the 36 registers are live during this pass only)
          const uint8_t* sW3 = smem + S::DWW + 9 * KF * 2;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const uint4 w4 = *reinterpret_cast<const uint4*>(sW3 + tp * 128 + (btid & 7) * 16);
            W3.w[tp][0] = *reinterpret_cast<const __half2*>(&w4.x);
            W3.w[tp][1] = *reinterpret_cast<const __half2*>(&w4.y);
            W3.w[tp][2] = *reinterpret_cast<const __half2*>(&w4.z);
            W3.w[tp][3] = *reinterpret_cast<const __half2*>(&w4.w);
          }
          dw_build_tile(sMid, smem + S::A + ab * 16384, W3, btid, min(IG_TH, H - w.ty0));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(a_full + 8 * ab);
      };
      int s = 0, ph = 0, j = 0;
      bool done = false, have_prev = false;
      IgItem prev_w = {0, 0, 0, 0};
      int prev_t = 0;
      for (int i = 0; !done; ++i) {
        IgItem w = {0, 0, 0, 0};
        for (int t = 0; t < T; ++t, ++j) {
          if (lane == 0) {
            bar_wait(in_full + 8 * s, ph);
            bar_wait(a2_empty, (j & 1) ^ 1);
          }
          __syncwarp();
          if (t == 0) {
            const int item = ig_queue_read(sQ, i);
            if (item < 0) {
              done = true;
              break;
            }
            w = ig_item(p, item);
          }
          df2_build_tile(smem + S::IN + s * S::IN_STRIDE, smem + S::A2, smem + S::DWW, btid);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            bar_arrive(a2_full);
            bar_arrive(in_empty + 8 * s);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
          if (have_prev) mid_and_predict(j - 1, prev_w, prev_t);
          prev_w = w;
          prev_t = t;
          have_prev = true;
        }
      }
      if (have_prev) mid_and_predict(j - 1, prev_w, prev_t);   // flush the pipeline, then pass the end marker on
      if (lane == 0) bar_arrive(a2_full);
    } else {
    DwWeights W;
    dw_load_weights(p.dw, btid & 7, W);
    int s = 0, ph = 0, j = 0;
    bool done = false;
    for (int i = 0; !done; ++i) {
      int rows_valid = IG_TH;
      for (int t = 0; t < T; ++t, ++j) {
        const int ab = j & 1;
        if (lane == 0) {
          bar_wait(in_full + 8 * s, ph);
          bar_wait(a_empty + 8 * ab, ((j >> 1) & 1) ^ 1);
        }
        __syncwarp();
        if (t == 0) {
          const int item = ig_queue_read(sQ, i);
          if (item < 0) {
            if (lane == 0) bar_arrive(a_full + 8 * ab);
            done = true;
            break;
          }
          const IgItem w = ig_queue_item(sQW, i);
          rows_valid = min(IG_TH, p.H[w.l] - w.ty0);
        }
        dw_build_tile(smem + S::IN + s * DW_IN_STRIDE, smem + S::A + ab * 16384, W, btid, rows_valid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    }
    }
  } else {
    // ===================== epilogue: thread = (pixel, anchors 3 cg .. 3 cg + 2) =====================
    df_reg_inc<kDfRegsEpi>();
    const int ew = warp - kDfFirstEpiWarp;
    const int etid = threadIdx.x - 32 * kDfFirstEpiWarp;
    const int q = warp & 3;         // TMEM lane quarter
    const int cg = ew >> 2;         // anchor triple
    const int m = q * 32 + lane;    // pixel (m / 8, m % 8) of the tile
    const bool elected = etid == 0;
    const float fT = (float)T, rT = 1.f / fT;
    float* const sOut = reinterpret_cast<float*>(smem + S::OUT);
    float* const sOut2 = reinterpret_cast<float*>(smem + S::OUT2);
    int j = 0;
    for (int i = 0;; ++i) {
      if (lane == 0) bar_wait(bar_tfull + 8 * (j & 1), (j >> 1) & 1);  // first sample of the item (or the end marker)
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) break;
      const IgItem w = ig_queue_item(sQW, i);
      const int H = p.H[w.l], W = p.W[w.l];
      const int oy = w.ty0 + (m >> 3), ox = w.tx0 + (m & 7);
      const bool ok = oy < H && ox < W;
      if (!BOX) {
        // ---- class head: 3 anchors x NC classes = CH logits per thread, at accumulator columns cg * CH .. ----
        constexpr int CH = 3 * NC, ROW = 9 * NC;  // logits per thread / per pixel
        constexpr bool TMA_OUT = (ROW * 4) % 16 == 0;  // TMA needs 16-byte strides: NC = 8 (288 B), not 7 (252 B)
        float sum[CH], x0[CH], s2[CH];
        for (int t = 0; t < T; ++t, ++j) {
          const int a = j & 1;
          if (lane == 0) bar_wait(bar_tfull + 8 * a, (j >> 1) & 1);
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NPAD + cg * CH);
          constexpr int NLD = (CH + 7) / 8;
          static_assert(2 * CH + NLD * 8 <= NPAD, "accumulator columns read past the thread's logits stay inside N");
          if constexpr (NC > 8) {
            // 30 logits x (sum, first sample, squared deviations) already fill the register file: the accumulator is read 8
            // columns at a time
            static_assert(CH % 2 == 0, "packed pairs");
#pragma unroll
            for (int u = 0; u < NLD; ++u) {
              uint32_t r8[8];
              ig_ld8(taddr + u * 8, r8);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const int c = u * 8 + e;
                if (c < CH) {
                  const float2 x = ig_add2(make_float2(__uint_as_float(r8[e]), __uint_as_float(r8[e + 1])),
                                           *reinterpret_cast<const float2*>(sBias + cg * CH + c));
                  if (t == 0) {
                    sum[c] = x.x; sum[c + 1] = x.y;
                    x0[c] = x.x; x0[c + 1] = x.y;
                    s2[c] = s2[c + 1] = 0.f;
                  } else {
                    const float2 sm = ig_add2(make_float2(sum[c], sum[c + 1]), x);
                    const float2 d = ig_sub2(x, make_float2(x0[c], x0[c + 1]));
                    const float2 qq = ig_fma2(d, d, make_float2(s2[c], s2[c + 1]));
                    sum[c] = sm.x; sum[c + 1] = sm.y;
                    s2[c] = qq.x; s2[c + 1] = qq.y;
                  }
                }
              }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(bar_tempty + 8 * a);
            continue;
          }
          uint32_t r[NLD][8];  // NLD x 8 columns from the thread's first one (the tail past CH is not used)
#pragma unroll
          for (int u = 0; u < NLD; ++u) ig_ld8(taddr + u * 8, r[u]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) bar_arrive(bar_tempty + 8 * a);
          if constexpr (NC == 8) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const float4 b0 = *reinterpret_cast<const float4*>(sBias + cg * 24 + u * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sBias + cg * 24 + u * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; e += 2) {  // packed pairs (FADD2 / FFMA2): the same IEEE operations, half the issue slots
                const int c = u * 8 + e;
                // = fma(acc, 1, bias) of the predict layer
                const float2 x = ig_add2(make_float2(__uint_as_float(r[u][e]), __uint_as_float(r[u][e + 1])), make_float2(bb[e], bb[e + 1]));
                if (t == 0) {
                  sum[c] = x.x; sum[c + 1] = x.y;
                  x0[c] = x.x; x0[c + 1] = x.y;
                  s2[c] = s2[c + 1] = 0.f;
                } else {
                  const float2 sm = ig_add2(make_float2(sum[c], sum[c + 1]), x);
                  const float2 d = ig_sub2(x, make_float2(x0[c], x0[c + 1]));
                  const float2 qq = ig_fma2(d, d, make_float2(s2[c], s2[c + 1]));
                  sum[c] = sm.x; sum[c + 1] = sm.y;
                  s2[c] = qq.x; s2[c + 1] = qq.y;
                }
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const float x = __fadd_rn(__uint_as_float(r[c >> 3][c & 7]), sBias[cg * CH + c]);
              if (t == 0) {
                sum[c] = x;
                x0[c] = x;
                s2[c] = 0.f;
              } else {
                sum[c] = __fadd_rn(sum[c], x);
                const float d = x - x0[c];
                s2[c] = fmaf(d, d, s2[c]);
              }
            }
          }
        }
        // ---- item done: mean / std / score / class of 128 pixels x 9 anchors leave through staging tiles ----
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous item's stores
        __syncwarp();
        df_epi_sync();
        float* st = sOut + m * ROW + cg * CH;
        float mean[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) mean[c] = df_div_t(sum[c], fT, rT);
        if constexpr (CH % 4 == 0 && ROW % 4 == 0) {
#pragma unroll
          for (int v = 0; v < CH / 4; ++v)
            reinterpret_cast<float4*>(st)[v] = make_float4(mean[4 * v], mean[4 * v + 1], mean[4 * v + 2], mean[4 * v + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < CH; ++c) st[c] = mean[c];
        }
#pragma unroll
        for (int ai = 0; ai < 3; ++ai) {
          float best = mean[ai * NC];
          int arg = 0;
#pragma unroll
          for (int c = 1; c < NC; ++c)
            if (mean[ai * NC + c] > best) {
              best = mean[ai * NC + c];
              arg = c;
            }
          sOut2[m * 9 + cg * 3 + ai] = df_sigmoid(best);
          reinterpret_cast<int32_t*>(sOut2 + 128 * 9)[m * 9 + cg * 3 + ai] = arg;
        }
        // a tile row (8 px) of a per-anchor tensor with `width` values per pixel is one contiguous run in global memory
        auto copy_rows = [&](const float* src, float* dst_base, int width) {
          for (int idx = etid; idx < IG_TH * IG_TW * width; idx += 384) {
            const int row = idx / (IG_TW * width), col = idx - row * (IG_TW * width);
            if (w.ty0 + row < H && w.tx0 + col / width < W)
              dst_base[((size_t)w.nb * (size_t)(p.N / 9) + (size_t)(p.pix_off[w.l] + (w.ty0 + row) * W + w.tx0)) * width + col] = src[idx];
          }
        };
        if constexpr (TMA_OUT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        df_epi_sync();
        if constexpr (TMA_OUT) {
          if (elected) {
            ig_tma_store(&maps.o[0][w.l], s32(sOut), 0, w.tx0, w.ty0, w.nb);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
        } else {
          copy_rows(sOut, p.mean_logits, ROW);
        }
        for (int idx = etid; idx < IG_TH * 72; idx += 384) {  // scores and classes: 8 px x 9 anchors per tile row
          const int row = idx / 72, col = idx - row * 72;
          if (w.ty0 + row < H && w.tx0 + col / 9 < W) {
            const size_t o = (size_t)w.nb * (size_t)p.N + 9ull * (size_t)(p.pix_off[w.l] + (w.ty0 + row) * W + w.tx0) + col;
            p.scores[o] = sOut2[idx];
            p.classes[o] = reinterpret_cast<const int32_t*>(sOut2 + 128 * 9)[idx];
          }
        }
        __syncwarp();
        df_epi_sync();  // the mean tile has been read
        // population std, shifted one-pass form: s1 = sum of the deviations from the first sample
        if constexpr (CH % 4 == 0 && ROW % 4 == 0) {
#pragma unroll
          for (int v = 0; v < CH / 4; ++v) {
            float sd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * v + e;
              const float s1 = sum[c] - fT * x0[c];
              sd[e] = df_sqrt(fmaxf(fmaf(-s1 * rT, s1, s2[c]), 0.f) * rT);
            }
            reinterpret_cast<float4*>(st)[v] = make_float4(sd[0], sd[1], sd[2], sd[3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float s1 = sum[c] - fT * x0[c];
            st[c] = df_sqrt(fmaxf(fmaf(-s1 * rT, s1, s2[c]), 0.f) * rT);
          }
        }
        if constexpr (TMA_OUT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        df_epi_sync();
        if constexpr (TMA_OUT) {
          if (elected) {
            ig_tma_store(&maps.o[1][w.l], s32(sOut), 0, w.tx0, w.ty0, w.nb);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          copy_rows(sOut, p.std_logits, ROW);
        }
        __syncwarp();
      } else {
        // ---- box head: per sample decode of 3 anchors x 2 axes, running statistics ----
        float sa[6], ca[6], sa2[6];
        {
          const long long n0 = ok ? 9ll * (p.pix_off[w.l] + (long long)oy * W + ox) + 3 * cg : 0;
#pragma unroll
          for (int ai = 0; ai < 3; ++ai) {
            const float4 an = __ldg(reinterpret_cast<const float4*>(p.anchors) + n0 + ai);
            sa[ai * 2] = an.z - an.x;
            ca[ai * 2] = 0.5f * (an.x + an.z);
            sa[ai * 2 + 1] = an.w - an.y;
            ca[ai * 2 + 1] = 0.5f * (an.y + an.w);
            sa2[ai * 2] = sa[ai * 2] * sa[ai * 2];
            sa2[ai * 2 + 1] = sa[ai * 2 + 1] * sa[ai * 2 + 1];
          }
        }
        float sum_lo[6], sum_hi[6], al[6], x0_lo[6], x0_hi[6], s1_lo[6], s1_hi[6], s2_lo[6], s2_hi[6];
        for (int t = 0; t < T; ++t, ++j) {
          const int a = j & 1;
          if (lane == 0) bar_wait(bar_tfull + 8 * a, (j >> 1) & 1);
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // columns: regression targets of anchor a at 4a..4a+3, sigmas at 36 + 4a..
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NPAD + cg * 12);
          uint32_t rt8[8], rt4[4], rs8[8], rs4[4];
          ig_ld8(taddr, rt8);
          df_ld4(taddr + 8, rt4);
          ig_ld8(taddr + 36, rs8);
          df_ld4(taddr + 44, rs4);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) bar_arrive(bar_tempty + 8 * a);
          float tv[12], sv[12];
#pragma unroll
          for (int e = 0; e < 12; ++e) {
            tv[e] = __fadd_rn(__uint_as_float(e < 8 ? rt8[e] : rt4[e - 8]), sBias[cg * 12 + e]);
            sv[e] = __fadd_rn(__uint_as_float(e < 8 ? rs8[e] : rs4[e - 8]), sBias[36 + cg * 12 + e]);
          }
#pragma unroll
          for (int ai = 0; ai < 3; ++ai)
#pragma unroll
            for (int ax = 0; ax < 2; ++ax) {
              const int k = ai * 2 + ax;
              float lo, hi, sd;
              df_decode_axis(sa[k], ca[k], sa2[k], tv[ai * 4 + ax], tv[ai * 4 + 2 + ax], sv[ai * 4 + ax], sv[ai * 4 + 2 + ax], lo, hi,
                             sd);
              if (t == 0) {
                sum_lo[k] = lo; sum_hi[k] = hi; al[k] = sd;
                x0_lo[k] = lo; x0_hi[k] = hi;
                s1_lo[k] = s1_hi[k] = s2_lo[k] = s2_hi[k] = 0.f;
              } else {
                sum_lo[k] = __fadd_rn(sum_lo[k], lo);
                sum_hi[k] = __fadd_rn(sum_hi[k], hi);
                al[k] = __fadd_rn(al[k], sd);
                const float e0 = lo - x0_lo[k], e1 = hi - x0_hi[k];
                s1_lo[k] += e0;
                s1_hi[k] += e1;
                s2_lo[k] = fmaf(e0, e0, s2_lo[k]);
                s2_hi[k] = fmaf(e1, e1, s2_hi[k]);
              }
            }
        }
        // ---- item done: boxes | albox in the first staging tile, mcbox in the second ----
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        df_epi_sync();
        float4* stb = reinterpret_cast<float4*>(sOut + m * 36 + cg * 12);
        float4* sta = reinterpret_cast<float4*>(sOut + 128 * 36 + m * 36 + cg * 12);
        float4* stm = reinterpret_cast<float4*>(sOut2 + m * 36 + cg * 12);
#pragma unroll
        for (int ai = 0; ai < 3; ++ai) {
          const int ky = ai * 2, kx = ai * 2 + 1;
          stb[ai] = make_float4(df_div_t(sum_lo[ky], fT, rT), df_div_t(sum_lo[kx], fT, rT), df_div_t(sum_hi[ky], fT, rT),
                                df_div_t(sum_hi[kx], fT, rT));
          const float ay = df_div_t(al[ky], fT, rT), ax2 = df_div_t(al[kx], fT, rT);
          sta[ai] = make_float4(ay, ax2, ay, ax2);
          stm[ai] = make_float4(df_sqrt(fmaxf(fmaf(-s1_lo[ky] * rT, s1_lo[ky], s2_lo[ky]), 0.f) * rT),
                                df_sqrt(fmaxf(fmaf(-s1_lo[kx] * rT, s1_lo[kx], s2_lo[kx]), 0.f) * rT),
                                df_sqrt(fmaxf(fmaf(-s1_hi[ky] * rT, s1_hi[ky], s2_hi[ky]), 0.f) * rT),
                                df_sqrt(fmaxf(fmaf(-s1_hi[kx] * rT, s1_hi[kx], s2_hi[kx]), 0.f) * rT));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        df_epi_sync();
        if (elected) {
          ig_tma_store(&maps.o[0][w.l], s32(sOut), 0, w.tx0, w.ty0, w.nb);
          ig_tma_store(&maps.o[1][w.l], s32(sOut + 128 * 36), 0, w.tx0, w.ty0, w.nb);
          ig_tma_store(&maps.o[2][w.l], s32(sOut2), 0, w.tx0, w.ty0, w.nb);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (L2)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

template <bool BOX, int NC, bool L2 = false>
int launch_dwf(udal_ctx* ctx, const DfMaps& maps, const DfParams& p, int grid) {
  using S = DfShapeOf<BOX, NC, L2>;
  UDAL_CUDA(cudaFuncSetAttribute(heads_dwf_kernel<BOX, NC, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
  heads_dwf_kernel<BOX, NC, L2><<<grid, kDfThreads, S::SMEM, ctx->stream>>>(maps, p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

}  // namespace

// ---- weight tables of one head (fp16 mode) -------------------------------------------------------------------------
// dwh_w: fp16 images [tower layers 2..R-1][64][64] | predict chunks [pred_chunks][dwh_rows][64] | fused predict [dwh_frows][64]
// dwh_f: fp32 predict bias per chunk [pred_chunks][dwh_rows] | fused bias [dwh_frows] | ones [96]
int udal_heads_dw_prepare(udal_ctx* ctx, int head) {
  const udal_config& c = ctx->cfg;
  udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(c.num_filters == KF && c.repeats >= 2, "fp16 tensor-core heads: fpn_num_filters 64, box_class_repeats >= 2");
  const int R = c.repeats;
  // stand-alone predict layer: one launch of up to 80 channels, or equal chunks of at most 64 (as the bf16 path)
  const int chunks = h.cout <= 80 ? 1 : (h.cout + KF - 1) / KF;
  const int chunk = chunks == 1 ? h.cout : (h.cout + chunks - 1) / chunks;
  const int rows = chunks == 1 ? (h.cout <= 64 ? 64 : 80) : 64;
  const int frows = h.cout <= 72 ? 80 : (h.cout <= 90 ? 96 : 0);   // fused predict + K2 kernels (A = 9; C = 7, 8, 10)
  h.dwh_chunks = chunks;
  h.dwh_chunk = chunk;
  h.dwh_rows = rows;
  h.dwh_frows = frows;
  if (h.dwh_w) UDAL_CUDA(cudaFree(h.dwh_w));
  if (h.dwh_f) UDAL_CUDA(cudaFree(h.dwh_f));
  h.dwh_w = nullptr;
  h.dwh_f = nullptr;
  const size_t n_img = (size_t)(R - 2) * KF * KF + (size_t)chunks * rows * KF + (size_t)frows * KF;
  const size_t n_f = (size_t)chunks * rows + frows + 96;
  UDAL_CUDA(cudaMalloc(&h.dwh_w, n_img * 2 + 1024));
  UDAL_CUDA(cudaMalloc(&h.dwh_f, n_f * sizeof(float)));
  __half* img = reinterpret_cast<__half*>(h.dwh_w);
  const int tb = 256;
  for (int r = 2; r < R; ++r) {
    dw_weights_kernel<<<(KF * KF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pw + (size_t)r * KF * KF, KF, 0, KF, KF, img + (size_t)(r - 2) * KF * KF);
    UDAL_CHECK_LAUNCH(ctx);
  }
  __half* pimg = img + (size_t)(R - 2) * KF * KF;
  for (int q = 0; q < chunks; ++q) {
    const int n0 = q * chunk, nc = h.cout - n0 < chunk ? h.cout - n0 : chunk;
    dw_weights_kernel<<<(rows * KF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pwp, h.cout, n0, nc, rows, pimg + (size_t)q * rows * KF);
    UDAL_CHECK_LAUNCH(ctx);
    dw_bias_kernel<<<1, 128, 0, ctx->stream>>>(h.bp, n0, nc, rows, h.dwh_f + (size_t)q * rows);
    UDAL_CHECK_LAUNCH(ctx);
  }
  if (frows) {
    dw_weights_kernel<<<(frows * KF + tb - 1) / tb, tb, 0, ctx->stream>>>(h.pwp, h.cout, 0, h.cout, frows, pimg + (size_t)chunks * rows * KF);
    UDAL_CHECK_LAUNCH(ctx);
    dw_bias_kernel<<<1, 128, 0, ctx->stream>>>(h.bp, 0, h.cout, frows, h.dwh_f + (size_t)chunks * rows);
    UDAL_CHECK_LAUNCH(ctx);
  }
  std::vector<float> one(96, 1.0f);
  UDAL_CUDA(cudaMemcpyAsync(h.dwh_f + (size_t)chunks * rows + frows, one.data(), 96 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  UDAL_CUDA(cudaStreamSynchronize(ctx->stream));
  return UDAL_OK;
}

int udal_heads_dw_fused_ok(const udal_ctx* ctx, int head) { return ctx->heads[head].dwh_frows != 0; }

// tower layer `layer` (2 <= layer < R; out[l] fp16 [NB,H,W,64] = swish(BN(sepconv(in))) * out_scale) or, layer == R, the
// predict layer writing fp32 [NB,H,W,cout] (stand-alone head sampler).  in[l]: fp16 [NB,H_l,W_l,64], dropout already applied.
int udal_heads_dw_layer(udal_ctx* ctx, int head, int layer, const void* const* in, int NB, const float* const* ep_scale,
                        const float* const* ep_bias, const float* const* out_scale, void* const* out) {
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(h.dwh_w && h.dwh_f, "fp16 head tables not built");
  const int R = c.repeats;
  const bool predict = layer == R;
  UDAL_REQUIRE(layer >= 2 && layer <= R, "heads_dw: layer %d", layer);
  const __half* img = reinterpret_cast<const __half*>(h.dwh_w);
  const __half* pimg = img + (size_t)(R - 2) * KF * KF;
  const float* ones = h.dwh_f + (size_t)h.dwh_chunks * h.dwh_rows + h.dwh_frows;
  const int launches = predict ? h.dwh_chunks : 1;
  for (int q = 0; q < launches; ++q) {
    DwMaps maps;
    DwParams p;
    memset(&p, 0, sizeof(p));
    memset(&maps, 0, sizeof(maps));
    UDAL_TRY(dw_fill_geometry(ctx, p, NB));
    const int n0 = predict ? q * h.dwh_chunk : 0;
    const int cout = predict ? (h.cout - n0 < h.dwh_chunk ? h.cout - n0 : h.dwh_chunk) : KF;
    const int ch_total = predict ? h.cout : KF;
    p.Cout = cout;
    p.ch_off = n0;
    p.ch_total = ch_total;
    // TMA needs 16-byte global strides: every fp16 layer qualifies, fp32 predictions when Cout % 4 == 0 (single chunk)
    p.tma_store = (!predict || ((cout & 3) == 0 && ch_total == cout)) ? 1 : 0;
    p.sc_stride = (!predict && out_scale) ? KF : 0;
    p.dw = predict ? h.dwp : h.dw + (size_t)layer * 9 * KF;
    p.wimg = predict ? pimg + (size_t)q * h.dwh_rows * KF : img + (size_t)(layer - 2) * KF * KF;
    const float* pbias = h.dwh_f + (size_t)q * h.dwh_rows;
    for (int l = 0; l < c.num_levels; ++l) {
      const int H = c.level_h[l], W = c.level_w[l];
      UDAL_TRY(encode_nhwc(encode, &maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, in[l], NB, H, W, KF, KF, IG_BOXW, IG_ROWS, false));
      if (!predict) {
        UDAL_TRY(encode_nhwc(encode, &maps.o[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out[l], NB, H, W, KF, KF, IG_TW, IG_TH, true));
      } else if (p.tma_store) {
        const int nfull = cout / 32, rem = cout % 32;
        if (nfull)
          UDAL_TRY(encode_nhwc(encode, &maps.o[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out[l], NB, H, W, cout, 32, IG_TW, IG_TH, true));
        if (rem)
          UDAL_TRY(encode_nhwc(encode, &maps.o2[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out[l], NB, H, W, cout, rem, IG_TW, IG_TH, false));
      }
      p.out[l] = out[l];
      p.out_scale[l] = (!predict && out_scale) ? out_scale[l] : ones;
      p.ep_scale[l] = predict ? ones : ep_scale[l];
      p.ep_bias[l] = predict ? pbias : ep_bias[l];
    }
    UDAL_TRY(udal_work_counter(ctx, &p.counter));
    const int grid = udal_persistent_grid(ctx, p.items);
    if (!predict) UDAL_TRY(launch_dw<DwTower>(ctx, maps, p, grid));
    else if (h.dwh_rows == 64) UDAL_TRY(launch_dw<DwPred64>(ctx, maps, p, grid));
    else UDAL_TRY(launch_dw<DwPred80>(ctx, maps, p, grid));
  }
  return UDAL_OK;
}

// predict layer of one head fused with the MC moments (class) / decode + MC moments (box): in[l] = last tower layer output
// [T*NB,H_l,W_l,64] fp16; writes the per-anchor tensors of `pre` that belong to the head.
struct DfL2Args {            // the last tower layer fused in (null: `in` is that layer's output)
  int layer;                 // its index (R - 1)
  const float* const* ep_scale;   // [L] -> [64] BN scale
  const float* const* ep_bias;    // [L] -> [64] folded bias
  const float* const* keep;       // [L] -> [T*NB][64] keep-scale of its own dropout
};
static int dw_fused_predict_impl(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const udal_prenms_out* pre,
                                 const DfL2Args* l2);
int udal_heads_dw_fused_predict(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const udal_prenms_out* pre) {
  return dw_fused_predict_impl(ctx, head, in, NB, T, pre, nullptr);
}
// in[l] = the INPUT of tower layer `layer` = R - 1 ([T*NB,H_l,W_l,64] fp16, dropout applied); that layer, the predict layer and
// K2 run in one kernel
int udal_heads_dw_fused_l2(udal_ctx* ctx, int head, int layer, const void* const* in, int NB, int T, const float* const* ep_scale,
                           const float* const* ep_bias, const float* const* keep, const udal_prenms_out* pre) {
  UDAL_REQUIRE(layer == ctx->cfg.repeats - 1 && layer >= 2, "fused tower layer: the last one (>= 2)");
  DfL2Args a = {layer, ep_scale, ep_bias, keep};
  return dw_fused_predict_impl(ctx, head, in, NB, T, pre, &a);
}
static int dw_fused_predict_impl(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const udal_prenms_out* pre,
                                 const DfL2Args* l2) {
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  UDAL_REQUIRE(ctx->anchors_set, "anchor table not set");
  const udal_config& c = ctx->cfg;
  const udal_head_weights_dev& h = ctx->heads[head];
  UDAL_REQUIRE(h.dwh_w && h.dwh_f && h.dwh_frows, "fused fp16 predict kernels: configuration not covered");
  const int R = c.repeats;
  DfMaps maps;
  DfParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  UDAL_TRY(dw_fill_geometry(ctx, p, NB));
  p.T = T;
  for (int l = 0; l < c.num_levels; ++l)
    UDAL_TRY(encode_nhwc(encode, &maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, in[l], T * NB, c.level_h[l], c.level_w[l], KF, KF,
                         l2 ? DF2_BOXW : IG_BOXW, l2 ? DF2_ROWS : IG_ROWS, false));
  if (l2) {
    p.dw2 = h.dw + (size_t)l2->layer * 9 * KF;
    p.wimg2 = reinterpret_cast<const __half*>(h.dwh_w) + (size_t)(l2->layer - 2) * KF * KF;
    for (int l = 0; l < c.num_levels; ++l) {
      p.ep2_scale[l] = l2->ep_scale[l];
      p.ep2_bias[l] = l2->ep_bias[l];
      p.keep2[l] = l2->keep[l];
    }
  }
  for (int l = 0; l <= UDAL_MAX_LEVELS; ++l) p.pix_off[l] = (int)ctx->level_pix_off[l < c.num_levels ? l : c.num_levels];
  p.dw = h.dwp;
  p.wimg = reinterpret_cast<const __half*>(h.dwh_w) + (size_t)(R - 2) * KF * KF + (size_t)h.dwh_chunks * h.dwh_rows * KF;
  p.bias = h.dwh_f + (size_t)h.dwh_chunks * h.dwh_rows;
  p.anchors = ctx->anchors;
  p.N = ctx->num_anchors;
  // [NB, H_l, W_l, ch] views of the level's slice of a per-anchor tensor with `ch` floats per pixel
  auto out_map = [&](CUtensorMap* map, const float* base, int l, int ch) -> int {
    const int H = c.level_h[l], W = c.level_w[l];
    const cuuint64_t gdim[4] = {(cuuint64_t)ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    const cuuint64_t gstr[3] = {(cuuint64_t)ch * 4, (cuuint64_t)W * ch * 4, (cuuint64_t)ctx->num_pixels * ch * 4};
    return encode_strided(encode, map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base + (size_t)ctx->level_pix_off[l] * ch, gdim, gstr, ch,
                          IG_TW, IG_TH, false);
  };
  UDAL_TRY(udal_work_counter(ctx, &p.counter));
  const int grid = udal_persistent_grid(ctx, p.items);
  if (head == UDAL_HEAD_CLASS) {
    UDAL_REQUIRE(pre->mean_logits && pre->std_logits && pre->scores && pre->classes, "fused class head: NULL output");
    p.mean_logits = pre->mean_logits;
    p.std_logits = pre->std_logits;
    p.scores = pre->scores;
    p.classes = pre->classes;
    if (c.num_classes == 8) {
      UDAL_REQUIRE(h.dwh_frows == 80, "fused class head: weight image");
      for (int l = 0; l < c.num_levels; ++l) {
        UDAL_TRY(out_map(&maps.o[0][l], pre->mean_logits, l, 72));
        UDAL_TRY(out_map(&maps.o[1][l], pre->std_logits, l, 72));
      }
      return l2 ? launch_dwf<false, 8, true>(ctx, maps, p, grid) : launch_dwf<false, 8>(ctx, maps, p, grid);
    }
    if (c.num_classes == 7) {
      UDAL_REQUIRE(h.dwh_frows == 80, "fused class head: weight image");
      return l2 ? launch_dwf<false, 7, true>(ctx, maps, p, grid) : launch_dwf<false, 7>(ctx, maps, p, grid);
    }
    UDAL_REQUIRE(c.num_classes == 10 && h.dwh_frows == 96, "fused class head: %d classes not covered", c.num_classes);
    return l2 ? launch_dwf<false, 10, true>(ctx, maps, p, grid) : launch_dwf<false, 10>(ctx, maps, p, grid);
  }
  UDAL_REQUIRE(pre->boxes && pre->albox && pre->mcbox, "fused box head: NULL output");
  UDAL_REQUIRE(h.cout == 72 && h.dwh_frows == 80, "fused box head: 8A = 72 channels");
  p.boxes = pre->boxes;
  p.albox = pre->albox;
  p.mcbox = pre->mcbox;
  for (int l = 0; l < c.num_levels; ++l) {
    UDAL_TRY(out_map(&maps.o[0][l], pre->boxes, l, 36));
    UDAL_TRY(out_map(&maps.o[1][l], pre->albox, l, 36));
    UDAL_TRY(out_map(&maps.o[2][l], pre->mcbox, l, 36));
  }
  return l2 ? launch_dwf<true, 8, true>(ctx, maps, p, grid) : launch_dwf<true, 8>(ctx, maps, p, grid);
}
