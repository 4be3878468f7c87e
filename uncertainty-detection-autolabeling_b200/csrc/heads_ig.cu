// K1, tower layers >= 2 and the predict layers (bf16 tensor-core mode): the separable 3x3 conv as ONE
// implicit GEMM on tcgen05 - the depthwise stage is folded into the dense contraction,
//
//     out[m][n] = sum_tap sum_k  in[pix(m) + tap][k] * ( dw[tap][k] * pw[k][n] * bn_scale[n] )
//
// so the CUDA cores do no convolution arithmetic at all.  Persistent, warp specialised, one CTA
// per SM:
//   warp 0  producer : ONE TMA box per work item (cp.async.bulk.tensor.4d, 128B swizzle, zero OOB
//                      fill = SAME padding): the 18x10-pixel halo tile, row pitch 10 px, into a
//                      3-stage ring; loads the 9 x [N x 64] weight image once, resident after that
//   warp 1  MMA      : one lane issues 36 x tcgen05.mma (M128, N = 64|80, K16) per tile; the A
//                      operand of tap (dy,dx) is the SAME shared-memory tile addressed through a
//                      UMMA descriptor whose start is shifted by (dy*pitch + dx) pixel rows and whose
//                      8-row group stride (SBO) is the row pitch; the 128B swizzle is a function of
//                      the absolute shared-memory address, so neither needs 1 KB alignment;
//                      accumulators double buffered in TMEM; tcgen05.commit releases the smem
//                      stage and publishes the accumulator
//   warps 2-9 epilogue: 2 groups x 4 warps; tcgen05.ld -> folded bias -> swish (tanh.approx) -> this
//                      layer's SpatialDropout2D keep-scale -> bf16 (or fp32 predictions) into a
//                      swizzled shared-memory staging tile -> TMA tensor store (coalesced, clips the
//                      ragged right / bottom edge); per-thread row stores only when the channel
//                      count breaks TMA's 16-byte stride rule (C = 7: 63 channels)
// Work item = (16x8-pixel tile, (sample,image)); items are strided over the CTAs.
//
// Reference arithmetic replaced: efficientdet_keras.py:448-483 / 628-664 (_conv_bn_act and the
// predict SeparableConv2D) for repeats >= 2; numerics as heads_tc.cu (bf16 operands, fp32 accum).
#include <cuda.h>
#include <cuda_bf16.h>

#include "udal_common.cuh"
#include "heads_umma.cuh"

namespace {

constexpr int kIgThreads = 320;       // producer warp, MMA warp, 2 x 4 epilogue warps
constexpr int kIgMaxStages = 3;

// compile-time shape of one kernel variant
//   NPAD   UMMA N (64 | 80)           NROWS  weight rows kept per tap (the MMA reads NPAD rows; rows
//   STAGES TMA ring depth                    >= NROWS alias the next tap and feed unused columns)
//   PREDICT epilogue flavour (compile time: no per-element branches)
template <int NPAD_, int NROWS_, int STAGES_, bool PREDICT_>
struct IgShape {
  static constexpr int NPAD = NPAD_, NROWS = NROWS_, PITCH = IG_BOXW, STAGES = STAGES_;
  static constexpr bool PREDICT = PREDICT_;  // false: tower layer (swish, dropout scale, bf16 out); true: fp32 predictions
  static constexpr int B_BYTES = 9 * NROWS * 128;
  static constexpr int STAGE_BYTES = IG_ROWS * PITCH * 128;
  static constexpr int STAGE_STRIDE = (STAGE_BYTES + 1023) / 1024 * 1024;
  // staging tile of one epilogue group: bf16 [128][64] (16 KB) or fp32 as 32-channel regions of
  // [128][32] (16 KB each, 128B swizzle) plus a dense remainder region [128][NROWS - 64]
  static constexpr int OUT_BYTES = !PREDICT ? 16384 : (NPAD == 64 ? 32768 : 32768 + 128 * (NROWS - 64) * 4);
  static constexpr int SM_B = 0;
  static constexpr int SM_IN = SM_B + B_BYTES;
  static constexpr int SM_OUT = SM_IN + STAGES * STAGE_STRIDE;
  static constexpr int SM_BAR = SM_OUT + 2 * OUT_BYTES;
  static constexpr int SM_FBS = SM_BAR + 640;  // barriers (96 B) + 2 x 64 keep-scales
  static constexpr int smem(int levels) { return SM_FBS + levels * 2 * NPAD * 4 + IG_QRING * 4 + 1024; }
  static_assert(B_BYTES % 1024 == 0 && OUT_BYTES % 1024 == 0, "swizzled regions must stay 1 KB aligned");
  static_assert(STAGES <= kIgMaxStages, "barrier slots");
};

struct IgParams {
  int num_levels, NB, items;             // items = sum_l tiles[l] * NB
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];     // prefix of tiles[l] * NB (items are level major)
  void* out[UDAL_MAX_LEVELS];            // [NB,H,W,64] bf16 or [NB,H,W,Cout] fp32
  const float* out_scale[UDAL_MAX_LEVELS];  // [NB,64] keep-scale of THIS layer's dropout (sc_stride 64) or ones (0)
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];  // ceil(2^32 / d) for the item decode
  int sc_stride;
  const float* ep_scale[UDAL_MAX_LEVELS];   // [NPAD] per-level BN scale (1 for the predict layer)
  const float* ep_bias[UDAL_MAX_LEVELS];    // [NPAD] folded bias
  const void* wimg;                      // bf16 [9][NROWS][64] pre-swizzled smem image (level independent)
  int Cout;
  int ch_off, ch_total;                  // predictions: this launch writes channels [ch_off, ch_off + Cout) of ch_total
  int tma_store;                         // predictions through the staging tile + TMA store (Cout % 4 == 0)
  int debug;                             // timing experiments only (wrong results): 1 = one tap, 2 = no epilogue math / stores, 4 = no TMA loads after the first ring fill
  int* counter;                          // zeroed work-item counter of this launch (dynamic claiming, heads_umma.cuh)
};

struct IgMaps {
  CUtensorMap m[UDAL_MAX_LEVELS];   // input halo boxes
  CUtensorMap o[UDAL_MAX_LEVELS];   // output: bf16 [64 ch, 8, 16] box, or fp32 32-channel regions
  CUtensorMap o2[UDAL_MAX_LEVELS];  // output: fp32 remainder region (Cout % 32 channels)
};

template <class S>
__global__ void __launch_bounds__(kIgThreads, 1) heads_ig_kernel(const __grid_constant__ IgMaps maps,
                                                                 const IgParams p) {
  constexpr int NPAD = S::NPAD, NROWS = S::NROWS, PITCH = S::PITCH, STAGES = S::STAGES;
  constexpr uint32_t kTmemCols = NPAD <= 64 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: full[3] @0  empty[3] @24  tfull[2] @48  tempty[2] @64  bfull @80  tmem slot @88
  const uint32_t bar0 = sb + S::SM_BAR;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 24, bar_tfull = bar0 + 48, bar_tempty = bar0 + 64, bar_b = bar0 + 80;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::SM_BAR + 88);
  float* sFb = reinterpret_cast<float*>(smem + S::SM_FBS);
  volatile int* sQ = reinterpret_cast<volatile int*>(smem + S::SM_FBS + p.num_levels * 2 * NPAD * 4);  // item-index ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(bar_full + 8 * i, 1);
      bar_init(bar_empty + 8 * i, 1);
    }
    bar_init(bar_tfull, 1);
    bar_init(bar_tfull + 8, 1);
    bar_init(bar_tempty, 4);      // one arrival per epilogue warp: per-thread arrivals on one mbarrier serialise
    bar_init(bar_tempty + 8, 4);  // in the shared-memory pipe, which the MMA operand fetch needs
    bar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + IG_QRING) sQ[threadIdx.x - 64] = -1;  // (slots are peeked ahead of their time)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + S::SM_BAR + 88),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue constants; with swish the 0.5 of x*sigmoid(x) = h*tanh(h) + h, h = x/2, is folded in
  const float half = S::PREDICT ? 1.0f : 0.5f;
  for (int e = threadIdx.x; e < p.num_levels * NPAD; e += kIgThreads) {
    const int l = e / NPAD, n = e - l * NPAD;
    sFb[(2 * l) * NPAD + n] = half * __ldg(p.ep_scale[l] + n);
    sFb[(2 * l + 1) * NPAD + n] = half * __ldg(p.ep_bias[l] + n);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int G = gridDim.x;

  if (warp == 0) {
    // ===================== producer =====================
    // (all three roles keep warp-uniform control flow: the TMA / MMA instructions take their operands from
    // the warp's uniform registers, so a lane that has left a divergent region must never run ahead of
    // the lane still inside it - every elected region ends in __syncwarp)
    if (ig_elect_one()) {
      bar_expect_tx(bar_b, S::B_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       sb + S::SM_B),
                   "l"(p.wimg), "r"(S::B_BYTES), "r"(bar_b)
                   : "memory");
    }
    __syncwarp();
    int s = 0, ph = 0, n_loaded = 0;
    for (int k = 0;; ++k) {
      const int item = ig_claim(p.counter, p.items, lane);
      const IgItem w = ig_item(p, item < 0 ? 0 : item);
      const uint32_t dst = sb + S::SM_IN + s * S::STAGE_STRIDE;
      const bool skip = (p.debug & 4) && n_loaded++ >= STAGES;
      if (ig_elect_one()) {
        bar_wait(bar_empty + 8 * s, ph ^ 1);  // stage free (first round passes immediately)
        sQ[k & (IG_QRING - 1)] = item;        // published by the arrival on the stage's full barrier
        if (item < 0) {
          sQ[(k + 1) & (IG_QRING - 1)] = -1;  // end of the stream, for both epilogue groups
          bar_arrive(bar_full + 8 * s);
        } else if (skip) {
          bar_arrive(bar_full + 8 * s);
        } else {
          bar_expect_tx(bar_full + 8 * s, IG_ROWS * IG_BOXW * 128);
          // the whole 18 x 10 pixel halo tile as one box; out-of-image pixels arrive as zeros
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(dst), "l"(&maps.m[w.l]), "r"(bar_full + 8 * s), "r"(0), "r"(w.tx0 - 1), "r"(w.ty0 - 1), "r"(w.nb)
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
    // The whole warp runs this loop: with warp-uniform control flow the descriptors stay in uniform
    // registers and each tcgen05.mma is a single UTCHMMA (a lane-divergent `if (lane == 0)` makes the
    // compiler wrap every instruction in a register-to-uniform waterfall loop, ~70 cycles per MMA
    // instead of the 48 the operand fetch needs).  One elected lane issues.
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
    if (lane == 0) bar_wait(bar_b, 0);  // weights resident
    __syncwarp();
    int s = 0, ph = 0;
    for (int it = 0;; ++it) {
      const int a = it & 1;
      const uint32_t in0 = sb + S::SM_IN + s * S::STAGE_STRIDE;
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * NPAD);
      if (lane == 0) bar_wait(bar_full + 8 * s, ph);  // halo tile landed (one poller: 32 would crowd the epilogue's smem traffic)
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
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if ((p.debug & 1) && tap != 0) continue;
          const int dy = tap / 3, dx = tap % 3;
          // the swizzle XOR is a function of the absolute shared-memory address bits, so a start
          // shifted by whole 128-byte rows needs no base offset (verified on hardware)
          const uint64_t adesc = ig_desc(in0 + (uint32_t)((dy * PITCH + dx) * 128), PITCH * 128, 0);
          const uint64_t bdesc = ig_desc(sb + S::SM_B + tap * NROWS * 128, 1024, 0);
#pragma unroll
          for (int k = 0; k < KF / 16; ++k)
            ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
        }
        ig_commit(bar_empty + 8 * s);   // smem stage reusable once these MMAs retire
        ig_commit(bar_tfull + 8 * a);   // accumulator ready
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ===================== epilogue: 2 groups x 4 warps, group g drains accumulator g =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // GEMM row = pixel (m / 8, m % 8) of the tile
    const bool elected = ((warp - 2) & 3) == 0 && lane == 0;
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
      // this item's dropout keep-scales -> this group's slot in shared memory (the previous tile's math of
      // the group ended before its second group barrier; the first barrier below publishes the slot)
      float* const sSc = reinterpret_cast<float*>(smem + S::SM_BAR + 96) + g * KF;  // inside the barrier block: 2 x 256 B
      if constexpr (!S::PREDICT) {
        if (((warp - 2) & 3) == 1 && lane < KF / 4) {  // (predicated, no divergence)
          const float4* sc = reinterpret_cast<const float4*>(p.out_scale[w.l] + (size_t)nb * p.sc_stride);
          reinterpret_cast<float4*>(sSc)[lane] = item == pre_item ? pre_sc : __ldg(sc + lane);
          // the group's next item (it + 2) is usually in the ring already (the producer runs a few items ahead): fetch its
          // keep-scales now, verified against the ring when the item is really due (a stale slot only costs a wasted load)
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
      if (p.debug & 2) {
        // timing experiment: no epilogue arithmetic, no stores
      } else if constexpr (!S::PREDICT) {
        // ---- tower layer: BN scale + folded bias (both halved) -> swish -> dropout keep-scale -> bf16 ----
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
          o.x = ig_pack(v[0], v[1]);
          o.y = ig_pack(v[2], v[3]);
          o.z = ig_pack(v[4], v[5]);
          o.w = ig_pack(v[6], v[7]);
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
        // ---- predictions whose channel count breaks TMA's 16-byte stride rule (C = 7: 63 channels): dense
        //      staging tile [128 px][Cout], then every tile row - one contiguous run of 8 px x Cout floats
        //      in global memory - is copied out with coalesced 4-byte stores ----
        const int H = p.H[w.l], W = p.W[w.l], Cout = p.Cout;
        float* stg = reinterpret_cast<float*>(ob);
        ig_group_sync(g);  // the previous tile's copy loop of this group is done with the staging tile
#pragma unroll
        for (int j = 0; j < NPAD / 8; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (j * 8 + i < Cout) stg[m * Cout + j * 8 + i] = fmaf(__uint_as_float(r[j][i]), ep_s[j * 8 + i], ep_b[j * 8 + i]);
        ig_group_sync(g);
        const int gt = threadIdx.x - 64 - g * 128;           // thread of the group
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
        if (elected && !(p.debug & 2)) {
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

// weight image: wimg[tap][n][k] = bf16( dw[tap][k] * wf[n][k] ) in the swizzled shared-memory layout,
// nrows rows per tap
__global__ void build_ig_weights_kernel(const float* __restrict__ dw, const float* __restrict__ wf, int nrows,
                                        __nv_bfloat16* __restrict__ wimg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * nrows * KF) return;
  const int k = i % KF, n = (i / KF) % nrows, tap = i / (KF * nrows);
  const float v = dw[tap * KF + k] * wf[(size_t)n * KF + k];
  const size_t byte = (size_t)tap * nrows * 128 + (size_t)n * 128 + (size_t)((((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
  wimg[byte / 2] = __float2bfloat16_rn(v);
}

template <class S>
int launch_ig(udal_ctx* ctx, const IgMaps& maps, const IgParams& p, int grid) {
  const int smem = S::smem(p.num_levels);
  UDAL_REQUIRE(smem <= kIgSmemLimit, "implicit-GEMM head kernel needs %d bytes of shared memory", smem);
  UDAL_CUDA(cudaFuncSetAttribute(heads_ig_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  heads_ig_kernel<S><<<grid, kIgThreads, smem, ctx->stream>>>(maps, p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// kernel variants: <NPAD, NROWS, STAGES, PREDICT>
using IgTower = IgShape<64, 64, 3, false>;    // tower layers
using IgPred64 = IgShape<64, 64, 3, true>;    // predict layers with <= 64 channels
using IgPred72 = IgShape<80, 72, 3, true>;    // predict layers with 65..72 channels (A = 9, C = 8; box+sigma)
using IgPred80 = IgShape<80, 80, 2, true>;    // predict layers with 73..80 channels
static_assert(IgTower::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit && IgPred64::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit &&
              IgPred80::smem(UDAL_MAX_LEVELS) <= kIgSmemLimit && IgPred72::smem(5) <= kIgSmemLimit, "shared-memory budget");

}  // namespace

// debug switches (ctypes-visible)
int udal_ig_tma_store = 1;  // 0: predictions by per-thread row stores even when TMA's stride rule holds
int udal_ig_debug = 0;      // timing experiments (IgParams::debug)

// weight rows per tap the ig kernels expect for a layer with `cout` output channels
int udal_heads_ig_rows(int cout, int num_levels) {
  if (cout <= 64) return 64;
  return (cout <= 72 && IgPred72::smem(num_levels) <= kIgSmemLimit) ? 72 : 80;
}

// builds the swizzled weight image of one layer: out must hold 9*nrows*64 bf16
int udal_heads_ig_build_weights(udal_ctx* ctx, const float* dw, const float* wf, int nrows, void* wimg) {
  const int total = 9 * nrows * KF;
  build_ig_weights_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(dw, wf, nrows, reinterpret_cast<__nv_bfloat16*>(wimg));
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// one tower (>= 2) or predict layer over ALL pyramid levels: in[l] [NB,H_l,W_l,64] bf16 (dropout already applied).
// predict = 0: out[l] bf16 [NB,H,W,64] = swish(BN(conv)) * out_scale (out_scale null: no dropout; `ones` = 64 floats
// of 1.0 on the device); predict = 1: out[l] fp32 [NB,H,W,cout] = conv + bias.
int udal_heads_ig_layer(udal_ctx* ctx, const void* const* in, int NB, const void* wimg, int rows,
                        const float* const* ep_scale, const float* const* ep_bias, int npad, int cout, int predict,
                        const float* const* out_scale, const float* ones, void* const* out, int ch_off, int ch_total) {
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  UDAL_REQUIRE(rows == udal_heads_ig_rows(cout, ctx->cfg.num_levels), "weight image was built for another kernel variant");
  if (ch_total == 0) ch_total = cout;
  UDAL_REQUIRE(ch_off >= 0 && ch_off + cout <= ch_total && (predict || ch_total == cout), "bad channel chunk");
  const udal_config& c = ctx->cfg;
  IgMaps maps;
  IgParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.num_levels = c.num_levels;
  p.NB = NB;
  // TMA needs 16-byte global strides: every bf16 layer qualifies, fp32 predictions when Cout % 4 == 0
  p.tma_store = (!predict || (udal_ig_tma_store && (cout & 3) == 0 && ch_total == cout)) ? 1 : 0;
  p.ch_off = ch_off;
  p.ch_total = ch_total;
  p.sc_stride = out_scale ? KF : 0;
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    UDAL_TRY(encode_nhwc(encode, &maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, in[l], NB, H, W, KF, KF, IG_BOXW, IG_ROWS,
                         true));
    if (!predict) {
      UDAL_TRY(encode_nhwc(encode, &maps.o[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out[l], NB, H, W, KF, KF, IG_TW, IG_TH, true));
    } else if (p.tma_store) {
      const int nfull = cout / 32, rem = cout % 32;
      if (nfull)
        UDAL_TRY(encode_nhwc(encode, &maps.o[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out[l], NB, H, W, cout, 32, IG_TW, IG_TH,
                             true));
      if (rem)
        UDAL_TRY(encode_nhwc(encode, &maps.o2[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out[l], NB, H, W, cout, rem, IG_TW,
                             IG_TH, false));
    }
    p.H[l] = H;
    p.W[l] = W;
    p.tiles_x[l] = (W + IG_TW - 1) / IG_TW;
    p.tiles[l] = p.tiles_x[l] * ((H + IG_TH - 1) / IG_TH);
    p.tiles_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles[l] - 1) / (uint64_t)p.tiles[l]);
    p.tiles_x_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles_x[l] - 1) / (uint64_t)p.tiles_x[l]);
    p.item_off[l] = off;
    UDAL_REQUIRE((int64_t)p.tiles[l] * NB * p.tiles[l] < (1ll << 32), "level %d: too many work items for the item decode", l);
    off += p.tiles[l] * NB;
    p.out[l] = out[l];
    p.out_scale[l] = out_scale ? out_scale[l] : ones;
    p.ep_scale[l] = ep_scale[l];
    p.ep_bias[l] = ep_bias[l];
  }
  for (int l = c.num_levels; l <= UDAL_MAX_LEVELS; ++l) p.item_off[l] = off;
  p.items = off;
  p.Cout = cout;
  p.wimg = wimg;
  p.debug = udal_ig_debug;
  UDAL_TRY(udal_work_counter(ctx, &p.counter));
  const int grid = udal_persistent_grid(ctx, p.items);
  if (!predict) return launch_ig<IgTower>(ctx, maps, p, grid);
  if (npad == 64) return launch_ig<IgPred64>(ctx, maps, p, grid);
  return rows == 72 ? launch_ig<IgPred72>(ctx, maps, p, grid) : launch_ig<IgPred80>(ctx, maps, p, grid);
}
