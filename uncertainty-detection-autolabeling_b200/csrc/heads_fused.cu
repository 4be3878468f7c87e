// K1 predict layer + K2 fused (bf16 tensor-core mode, serving configuration): the class / box(+sigma)
// predict layers as implicit GEMMs on tcgen05 (as heads_ig.cu) whose epilogue never writes the [T,...]
// head outputs to HBM.  The T samples of a (16x8-pixel tile, image) run back to back on one CTA and the
// epilogue threads keep the Monte-Carlo statistics of their anchors in registers:
//
//   class head : (7 or 8 classes) per (anchor, class) logit the sequential fp32 sum (-> mean, bit-identical to the
//                stand-alone decode kernel), the first sample and the sum of squared deviations from it
//                (-> population std); at the last sample also argmax / sigmoid score per anchor
//                (utils_extra.py:220-244, postprocess.py:123-135, 284)
//   box head   : per sample the decode with exact (closed-form) moment propagation of
//                utils_box.py:125-160 in fp32 (exp through ex2.approx, exp(v) - 1 through a series for
//                small v: ~1e-6 relative, inside the 1e-4 contract; the fp64 form of decode_math.cuh on
//                12 epilogue warps would bound the kernel at 3x the MMA time), then the same running
//                statistics on the decoded corners and the mean of the aleatoric std
//                (postprocess.py:297-331)
//
// Roles: warp 0 TMA producer (4-stage ring of 18x10-pixel halo tiles), warp 1 MMA issuer (36 x
// tcgen05.mma M128 N80 K16 per sample, accumulators double buffered in TMEM), warps 2-13 epilogue:
// 4 TMEM lane quarters x 3 anchor triples; thread = (pixel, 3 anchors).
// Against predict layers + decode_moments on the same activations: mean logits, scores and classes are
// bit-identical; the standard deviations come from a one-pass (shifted) variance (~1e-6 relative) and the
// box quantities from the fp32 decode (~1e-6 relative).
// Per-anchor results leave through shared-memory staging tiles and TMA tensor stores ([B, H_l, W_l, 72] /
// [.., 36] views of the [B, N, C] / [B, N, 4] outputs), which also clip the ragged tile edges.
#include <type_traits>

#include "udal_common.cuh"
#include "heads_umma.cuh"
#include "decode_math.cuh"

namespace {

constexpr int kFuThreads = 64 + 12 * 32;
constexpr int FU_STAGE = (IG_ROWS * IG_BOXW * 128 + 1023) / 1024 * 1024;
constexpr int FU_MAX_STAGES = 4;
// compile-time shape of one kernel variant (offsets from a 1024-byte aligned base)
//   NPAD   UMMA N                       NROWS  weight rows kept per tap (multiple of 8)
//   STAGES TMA ring depth               OUT_BYTES / OUT2_BYTES  the two staging areas
template <int NPAD_, int NROWS_, int STAGES_, int OUT_BYTES_, int OUT2_BYTES_>
struct FuShape {
  static constexpr int NPAD = NPAD_, NROWS = NROWS_, STAGES = STAGES_;
  static constexpr int B = 0;
  static constexpr int B_BYTES = 9 * NROWS * 128;
  static constexpr int IN = B + B_BYTES;
  static constexpr int OUT = IN + STAGES * FU_STAGE;   // class: [128 px][9 NC] fp32; box: boxes | albox, 2 x [128 px][36]
  static constexpr int OUT2 = OUT + OUT_BYTES_;        // class: scores [128][9] fp32, classes [128][9] i32; box: mcbox [128][36]
  static constexpr int BAR = OUT2 + OUT2_BYTES_;       // barriers + tmem slot (128 B)
  static constexpr int TBL = BAR + 128;                // 64 doubles: 2^(j/64) table of exp_fast
  static constexpr int BIAS = TBL + 512;               // [NPAD] fp32 predict bias
  static constexpr int QUEUE = BIAS + NPAD * 4;         // item-index ring (IG_QRING ints)
  static constexpr int SMEM = QUEUE + IG_QRING * 4 + 1024;
  static_assert(B_BYTES % 1024 == 0 && STAGES <= FU_MAX_STAGES && 2 * NPAD <= 256, "layout");
  static_assert(SMEM <= kIgSmemLimit, "shared-memory budget");
};
// A = 9 anchors: 63 / 72 logits (7 / 8 classes) and the 72 box + sigma channels share one shape; 10 classes (90 logits,
// e.g. the BDD100K map) need N = 96, whose 108 KB weight image leaves room for a 2-stage ring
using FuShape72 = FuShape<80, 72, 3, 128 * 72 * 4, 128 * 36 * 4>;
using FuShape96 = FuShape<96, 96, 2, 128 * 90 * 4, 128 * 9 * 8>;
template <bool BOX, int NC>
using FuShapeOf = typename std::conditional<(BOX || NC <= 8), FuShape72, FuShape96>::type;

struct FuParams {
  int num_levels, NB, T, items;          // NB = images; items = sum_l tiles[l] * NB (level major)
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];
  int pix_off[UDAL_MAX_LEVELS + 1];      // prefix of H_l * W_l
  const void* wimg;                      // bf16 [9][NROWS][64] swizzled weight image of the predict layer
  const float* bias;                     // [NPAD]
  const float* anchors;                  // [N,4]
  int* counter;                          // zeroed work-item counter of this launch (dynamic claiming, heads_umma.cuh)
  long long N;                           // anchors per image
  float* mean_logits;                    // class head outputs [NB,N,NC]
  float* std_logits;
  float* scores;                         // [NB,N]
  int32_t* classes;
  float* boxes;                          // box head outputs [NB,N,4]
  float* albox;
  float* mcbox;
};

struct FuMaps {
  CUtensorMap m[UDAL_MAX_LEVELS];        // [T*NB,H,W,64] bf16 layer-(R-1) output, box {64,10,18,1}, 128B swizzle
  // outputs as [NB,H_l,W_l,ch] views of the level's slice of the per-anchor tensors, box {ch,8,16,1}, no swizzle
  CUtensorMap o[3][UDAL_MAX_LEVELS];     // class: mean_logits, std_logits (ch = 72); box: boxes, albox, mcbox (ch = 36)
};

__device__ __forceinline__ void fu_epi_sync() {  // the 384 epilogue threads
  asm volatile("bar.sync 1, 384;" ::: "memory");
}
__device__ __forceinline__ float fu_exp(float x) {  // ex2.approx: ~2^-22 relative
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
__device__ __forceinline__ float fu_expm1(float v) {  // v = sigma^2 >= 0
  if (v < 0.25f) {
    float q = fmaf(v, 1.f / 720.f, 1.f / 120.f);
    q = fmaf(q, v, 1.f / 24.f);
    q = fmaf(q, v, 1.f / 6.f);
    q = fmaf(q, v, 0.5f);
    q = fmaf(q, v, 1.f);
    return q * v;
  }
  return fu_exp(v) - 1.f;
}
__device__ __forceinline__ float fu_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// one axis of utils_box.py:125-160 (l-norm), fp32: sa = anchor size, ca = anchor centre, sa2 = sa * sa
__device__ __forceinline__ void fu_decode_axis(float sa, float ca, float sa2, float t_c, float t_s, float s_c, float s_s,
                                               float& lo, float& hi, float& sd) {
  const float vs = s_s * s_s, vc = s_c * s_c;
  const float c = fmaf(t_c, sa, ca);
  const float e = fu_exp(fmaf(0.5f, vs, t_s));
  const float half = 0.5f * e * sa;
  lo = c - half;
  hi = c + half;
  // Var(centre) + Var(size) / 4 with Var(size) = (exp(v) - 1) exp(2 t + v) sa^2
  sd = fu_sqrt(sa2 * fmaf(0.25f * fu_expm1(vs), e * e, vc));
}

__device__ __forceinline__ void fu_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

// NC = classes per anchor of the class head (7: KITTI map of the reference YAMLs, 8, 10: BDD100K); the box head
// ignores it
template <bool BOX, int NC>
__global__ void __launch_bounds__(kFuThreads, 1) heads_fused_kernel(const __grid_constant__ FuMaps maps, const FuParams p) {
  using S = FuShapeOf<BOX, NC>;
  constexpr int FU_NPAD = S::NPAD, FU_NROWS = S::NROWS, FU_STAGES = S::STAGES, FU_B = S::B, FU_B_BYTES = S::B_BYTES,
                FU_IN = S::IN, FU_OUT = S::OUT, FU_OUT2 = S::OUT2, FU_BAR = S::BAR, FU_TBL = S::TBL, FU_BIAS = S::BIAS;
  volatile int* sQ = nullptr;  // item-index ring, set below
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: full[4] @0  empty[4] @32  tfull[2] @64  tempty[2] @80  bfull @96  tmem slot @104
  const uint32_t bar0 = sb + FU_BAR;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 32, bar_tfull = bar0 + 64, bar_tempty = bar0 + 80, bar_b = bar0 + 96;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FU_BAR + 104);
  double* sTbl = reinterpret_cast<double*>(smem + FU_TBL);
  float* sBias = reinterpret_cast<float*>(smem + FU_BIAS);
  sQ = reinterpret_cast<volatile int*>(smem + S::QUEUE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FU_STAGES; ++i) {
      bar_init(bar_full + 8 * i, 1);
      bar_init(bar_empty + 8 * i, 1);
    }
    bar_init(bar_tfull, 1);
    bar_init(bar_tfull + 8, 1);
    bar_init(bar_tempty, 12);   // one arrival per epilogue warp
    bar_init(bar_tempty + 8, 12);
    bar_init(bar_b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sb + FU_BAR + 104) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) sTbl[threadIdx.x] = kExp2Table[threadIdx.x];
  if (threadIdx.x < FU_NPAD) sBias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer =====================
    if (ig_elect_one()) {
      bar_expect_tx(bar_b, FU_B_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + FU_B),
                   "l"(p.wimg), "r"(FU_B_BYTES), "r"(bar_b)
                   : "memory");
    }
    __syncwarp();
    int s = 0, ph = 0;
    for (int i = 0;; ++i) {
      const int item = ig_claim(p.counter, p.items, lane);
      if (item < 0) {  // end of the stream: an empty "sample" carries it to the MMA warp
        if (ig_elect_one()) {
          bar_wait(bar_empty + 8 * s, ph ^ 1);
          sQ[i & (IG_QRING - 1)] = -1;
          bar_arrive(bar_full + 8 * s);
        }
        __syncwarp();
        break;
      }
      const IgItem w = ig_item(p, item);
      for (int t = 0; t < T; ++t) {
        if (ig_elect_one()) {
          bar_wait(bar_empty + 8 * s, ph ^ 1);
          if (t == 0) sQ[i & (IG_QRING - 1)] = item;  // published by the arrival on the first sample's full barrier
          bar_expect_tx(bar_full + 8 * s, IG_ROWS * IG_BOXW * 128);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(sb + FU_IN + s * FU_STAGE), "l"(&maps.m[w.l]), "r"(bar_full + 8 * s), "r"(0), "r"(w.tx0 - 1), "r"(w.ty0 - 1),
              "r"(t * p.NB + w.nb)
              : "memory");
        }
        __syncwarp();
        if (++s == FU_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FU_NPAD >> 3) << 17) | ((128u >> 4) << 24);
    if (lane == 0) bar_wait(bar_b, 0);
    __syncwarp();
    int j = 0, s = 0, ph = 0;
    for (int i = 0;; ++i) {
      if (lane == 0) bar_wait(bar_full + 8 * s, ph);  // first sample of the item (or the end marker) landed
      __syncwarp();
      if (ig_queue_read(sQ, i) < 0) {
        if (ig_elect_one()) {  // wake the epilogue: its next accumulator "arrives" empty
          bar_wait(bar_tempty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);
          bar_arrive(bar_tfull + 8 * (j & 1));
        }
        __syncwarp();
        break;
      }
      for (int t = 0; t < T; ++t, ++j) {
        const int a = j & 1;
        const uint32_t in0 = sb + FU_IN + s * FU_STAGE;
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * FU_NPAD);
        if (ig_elect_one()) {
          bar_wait(bar_tempty + 8 * a, ((j >> 1) & 1) ^ 1);
          bar_wait(bar_full + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint64_t adesc = ig_desc(in0 + (uint32_t)((dy * IG_BOXW + dx) * 128), IG_BOXW * 128, 0);
            const uint64_t bdesc = ig_desc(sb + FU_B + tap * FU_NROWS * 128, 1024, 0);
#pragma unroll
            for (int k = 0; k < KF / 16; ++k)
              ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
          }
          ig_commit(bar_empty + 8 * s);
          ig_commit(bar_tfull + 8 * a);
        }
        __syncwarp();
        if (++s == FU_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: thread = (pixel, anchors 3 cg .. 3 cg + 2) =====================
    const int ew = warp - 2;
    const int etid = threadIdx.x - 64;
    const int q = warp & 3;         // TMEM lane quarter
    const int cg = ew >> 2;         // anchor triple
    const int m = q * 32 + lane;    // pixel (m / 8, m % 8) of the tile
    const bool elected = etid == 0;
    const float fT = (float)T, rT = 1.f / fT;
    float* const sOut = reinterpret_cast<float*>(smem + FU_OUT);
    float* const sOut2 = reinterpret_cast<float*>(smem + FU_OUT2);
    int j = 0;
    for (int i = 0;; ++i) {
      if (lane == 0) bar_wait(bar_tfull + 8 * (j & 1), (j >> 1) & 1);  // first sample of the item (or the end marker)
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) break;
      const IgItem w = ig_item(p, item);
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
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * FU_NPAD + cg * CH);
          constexpr int NLD = (CH + 7) / 8;
          static_assert(2 * CH + NLD * 8 <= FU_NPAD, "accumulator columns read past the thread's logits stay inside N");
          if constexpr (NC > 8) {
            // 30 logits x (sum, first sample, squared deviations) already fill the register file: the accumulator is
            // read 8 columns at a time (the MMA of the next sample takes ~2000 cycles, the extra load latency hides)
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
                    const float2 q = ig_fma2(d, d, make_float2(s2[c], s2[c + 1]));
                    sum[c] = sm.x; sum[c + 1] = sm.y;
                    s2[c] = q.x; s2[c + 1] = q.y;
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
                  const float2 q = ig_fma2(d, d, make_float2(s2[c], s2[c + 1]));
                  sum[c] = sm.x; sum[c + 1] = sm.y;
                  s2[c] = q.x; s2[c + 1] = q.y;
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
        fu_epi_sync();
        float* st = sOut + m * ROW + cg * CH;
        float mean[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) mean[c] = __fdiv_rn(sum[c], fT);
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
          sOut2[m * 9 + cg * 3 + ai] = sigmoid_ref(best);
          reinterpret_cast<int32_t*>(sOut2 + 128 * 9)[m * 9 + cg * 3 + ai] = arg;
        }
        // a tile row (8 px) of a per-anchor tensor with `w` values per pixel is one contiguous run in global memory
        auto copy_rows = [&](const float* src, float* dst_base, int width) {
          for (int idx = etid; idx < IG_TH * IG_TW * width; idx += 384) {
            const int row = idx / (IG_TW * width), col = idx - row * (IG_TW * width);
            if (w.ty0 + row < H && w.tx0 + col / width < W)
              dst_base[((size_t)w.nb * (size_t)(p.N / 9) + (size_t)(p.pix_off[w.l] + (w.ty0 + row) * W + w.tx0)) * width + col] = src[idx];
          }
        };
        if constexpr (TMA_OUT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fu_epi_sync();
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
        fu_epi_sync();  // the mean tile has been read
        // population std, shifted one-pass form: s1 = sum of the deviations from the first sample
        if constexpr (CH % 4 == 0 && ROW % 4 == 0) {
#pragma unroll
          for (int v = 0; v < CH / 4; ++v) {
            float sd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * v + e;
              const float s1 = sum[c] - fT * x0[c];
              sd[e] = fu_sqrt(fmaxf(fmaf(-s1 * rT, s1, s2[c]), 0.f) * rT);
            }
            reinterpret_cast<float4*>(st)[v] = make_float4(sd[0], sd[1], sd[2], sd[3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float s1 = sum[c] - fT * x0[c];
            st[c] = fu_sqrt(fmaxf(fmaf(-s1 * rT, s1, s2[c]), 0.f) * rT);
          }
        }
        if constexpr (TMA_OUT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fu_epi_sync();
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
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * FU_NPAD + cg * 12);
          uint32_t rt8[8], rt4[4], rs8[8], rs4[4];
          ig_ld8(taddr, rt8);
          fu_ld4(taddr + 8, rt4);
          ig_ld8(taddr + 36, rs8);
          fu_ld4(taddr + 44, rs4);
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
              fu_decode_axis(sa[k], ca[k], sa2[k], tv[ai * 4 + ax], tv[ai * 4 + 2 + ax], sv[ai * 4 + ax], sv[ai * 4 + 2 + ax], lo, hi,
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
        fu_epi_sync();
        float4* stb = reinterpret_cast<float4*>(sOut + m * 36 + cg * 12);
        float4* sta = reinterpret_cast<float4*>(sOut + 128 * 36 + m * 36 + cg * 12);
        float4* stm = reinterpret_cast<float4*>(sOut2 + m * 36 + cg * 12);
#pragma unroll
        for (int ai = 0; ai < 3; ++ai) {
          const int ky = ai * 2, kx = ai * 2 + 1;
          stb[ai] = make_float4(__fdiv_rn(sum_lo[ky], fT), __fdiv_rn(sum_lo[kx], fT), __fdiv_rn(sum_hi[ky], fT),
                                __fdiv_rn(sum_hi[kx], fT));
          const float ay = __fdiv_rn(al[ky], fT), ax2 = __fdiv_rn(al[kx], fT);
          sta[ai] = make_float4(ay, ax2, ay, ax2);
          stm[ai] = make_float4(fu_sqrt(fmaxf(fmaf(-s1_lo[ky] * rT, s1_lo[ky], s2_lo[ky]), 0.f) * rT),
                                fu_sqrt(fmaxf(fmaf(-s1_lo[kx] * rT, s1_lo[kx], s2_lo[kx]), 0.f) * rT),
                                fu_sqrt(fmaxf(fmaf(-s1_hi[ky] * rT, s1_hi[ky], s2_hi[ky]), 0.f) * rT),
                                fu_sqrt(fmaxf(fmaf(-s1_hi[kx] * rT, s1_hi[kx], s2_hi[kx]), 0.f) * rT));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fu_epi_sync();
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace

int udal_run_fused = 1;  // 0: udal_run always goes through predict layers + decode_moments

// true if the fused predict + decode kernels cover this configuration (the serving default)
int udal_heads_fused_ok(const udal_ctx* ctx) {
  const udal_config& c = ctx->cfg;
  return udal_run_fused && (c.heads_mode == UDAL_HEADS_BF16_TC || c.heads_mode == UDAL_HEADS_FP16_TC) && c.repeats >= 2 && c.num_filters == KF && c.anchors_per_loc == 9 &&
         (c.num_classes == 8 || c.num_classes == 7 || c.num_classes == 10) && c.loss_attenuation && c.decode_method == UDAL_DECODE_LNORM && c.cls_mc && c.box_mc &&
         c.max_nms_inputs == 0 && c.mc_samples >= 2;
}

// predict layer of one head fused with the MC moments (class) / decode + MC moments (box):
// in[l] = last tower layer output [T*NB,H_l,W_l,64] bf16, wimg = 72-row weight image, bias [80];
// writes the per-anchor tensors of `pre` that belong to the head.
int udal_heads_fused_predict(udal_ctx* ctx, int head, const void* const* in, int NB, int T, const void* wimg, int rows,
                             const float* bias, const udal_prenms_out* pre) {
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const bool wide = head == UDAL_HEAD_CLASS && ctx->cfg.num_classes > 8;  // 90 logits per pixel: the N = 96 shape
  UDAL_REQUIRE(rows == (wide ? FuShape96::NROWS : FuShape72::NROWS), "fused predict kernel: weight image with %d rows per tap", rows);
  UDAL_REQUIRE(ctx->anchors_set, "anchor table not set");
  const udal_config& c = ctx->cfg;
  FuMaps maps;
  FuParams p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.num_levels = c.num_levels;
  p.NB = NB;
  p.T = T;
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    UDAL_TRY(encode_nhwc(encode, &maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, in[l], T * NB, H, W, KF, KF, IG_BOXW, IG_ROWS,
                         true));
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
  for (int l = 0; l <= UDAL_MAX_LEVELS; ++l) p.pix_off[l] = (int)ctx->level_pix_off[l < c.num_levels ? l : c.num_levels];
  p.items = off;
  p.wimg = wimg;
  p.bias = bias;
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
      for (int l = 0; l < c.num_levels; ++l) {
        UDAL_TRY(out_map(&maps.o[0][l], pre->mean_logits, l, 72));
        UDAL_TRY(out_map(&maps.o[1][l], pre->std_logits, l, 72));
      }
      UDAL_CUDA(cudaFuncSetAttribute(heads_fused_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuShape72::SMEM));
      heads_fused_kernel<false, 8><<<grid, kFuThreads, FuShape72::SMEM, ctx->stream>>>(maps, p);
    } else if (c.num_classes == 7) {
      UDAL_CUDA(cudaFuncSetAttribute(heads_fused_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuShape72::SMEM));
      heads_fused_kernel<false, 7><<<grid, kFuThreads, FuShape72::SMEM, ctx->stream>>>(maps, p);
    } else {
      UDAL_REQUIRE(c.num_classes == 10, "fused class head: %d classes not covered", c.num_classes);
      UDAL_CUDA(cudaFuncSetAttribute(heads_fused_kernel<false, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuShape96::SMEM));
      heads_fused_kernel<false, 10><<<grid, kFuThreads, FuShape96::SMEM, ctx->stream>>>(maps, p);
    }
  } else {
    UDAL_REQUIRE(pre->boxes && pre->albox && pre->mcbox, "fused box head: NULL output");
    p.boxes = pre->boxes;
    p.albox = pre->albox;
    p.mcbox = pre->mcbox;
    for (int l = 0; l < c.num_levels; ++l) {
      UDAL_TRY(out_map(&maps.o[0][l], pre->boxes, l, 36));
      UDAL_TRY(out_map(&maps.o[1][l], pre->albox, l, 36));
      UDAL_TRY(out_map(&maps.o[2][l], pre->mcbox, l, 36));
    }
    UDAL_CUDA(cudaFuncSetAttribute(heads_fused_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuShape72::SMEM));
    heads_fused_kernel<true, 8><<<grid, kFuThreads, FuShape72::SMEM, ctx->stream>>>(maps, p);
  }
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
