// K1, tower layer 1 (bf16 tensor-core mode): the layer whose input - the layer-0 output - does not
// depend on the MC sample.  Persistent, warp specialised, one CTA per SM; work item = (16x8-pixel
// tile, image), and the T samples of the item are produced back to back from ONE depthwise pass:
//
//   warp 0      producer : TMA box (18 x 10 px halo tile of the bf16 layer-0 output, zero OOB fill)
//                          into a 2-stage ring
//   warps 2-5   builders : depthwise 3x3 on the CUDA cores (fp32) -> A operand [128 x 64] bf16 in the
//                          K-major 128B-swizzled UMMA layout (double buffered over items); then, per
//                          sample t, the B operand = resident bf16(W * bn_scale * 1/(1-rate)) image of
//                          the level with the columns of the channels dropped by SpatialDropout2D
//                          (mask r = 0 of sample t) zeroed - a masked 8 KB copy, no arithmetic - into
//                          a 4-deep ring
//   warp 1      MMA      : 4 x tcgen05.mma (M128 N64 K16) per sample, accumulators 4-deep in TMEM
//   warps 6-21  epilogue : 2 groups x 8 warps (lane quarter x column half): tcgen05.ld -> folded bias ->
//                          swish (tanh.approx) -> this layer's keep-scale -> bf16 staging tile -> TMA
//                          tensor store
// (Measured with in-kernel cycle counters: the epilogue arithmetic - ~250 instructions per thread and
// tile - and the depthwise pass bound the kernel, not the tensor pipe or HBM.)
//
// Dropout of the layer-0 output is a per-(sample, image, channel) factor in {0, 1/(1-rate)}; it
// commutes with the depthwise conv, so it lives in the rows of B and the depthwise result is shared
// by all samples (efficientdet_keras.py:448-483 / 628-664, utils_extra.py:142-198).
// Numerics identical to the per-tile kernel it replaces: bf16(w * scale) operands, fp32 accumulation.
#include "udal_common.cuh"
#include "heads_umma.cuh"

namespace {

constexpr int kL1Threads = 64 + 128 + 512;  // producer + MMA, builders, epilogue
constexpr int L1_STAGE = (IG_ROWS * IG_BOXW * 128 + 1023) / 1024 * 1024;  // 23 552
// shared memory map (offsets from a 1024-byte aligned base)
constexpr int L1_A = 0;                          // 2 x [128][128 B] depthwise output (UMMA A)
constexpr int L1_BT = L1_A + 2 * 16384;          // 4 x [64][128 B] masked weights (UMMA B)
constexpr int L1_OUT = L1_BT + 4 * 8192;         // 2 x [128][128 B] staging tiles (one per epilogue group)
constexpr int L1_IN = L1_OUT + 2 * 16384;        // 2 x halo tile, linear [18][10][64] bf16
constexpr int L1_BAR = L1_IN + 2 * L1_STAGE;     // mbarriers + tmem slot (168 B), item-index ring (64 B at +192)
constexpr int L1_DW = L1_BAR + 256;              // [9][64] fp32 depthwise weights
constexpr int L1_SC = L1_DW + 9 * KF * 4;        // 2 x [64] fp32 keep-scales of the tile being stored
constexpr int L1_FB = L1_SC + 2 * KF * 4;        // [levels][64] fp32 folded bias (halved)
constexpr int L1_MSK = L1_FB + UDAL_MAX_LEVELS * KF * 4;  // 2 x [kL1MaxT samples][8] bytes: keep bits of the item's samples (per k chunk)
constexpr int kL1MaxT = 64;
constexpr int L1_WC = (L1_MSK + 2 * kL1MaxT * 8 + 1023) / 1024 * 1024;  // [levels][64][128 B] bf16 weights
constexpr int l1_smem(int levels) { return L1_WC + levels * 8192 + 1024; }
constexpr int kL1MaxLevels = UDAL_MAX_LEVELS;  // resident weight images: 8 KB per pyramid level
static_assert(l1_smem(kL1MaxLevels) <= kIgSmemLimit, "shared-memory budget");

struct L1Params {
  int num_levels, NB, items;             // NB = images; items = sum_l tiles[l] * NB (level major)
  int H[UDAL_MAX_LEVELS], W[UDAL_MAX_LEVELS], tiles_x[UDAL_MAX_LEVELS], tiles[UDAL_MAX_LEVELS];
  int item_off[UDAL_MAX_LEVELS + 1];
  uint32_t tiles_magic[UDAL_MAX_LEVELS], tiles_x_magic[UDAL_MAX_LEVELS];
  const float* wf[UDAL_MAX_LEVELS];      // [64 n][64 k] fp32 pointwise * BN scale
  const float* fb[UDAL_MAX_LEVELS];      // [64] folded bias
  const float* in_scale[UDAL_MAX_LEVELS];   // [T*NB][64] keep-scale of the layer-0 dropout (or ones, sc_stride 0)
  const float* out_scale[UDAL_MAX_LEVELS];  // [T*NB][64] keep-scale of this layer's dropout (or ones)
  const float* dw;                       // [9][64]
  int T, sc_stride;
  float inv_keep;                        // the non-zero value of in_scale
  int* counter;                          // zeroed work-item counter of this launch (dynamic claiming, heads_umma.cuh)
};

struct L1Maps {
  CUtensorMap in[UDAL_MAX_LEVELS];    // [NB,H,W,64] bf16, box {64,10,18,1}, no swizzle
  CUtensorMap out[UDAL_MAX_LEVELS];   // [T*NB,H,W,64] bf16, box {64,8,16,1}, 128B swizzle
};

// FP16: 16-bit format of the activations in / out and of the GEMM operands (false: bf16)
template <bool FP16>
__global__ void __launch_bounds__(kL1Threads, 1) heads_l1_kernel(const __grid_constant__ L1Maps maps, const L1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = s32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t sb = s32(smem);
  // barriers: in_full[2] @0  in_empty[2] @16  a_full[2] @32  a_empty[2] @48  b_full[4] @64  tfull[4] @96
  //           tempty[4] @128  tmem slot @160
  const uint32_t bar0 = sb + L1_BAR;
  const uint32_t in_full = bar0, in_empty = bar0 + 16, a_full = bar0 + 32, a_empty = bar0 + 48, b_full = bar0 + 64,
                 tfull = bar0 + 96, tempty = bar0 + 128;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L1_BAR + 160);
  volatile int* sQ = reinterpret_cast<volatile int*>(smem + L1_BAR + 192);  // item-index ring (IG_QRING ints)
  float* sDw = reinterpret_cast<float*>(smem + L1_DW);
  float* sFb = reinterpret_cast<float*>(smem + L1_FB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      bar_init(in_full + 8 * i, 1);
      bar_init(in_empty + 8 * i, 4);   // one arrival per builder warp (128 per-thread arrivals on one mbarrier
      bar_init(a_full + 8 * i, 4);     // serialise in the shared-memory pipe and cost more than the work)
      bar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      bar_init(b_full + 8 * i, 4);
      bar_init(tfull + 8 * i, 1);
      bar_init(tempty + 8 * i, 8);     // one arrival per epilogue warp of the group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sb + L1_BAR + 160) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident tables: depthwise weights, halved folded bias (x*sigmoid(x) = h*tanh(h) + h with h = x/2) and, per
  // level, the bf16 image of (pointwise * BN scale * 1/(1-rate)) in the swizzled K-major layout
  for (int e = threadIdx.x; e < 9 * KF; e += kL1Threads) sDw[e] = __ldg(p.dw + e);
  for (int e = threadIdx.x; e < p.num_levels * KF; e += kL1Threads) sFb[e] = 0.5f * __ldg(p.fb[e / KF] + (e % KF));
  for (int e = threadIdx.x; e < p.num_levels * KF * 8; e += kL1Threads) {
    const int l = e / (KF * 8), n = (e / 8) % KF, c = e % 8;
    const float4* src = reinterpret_cast<const float4*>(p.wf[l] + (size_t)n * KF + c * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    uint4 v;
    v.x = ig_pack16<FP16>(a.x * p.inv_keep, a.y * p.inv_keep);
    v.y = ig_pack16<FP16>(a.z * p.inv_keep, a.w * p.inv_keep);
    v.z = ig_pack16<FP16>(b.x * p.inv_keep, b.y * p.inv_keep);
    v.w = ig_pack16<FP16>(b.z * p.inv_keep, b.w * p.inv_keep);
    *reinterpret_cast<uint4*>(smem + L1_WC + l * 8192 + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer (warp-uniform loop, one elected lane issues) =====================
    for (int i = 0;; ++i) {
      const int item = ig_claim(p.counter, p.items, lane);
      const IgItem w = ig_item(p, item < 0 ? 0 : item);
      const int s = i & 1;
      if (ig_elect_one()) {
        bar_wait(in_empty + 8 * s, ((i >> 1) & 1) ^ 1);
        sQ[i & (IG_QRING - 1)] = item;  // published by the arrival on the stage's full barrier
        if (item < 0) {
          sQ[(i + 1) & (IG_QRING - 1)] = -1;  // end of the stream, for both epilogue groups
          bar_arrive(in_full + 8 * s);
        } else {
          bar_expect_tx(in_full + 8 * s, IG_ROWS * IG_BOXW * 128);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(sb + L1_IN + s * L1_STAGE), "l"(&maps.in[w.l]), "r"(in_full + 8 * s), "r"(0), "r"(w.tx0 - 1), "r"(w.ty0 - 1),
              "r"(w.nb)
              : "memory");
        }
      }
      __syncwarp();
      if (item < 0) break;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ig_idesc<FP16>(64);
    int j = 0;
    for (int i = 0;; ++i) {
      const int ab = i & 1;
      const uint64_t adesc = ig_desc(sb + L1_A + ab * 16384, 1024, 0);
      if (lane == 0) bar_wait(a_full + 8 * ab, (i >> 1) & 1);        // depthwise output of this item in place
      __syncwarp();
      if (ig_queue_read(sQ, i) < 0) {
        // end of the stream (the builders arrived without an A tile): wake both epilogue groups
        if (ig_elect_one()) {
          for (int dj = 0; dj < 2; ++dj) {
            const int q = (j + dj) & 3;
            bar_wait(tempty + 8 * q, (((j + dj) >> 2) & 1) ^ 1);
            bar_arrive(tfull + 8 * q);
          }
        }
        __syncwarp();
        break;
      }
      for (int t = 0; t < T; ++t, ++j) {
        const int q = j & 3;
        const uint64_t bdesc = ig_desc(sb + L1_BT + q * 8192, 1024, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(q * KF);
        if (ig_elect_one()) {
          bar_wait(b_full + 8 * q, (j >> 2) & 1);                    // masked weights of this sample in place
          bar_wait(tempty + 8 * q, ((j >> 2) & 1) ^ 1);              // accumulator drained
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < KF / 16; ++k)
            ig_mma(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
          ig_commit(tfull + 8 * q);                                  // accumulator ready, B slot reusable
          if (t == T - 1) ig_commit(a_empty + 8 * ab);               // A buffer reusable
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== builders: depthwise once per item, masked weights once per sample =====================
    const int tid = threadIdx.x - 64;
    const int q4 = tid & 15, x = tid >> 4;  // depthwise: channel quad, tile column
    const int kc = tid & 7, n0 = tid >> 3;  // weight copy: 16-byte k chunk, first row
    float wgt[9][4];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      const float4 w4 = *reinterpret_cast<const float4*>(sDw + tp * KF + q4 * 4);
      wgt[tp][0] = w4.x; wgt[tp][1] = w4.y; wgt[tp][2] = w4.z; wgt[tp][3] = w4.w;
    }
    int j = 0;
    for (int i = 0;; ++i) {
      const int s = i & 1, ab = i & 1;
      if (lane == 0) {
        bar_wait(in_full + 8 * s, (i >> 1) & 1);          // halo tile landed
        bar_wait(a_empty + 8 * ab, ((i >> 1) & 1) ^ 1);   // the MMAs of item i-2 are done with this A buffer
      }
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) {  // end of the stream: pass it on to the MMA warp through the A barrier
        if (lane == 0) bar_arrive(a_full + 8 * ab);
        break;
      }
      const IgItem w = ig_item(p, item);
      // keep bits of ALL samples of the item, one byte per (sample, 8-channel chunk), fetched now (one global-load latency
      // per item, under the depthwise pass) - a fetch per sample sat in the serial chain B copy -> MMA -> epilogue
      uint8_t* const sMsk = smem + L1_MSK + ab * (kL1MaxT * 8);
      {
        const float* isc = p.in_scale[w.l];
        for (int t = tid >> 3; t < T; t += 16) {
          const size_t row = ((size_t)t * p.NB + w.nb) * p.sc_stride + kc * 8;
          const float4 m0 = __ldg(reinterpret_cast<const float4*>(isc + row));
          const float4 m1 = __ldg(reinterpret_cast<const float4*>(isc + row + 4));
          sMsk[t * 8 + kc] = (uint8_t)((m0.x != 0.f ? 1u : 0u) | (m0.y != 0.f ? 2u : 0u) | (m0.z != 0.f ? 4u : 0u) | (m0.w != 0.f ? 8u : 0u) |
                                       (m1.x != 0.f ? 16u : 0u) | (m1.y != 0.f ? 32u : 0u) | (m1.z != 0.f ? 64u : 0u) |
                                       (m1.w != 0.f ? 128u : 0u));
        }
      }
      {
        const uint8_t* sIn = smem + L1_IN + s * L1_STAGE;
        uint8_t* sA = smem + L1_A + ab * 16384;
        float acc[IG_TH][4];
#pragma unroll
        for (int y = 0; y < IG_TH; ++y)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[y][c] = 0.f;
#pragma unroll
        for (int r = 0; r < IG_ROWS; ++r) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint2 raw2 = *reinterpret_cast<const uint2*>(sIn + (size_t)(r * IG_BOXW + x + dx) * 128 + q4 * 8);
            const float2 v01 = ig_unpack16<FP16>(raw2.x), v23 = ig_unpack16<FP16>(raw2.y);
            const float v[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int y = r - dy;
              if (y >= 0 && y < IG_TH) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[y][c] = fmaf(v[c], wgt[dy * 3 + dx][c], acc[y][c]);
              }
            }
          }
          if (r >= 2) {
            const int y = r - 2;
            const int m = y * IG_TW + x;
            uint2 o;
            o.x = ig_pack16<FP16>(acc[y][0], acc[y][1]);
            o.y = ig_pack16<FP16>(acc[y][2], acc[y][3]);
            *reinterpret_cast<uint2*>(sA + (size_t)m * 128 + (((q4 >> 1) ^ (m & 7)) << 4) + (q4 & 1) * 8) = o;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A tile -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        bar_arrive(a_full + 8 * ab);
        bar_arrive(in_empty + 8 * s);
      }
      asm volatile("bar.sync 5, 128;" ::: "memory");  // the builders' keep bits of this item are in place
      const uint8_t* sWc = smem + L1_WC + w.l * 8192;
      for (int t = 0; t < T; ++t, ++j) {
        const int q = j & 3;
        // bit masks of this thread's 8 channels: keep-scale is either 0 or 1/(1-rate)
        const uint32_t kb = sMsk[t * 8 + kc];
        uint4 msk;
        msk.x = (kb & 1u ? 0x0000ffffu : 0u) | (kb & 2u ? 0xffff0000u : 0u);
        msk.y = (kb & 4u ? 0x0000ffffu : 0u) | (kb & 8u ? 0xffff0000u : 0u);
        msk.z = (kb & 16u ? 0x0000ffffu : 0u) | (kb & 32u ? 0xffff0000u : 0u);
        msk.w = (kb & 64u ? 0x0000ffffu : 0u) | (kb & 128u ? 0xffff0000u : 0u);
        if (lane == 0) bar_wait(tfull + 8 * q, ((j >> 2) & 1) ^ 1);  // the MMA of sample j-4 is done with this slot
        __syncwarp();
        uint8_t* sBt = smem + L1_BT + q * 8192;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = n0 + 16 * u;
          const uint32_t off = (uint32_t)(n * 128 + ((kc ^ (n & 7)) << 4));
          uint4 v = *reinterpret_cast<const uint4*>(sWc + off);
          v.x &= msk.x; v.y &= msk.y; v.z &= msk.z; v.w &= msk.w;
          *reinterpret_cast<uint4*>(sBt + off) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(b_full + 8 * q);
      }
    }
  } else {
    // ===================== epilogue: 2 groups x 8 warps =====================
    const int ew = warp - 6;
    const int g = ew >> 3;                  // group: handles the samples with (j & 1) == g
    const int wq = warp & 3;                // TMEM lane quarter this warp may access
    const int hc = (ew >> 2) & 1;           // column half
    const int m = wq * 32 + lane;           // GEMM row = pixel (m / 8, m % 8)
    const bool elected = (ew & 7) == 0 && lane == 0;
    float* const sSc = reinterpret_cast<float*>(smem + L1_SC) + g * KF;
    const uint32_t swz = (uint32_t)(m & 7);
    // keep-scales of the group's next sample of the same item are fetched one sample ahead by 16 lanes of one warp: the
    // global-load latency (the whole group would otherwise wait for it at its first barrier) hides under the current sample
    const bool sc_loader = (ew & 7) == 1 && lane < KF / 4;
    float4 sc_next = make_float4(0.f, 0.f, 0.f, 0.f);
    bool have_next = false;
    int i = 0, t = g;  // sample j = g, g + 2, ... of this CTA's item stream is sample t of its i-th item
    for (int j = g;; j += 2, t += 2) {
      while (t >= T) {
        t -= T;
        ++i;
        have_next = false;
      }
      const int q = j & 3;
      if (lane == 0) bar_wait(tfull + 8 * q, (j >> 2) & 1);
      __syncwarp();
      const int item = ig_queue_read(sQ, i);
      if (item < 0) break;
      {
        const IgItem w = ig_item(p, item);
        const float* fbv = sFb + w.l * KF + hc * 32;
        const int nb = t * p.NB + w.nb;
        uint8_t* const ob = smem + L1_OUT + g * 16384;
        // this sample's keep-scales -> the group's slot (published by the first group barrier below; the
        // previous tile's readers passed its second barrier); then the fetch for sample t + 2 of the item
        if (sc_loader) {
          const float* row = p.out_scale[w.l] + (size_t)nb * p.sc_stride;
          reinterpret_cast<float4*>(sSc)[lane] = have_next ? sc_next : __ldg(reinterpret_cast<const float4*>(row) + lane);
          if (t + 2 < T) sc_next = __ldg(reinterpret_cast<const float4*>(row + (size_t)2 * p.NB * p.sc_stride) + lane);
        }
        have_next = t + 2 < T;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(q * KF + hc * 32);
        uint32_t r[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) ig_ld8(taddr + u * 8, r[u]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) bar_arrive(tempty + 8 * q);
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging tile free again
        __syncwarp();
        ig_group_sync256(g);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 f0 = *reinterpret_cast<const float4*>(fbv + u * 8);
          const float4 f1 = *reinterpret_cast<const float4*>(fbv + u * 8 + 4);
          const float4 s0 = *reinterpret_cast<const float4*>(sSc + hc * 32 + u * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(sSc + hc * 32 + u * 8 + 4);
          const float fb8[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          const float sc8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {  // packed pairs (FFMA2 / FMUL2): the same IEEE operations, half the issue slots
            const float2 h = ig_fma2(make_float2(__uint_as_float(r[u][e]), __uint_as_float(r[u][e + 1])), make_float2(0.5f, 0.5f),
                                     make_float2(fb8[e], fb8[e + 1]));
            const float2 sw = ig_mul2(ig_fma2(h, make_float2(ig_tanh(h.x), ig_tanh(h.y)), h), make_float2(sc8[e], sc8[e + 1]));
            v[e] = sw.x;
            v[e + 1] = sw.y;
          }
          uint4 o;
          o.x = ig_pack16<FP16>(v[0], v[1]);
          o.y = ig_pack16<FP16>(v[2], v[3]);
          o.z = ig_pack16<FP16>(v[4], v[5]);
          o.w = ig_pack16<FP16>(v[6], v[7]);
          *reinterpret_cast<uint4*>(ob + m * 128 + (((uint32_t)(hc * 4 + u) ^ swz) << 4)) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        ig_group_sync256(g);
        if (elected) {
          ig_tma_store(&maps.out[w.l], s32(ob), 0, w.tx0, w.ty0, nb);
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

// Tower layer 1 over all pyramid levels: in[l] = layer-0 output [NB,H_l,W_l,64] bf16 (no dropout applied),
// out[l] [T*NB,H_l,W_l,64] bf16 = swish(BN(sepconv(in * in_scale_t))) * out_scale_t for t < T.
// in_scale / out_scale: per level [T*NB][64] keep-scales (values 0 or inv_keep), or null = no dropout
// (`ones`: 64 floats of 1.0 on the device).  wf[l] [64][64] fp32 (pointwise * BN scale, n-major), fb[l] [64].
int udal_heads_l1_layer(udal_ctx* ctx, const void* const* in, int NB, int T, const float* dw, const float* const* wf,
                        const float* const* fb, const float* const* in_scale, const float* const* out_scale,
                        const float* ones, float inv_keep, void* const* out) {
  const bool fp16 = ctx->cfg.heads_mode == UDAL_HEADS_FP16_TC;
  const CUtensorMapDataType dt16 = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  EncodeTiledFn encode = get_encode();
  UDAL_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  UDAL_REQUIRE((in_scale == nullptr) == (out_scale == nullptr), "layer 1: both dropout scale tables or none");
  const udal_config& c = ctx->cfg;
  L1Maps maps;
  L1Params p;
  memset(&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.num_levels = c.num_levels;
  p.NB = NB;
  UDAL_REQUIRE(T >= 1 && T <= kL1MaxT, "tensor-core heads: at most %d MC samples (got %d) - use heads_mode fp32", kL1MaxT, T);
  p.T = T;
  p.sc_stride = in_scale ? KF : 0;
  p.inv_keep = in_scale ? inv_keep : 1.0f;
  p.dw = dw;
  int off = 0;
  for (int l = 0; l < c.num_levels; ++l) {
    const int H = c.level_h[l], W = c.level_w[l];
    UDAL_TRY(encode_nhwc(encode, &maps.in[l], dt16, 2, in[l], NB, H, W, KF, KF, IG_BOXW, IG_ROWS, false));
    UDAL_TRY(encode_nhwc(encode, &maps.out[l], dt16, 2, out[l], T * NB, H, W, KF, KF, IG_TW, IG_TH, true));
    p.H[l] = H;
    p.W[l] = W;
    p.tiles_x[l] = (W + IG_TW - 1) / IG_TW;
    p.tiles[l] = p.tiles_x[l] * ((H + IG_TH - 1) / IG_TH);
    p.tiles_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles[l] - 1) / (uint64_t)p.tiles[l]);
    p.tiles_x_magic[l] = (uint32_t)((0x100000000ull + (uint64_t)p.tiles_x[l] - 1) / (uint64_t)p.tiles_x[l]);
    UDAL_REQUIRE((int64_t)p.tiles[l] * NB * p.tiles[l] < (1ll << 32), "level %d: too many work items for the item decode", l);
    p.item_off[l] = off;
    off += p.tiles[l] * NB;
    p.wf[l] = wf[l];
    p.fb[l] = fb[l];
    p.in_scale[l] = in_scale ? in_scale[l] : ones;
    p.out_scale[l] = out_scale ? out_scale[l] : ones;
  }
  for (int l = c.num_levels; l <= UDAL_MAX_LEVELS; ++l) p.item_off[l] = off;
  p.items = off;
  UDAL_TRY(udal_work_counter(ctx, &p.counter));
  const int grid = udal_persistent_grid(ctx, p.items);
  UDAL_REQUIRE(c.num_levels <= kL1MaxLevels, "the tensor-core head sampler keeps the weights of at most %d pyramid levels "
               "resident (got %d) - use heads_mode fp32", kL1MaxLevels, c.num_levels);
  const int smem = l1_smem(c.num_levels);
  if (fp16) {
    UDAL_CUDA(cudaFuncSetAttribute(heads_l1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    heads_l1_kernel<true><<<grid, kL1Threads, smem, ctx->stream>>>(maps, p);
  } else {
    UDAL_CUDA(cudaFuncSetAttribute(heads_l1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    heads_l1_kernel<false><<<grid, kL1Threads, smem, ctx->stream>>>(maps, p);
  }
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
