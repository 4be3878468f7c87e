// Device / host helpers shared by the persistent tcgen05 head kernels (heads_ig.cu, heads_l1.cu): mbarrier,
// TMA, UMMA descriptor and issue wrappers, the work-item decode and tensor-map encoding.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "udal_common.cuh"

namespace {

constexpr int IG_TH = 16, IG_TW = 8;               // output tile: 16 rows x 8 px = 128 GEMM rows
constexpr int IG_ROWS = IG_TH + 2;                 // staged halo tile: 18 rows
constexpr int IG_BOXW = IG_TW + 2;                 // pixels loaded per row
constexpr int KF = 64;
constexpr int kIgSmemLimit = 232448;  // 227 KB opt-in limit of sm_100

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// (a suspend-time hint on try_wait was measured in round 2: no gain for the fused kernels, the tower kernel 10 % slower)
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
__device__ __forceinline__ bool ig_elect_one() {  // true in exactly one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
// 16-bit operand format of the tensor-core heads: bf16 (range of fp32, 8-bit significand) or fp16 (11-bit significand: 8x
// smaller rounding error; what the reference's own GPU export computes in, mixed_float16, infer_lib.py:429-431)
template <bool FP16>
__device__ __forceinline__ uint32_t ig_pack16(float lo, float hi) {
  if constexpr (FP16) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  } else {
    return ig_pack(lo, hi);
  }
}
template <bool FP16>
__device__ __forceinline__ float2 ig_unpack16(uint32_t v) {
  if constexpr (FP16) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  } else {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  }
}
// tcgen05 instruction descriptor, kind::f16: fp32 accumulate, A / B both bf16 (format 1) or fp16 (format 0), K-major, M = 128
template <bool FP16>
__host__ __device__ constexpr uint32_t ig_idesc(int n) {
  return (1u << 4) | (FP16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ float ig_swish_h(float h) {  // h = x / 2
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);  // x * sigmoid(x) = h * tanh(h) + h
}

// two independent fp32 FMAs in one issue slot (FFMA2): acc = a * b + acc per half; bit-identical to fmaf per element
__device__ __forceinline__ void ig_ffma2(float2& acc, const float2 a, const float2 b) {
  unsigned long long c = *reinterpret_cast<unsigned long long*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(c)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  acc = *reinterpret_cast<float2*>(&c);
}

__device__ __forceinline__ float2 ig_fma2(const float2 a, const float2 b, const float2 c) {  // a * b + c per half
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ig_add2(const float2 a, const float2 b) {  // FADD2
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ig_sub2(const float2 a, const float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ig_mul2(const float2 a, const float2 b) {  // FMUL2
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float ig_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

struct IgItem {
  int l, nb, ty0, tx0;
};
// n / d and n % d through magic = ceil(2^32 / d) (0 stands for d = 1); exact for n * d < 2^32 after one
// correction step
__device__ __forceinline__ void ig_divmod(int n, int d, uint32_t magic, int& q, int& r) {
  q = magic ? (int)__umulhi((uint32_t)n, magic) : n;
  r = n - q * d;
  if (r < 0) {
    --q;
    r += d;
  }
}
template <class P>
__device__ __forceinline__ IgItem ig_item(const P& p, int item) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < UDAL_MAX_LEVELS; ++i)
    if (i < p.num_levels && item >= p.item_off[i]) l = i;
  const int r = item - p.item_off[l];
  int nb, tile, ty, tx;
  ig_divmod(r, p.tiles[l], p.tiles_magic[l], nb, tile);
  ig_divmod(tile, p.tiles_x[l], p.tiles_x_magic[l], ty, tx);
  IgItem it;
  it.l = l;
  it.nb = nb;
  it.ty0 = ty * IG_TH;
  it.tx0 = tx * IG_TW;
  return it;
}

// Dynamic work distribution of the persistent kernels: the producer warp claims item indices from a global counter
// (a CTA that starts late - its SM was still busy with another stream's kernel - simply claims fewer) and publishes
// them in a shared-memory ring; every other role reads item k from slot k % IG_QRING once the barrier that orders it
// behind the producer has completed.  -1 ends the stream (written to two slots: both epilogue groups see it).
constexpr int IG_QRING = 16;  // > deepest look-ahead between the producer and the epilogue (stages + accumulators)
__device__ __forceinline__ int ig_claim(int* counter, int items, int lane) {  // whole warp, converged
  int v = 0;
  if (lane == 0) v = atomicAdd(counter, 1);
  v = __shfl_sync(0xffffffffu, v, 0);
  return v < items ? v : -1;
}
__device__ __forceinline__ int ig_queue_read(const volatile int* q, int k) { return q[k & (IG_QRING - 1)]; }
// The producer also publishes the DECODED item next to its index (a second ring of int4): the level search, the two magic
// divisions and the constant-bank look-ups of ig_item are ~60 dependent instructions - 10 % of a builder warp's time per
// tile when every role repeats them (round-2 ncu source view).  Same ordering as the index ring: written before the arrival
// on the stage's full barrier, read behind it.
__device__ __forceinline__ void ig_queue_put(volatile int* q, int4* qw, int k, int item, const IgItem& w) {
  qw[k & (IG_QRING - 1)] = make_int4(w.l, w.nb, w.ty0, w.tx0);
  q[k & (IG_QRING - 1)] = item;
}
__device__ __forceinline__ IgItem ig_queue_item(const int4* qw, int k) {
  int4 v;
  asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"((uint32_t)__cvta_generic_to_shared(qw + (k & (IG_QRING - 1)))));
  IgItem it;
  it.l = v.x;
  it.nb = v.y;
  it.ty0 = v.z;
  it.tx0 = v.w;
  return it;
}

__device__ __forceinline__ void ig_group_sync(int g) {  // the 128 threads of epilogue group g
  asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
}
__device__ __forceinline__ void ig_group_sync256(int g) {  // 256-thread epilogue groups
  asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
}
__device__ __forceinline__ void ig_tma_store(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// ---- depthwise 3x3 on the CUDA cores in packed fp16 (heads_dw.cu) ------------------------------------------------------
// 128 builder threads turn one staged halo tile - linear [18][10] pixels x 64 channels fp16 (128 B per pixel) - into the A
// operand of the pointwise GEMM: [128 px][64 ch] fp16, K-major, 128B swizzle.  thread = 8-channel group (one 16-byte chunk)
// x column pair x row quarter: a 4-row x 2-column patch of outputs, fed by a sliding window over 6 x 4 input pixels (24
// LDS.128 for 8 outputs - the 9-fold tap reuse happens in registers), 288 HFMA2 (two fp16 FMAs per lane and issue slot: twice
// the fp32 rate).  fp16 accumulation of the 9 taps adds ~sqrt(9) * 2^-12 relative to a result that is rounded to fp16 (2^-12)
// for the tensor core anyway.
constexpr int kDwBuilderThreads = 128;
struct DwWeights {
  __half2 w[9][4];  // this thread's 8 channels of the 9 taps
};
__device__ __forceinline__ void dw_load_weights(const float* __restrict__ dw /* [9][64] fp32 */, int cg, DwWeights& W) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(dw + tp * KF + cg * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dw + tp * KF + cg * 8 + 4));
    W.w[tp][0] = __floats2half2_rn(a.x, a.y);
    W.w[tp][1] = __floats2half2_rn(a.z, a.w);
    W.w[tp][2] = __floats2half2_rn(b.x, b.y);
    W.w[tp][3] = __floats2half2_rn(b.z, b.w);
  }
}
// rows_valid: tile rows that lie inside the image (rows past it are never stored: their A rows stay as they are)
__device__ __forceinline__ void dw_build_tile(const uint8_t* __restrict__ sIn, uint8_t* __restrict__ sA, const DwWeights& W, int btid,
                                              int rows_valid) {
  const int cg = btid & 7, x0 = 2 * ((btid >> 3) & 3), y0 = 4 * (btid >> 5);
  if (y0 >= rows_valid) return;  // warp-uniform: a warp is one row quarter
  __half2 acc[4][2][4];
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int oc = 0; oc < 2; ++oc)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[y][oc][q] = __floats2half2_rn(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    uint4 in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      in[c] = *reinterpret_cast<const uint4*>(sIn + (size_t)((y0 + r) * IG_BOXW + x0 + c) * 128 + cg * 16);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = r - dy;
      if (y >= 0 && y < 4) {
#pragma unroll
        for (int oc = 0; oc < 2; ++oc)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const __half2* v = reinterpret_cast<const __half2*>(&in[oc + dx]);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[y][oc][q] = __hfma2(v[q], W.w[dy * 3 + dx][q], acc[y][oc][q]);
          }
      }
    }
    if (r >= 2) {
      const int y = r - 2;
#pragma unroll
      for (int oc = 0; oc < 2; ++oc) {
        const int m = (y0 + y) * IG_TW + x0 + oc;
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t*>(&acc[y][oc][0]);
        o.y = *reinterpret_cast<const uint32_t*>(&acc[y][oc][1]);
        o.z = *reinterpret_cast<const uint32_t*>(&acc[y][oc][2]);
        o.w = *reinterpret_cast<const uint32_t*>(&acc[y][oc][3]);
        *reinterpret_cast<uint4*>(sA + (size_t)m * 128 + (((uint32_t)cg ^ (uint32_t)(m & 7)) << 4)) = o;
      }
    }
  }
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

// 4-D tensor map with explicit byte strides of dimensions 1..3 and a {b0, b1, b2, 1} box
int encode_strided(EncodeTiledFn encode, CUtensorMap* map, CUtensorMapDataType dt, const void* base, const cuuint64_t (&gdim)[4],
                   const cuuint64_t (&gstr)[3], int b0, int b1, int b2, bool swizzle128) {
  const cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = encode(map, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    udal_set_error("cuTensorMapEncodeTiled failed (%d) for dims {%llu,%llu,%llu,%llu} box {%d,%d,%d}", (int)r,
                   (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                   (unsigned long long)gdim[3], b0, b1, b2);
    return UDAL_ERR_INVALID;
  }
  return UDAL_OK;
}

// [NB,H,W,ch] tensor map with a {c_box, x_box, y_box, 1} box
int encode_nhwc(EncodeTiledFn encode, CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, int NB, int H,
                int W, int ch, int c_box, int x_box, int y_box, bool swizzle128) {
  const cuuint64_t gdim[4] = {(cuuint64_t)ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  const cuuint64_t gstr[3] = {(cuuint64_t)ch * esize, (cuuint64_t)W * ch * esize, (cuuint64_t)H * W * ch * esize};
  const cuuint32_t box[4] = {(cuuint32_t)c_box, (cuuint32_t)x_box, (cuuint32_t)y_box, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = encode(map, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    udal_set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box {%d,%d,%d}", (int)r, NB, H, W, ch, c_box, x_box,
                   y_box);
    return UDAL_ERR_INVALID;
  }
  return UDAL_OK;
}

}  // namespace
