// K4, global variant: tf.raw_ops.NonMaxSuppressionV5 (hard and gaussian soft NMS) over all N anchors of an image as ONE
// CTA-wide cooperative kernel per image - score pre-filter, selection and exactness check in a single launch.
//
// Replaces the TF kernel called at reference src/postprocess.py:392-400 from postprocess_global (:472-621).
//
// The TF kernel is a lazy max-heap loop (oracle/nms_v5.c): pop the best candidate, decay it by the boxes selected since
// it was last looked at (newest first), select it if its score did not change, re-insert it otherwise.  The selections
// are inherently sequential, the work between two selections is not.  With the selected set S fixed ("epoch" e = |S|),
// let s'(c) be the score candidate c would have after its pending decay chain S[e-1] .. S[begin(c)].  Then
//   * the next selection is c* = argmax_c s'(c)  (ties: lower box index - the heap's rule), and
//   * exactly the candidates whose CURRENT key (score, index) precedes the key (s'(c*), index(c*)) are popped before it:
//     they take their decayed score (or die at the threshold), their begin becomes e; everybody else is untouched.
// (Proof sketch: keys only decrease; the heap pops in key order; a popped candidate re-enters with s' and is selected
// on its next pop - empty pending range - unless something with a larger key is still ahead; the first key that
// survives its own pop unchanged is the maximum of s'.)  One epoch is therefore: every thread evaluates the chains of
// its candidates that can matter, a block-wide arg-max, a commit.  Candidates evaluated in the previous epoch have a
// one-box chain ("hot"); a candidate whose current score is already behind the hot maximum cannot be popped and is
// skipped, so the total IoU work stays bounded by (candidates x selections) and is in practice a small multiple of the
// candidate count.  The fp32 product of the decay weights is taken in exactly the oracle's order: results are bit-exact.
//
// Candidates: the K best scores of the image (in-CTA radix select over the ordered-uint keys, no sort needed - the
// epoch loop never looks at an order), resident in shared memory.  The truncation is provably exact when every
// selection scored above the best excluded candidate; otherwise the segment is flagged and redone over all N
// candidates by the same loop with its state in global memory (nms_epoch_full_kernel, early exit when not flagged).
//
// Arithmetic = oracle/nms_v5.c: fp32 IoU without "+1", weight = fp32(exp(fp64((scale*u)*u))).
#include <math_constants.h>

#include "fast_math64.cuh"
#include "udal_common.cuh"

namespace {

constexpr int kCtaThreads = 1024;
constexpr int kCtaWarps = kCtaThreads / 32;
constexpr int kHistBins = 2048;

struct EpochParams {
  const float* boxes;    // [S,n,4]
  const float* scores;   // [S,n]
  int segments, n, max_out, cap;   // cap = candidates the shared-memory variant keeps
  float iou_thr, score_thr, scale;
  int soft, variant_old;
  int32_t* sel_row;      // [S,max_out] zero padded
  float* sel_scores;     // [S,max_out]
  int32_t* valid;        // [S]
  int32_t* flag;         // [S] 1 = the truncated run is not provably exact (written by the CTA kernel, read by the full one)
  // full variant: state in global memory, [S,n] each
  float* g_cur;
  float* g_sp;
  uint32_t* g_meta;
};

__device__ __forceinline__ uint32_t ordered_key(float f) {  // monotone float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// heap key of (score, index): larger = popped earlier (score descending, index ascending)
__device__ __forceinline__ unsigned long long heap_key(float s, int idx) {
  return ((unsigned long long)ordered_key(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

__device__ __forceinline__ float iou_v5(const float4 a, const float4 b) {
  const float ay0 = fminf(a.x, a.z), ax0 = fminf(a.y, a.w);
  const float ay1 = fmaxf(a.x, a.z), ax1 = fmaxf(a.y, a.w);
  const float by0 = fminf(b.x, b.z), bx0 = fminf(b.y, b.w);
  const float by1 = fmaxf(b.x, b.z), bx1 = fmaxf(b.y, b.w);
  const float area_a = __fmul_rn(__fsub_rn(ay1, ay0), __fsub_rn(ax1, ax0));
  const float area_b = __fmul_rn(__fsub_rn(by1, by0), __fsub_rn(bx1, bx0));
  if (area_a <= 0.f || area_b <= 0.f) return 0.f;
  const float iy0 = fmaxf(ay0, by0), ix0 = fmaxf(ax0, bx0);
  const float iy1 = fminf(ay1, by1), ix1 = fminf(ax1, bx1);
  const float ih = fmaxf(__fsub_rn(iy1, iy0), 0.f), iw = fmaxf(__fsub_rn(ix1, ix0), 0.f);
  const float inter = __fmul_rn(ih, iw);
  if (inter == 0.f) return 0.f;  // (= 0 / union: skips the division for the disjoint pairs, the common case)
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// the oracle's decay chain of one candidate over the selections sel[from] .. sel[to] (newest first); returns the decayed
// score, dead = the candidate leaves the heap when popped (hard suppression or score at / below the threshold)
__device__ __forceinline__ float decay_chain(float s, const float4 box, int from, int to, const float4* sel_box,
                                             const EpochParams& p, const double* tbl, bool& dead) {
  dead = false;
  for (int q = from; q >= to; --q) {
    const float u = iou_v5(box, sel_box[q]);
    if (u == 0.f && p.soft && !p.variant_old) continue;  // weight exactly 1, no hard rule in this mode
    float w = (u == 0.f || p.scale == 0.f) ? 1.f : (float)exp_fast((double)__fmul_rn(__fmul_rn(p.scale, u), u), tbl);
    bool hard;
    if (p.variant_old) {
      if (!(u <= p.iou_thr)) w = 0.f;
      hard = u >= p.iou_thr;
    } else {
      if (!(p.soft || u <= p.iou_thr)) w = 0.f;
      hard = !p.soft && u > p.iou_thr;
    }
    s = __fmul_rn(s, w);
    if (hard) {
      dead = true;
      return s;
    }
    if (s <= p.score_thr) break;
  }
  if (!(s > p.score_thr)) dead = true;
  return s;
}

// block-wide maximum of a 64-bit key: warp reduction, one shared-memory atomic per warp
__device__ __forceinline__ void block_max_push(unsigned long long key, unsigned long long* slot, int lane) {
  const uint32_t hi = (uint32_t)(key >> 32);
  const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mhi ? (uint32_t)key : 0u;
  const uint32_t mlo = __reduce_max_sync(0xffffffffu, lo);
  if (lane == 0) {
    const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
    if (m) atomicMax(slot, m);
  }
}

// meta word of a candidate: begin (low 16 bits) | epoch stamp of its last evaluation + 1 (high 16 bits)
__device__ __forceinline__ int meta_begin(uint32_t m) { return (int)(m & 0xffffu); }
__device__ __forceinline__ int meta_stamp(uint32_t m) { return (int)(m >> 16); }

// The epoch loop.  Candidate j: box bx[j], index ix ? ix[j] : j, current score cur[j] (-inf = not in the heap), meta[j],
// sp[j] = decayed score of this epoch's evaluation.  Returns through shared variables: nsel, last selected key.
// slots: 4 x u64 in shared memory (zeroed); sel_box: [max_out] in shared memory.
template <bool GLOBAL>
__device__ __forceinline__ void epoch_loop(const EpochParams& p, int cnt, const float4* __restrict__ bx, const int32_t* ix,
                                           float* cur, float* sp, uint32_t* meta, float4* sel_box, unsigned long long* slots,
                                           const double* tbl, int32_t* out_row, float* out_score, int& nsel_out,
                                           float& last_out, bool& emptied_out) {
  const int tid = threadIdx.x, lane = tid & 31;
  int nsel = 0;
  float last = CUDART_INF_F;
  bool emptied = false;
  for (int e = 0; e < p.max_out; ++e) {
    unsigned long long* slot1 = slots + 2 * (e & 1);
    unsigned long long* slot2 = slot1 + 1;
    // ---- round 1: candidates whose pending chain is empty or the newest selection only ----
    unsigned long long best = 0;
    for (int j = tid; j < cnt; j += kCtaThreads) {
      const float s = cur[j];
      if (s == -CUDART_INF_F) continue;
      const uint32_t m = meta[j];
      const int b = meta_begin(m);
      if (b < e - 1) continue;  // cold
      float v = s;
      bool dead = false;
      if (b == e - 1) v = decay_chain(s, bx[j], e - 1, e - 1, sel_box, p, tbl, dead);
      sp[j] = dead ? -CUDART_INF_F : v;
      meta[j] = (uint32_t)b | ((uint32_t)(e + 1) << 16);
      if (!dead) {
        const unsigned long long k = heap_key(v, ix ? ix[j] : j);
        best = k > best ? k : best;
      }
    }
    block_max_push(best, slot1, lane);
    __syncthreads();
    const unsigned long long m1 = *slot1;
    // ---- round 2: cold candidates that would be popped before the round-1 maximum ----
    best = 0;
    int worked = 0;
    for (int j = tid; j < cnt; j += kCtaThreads) {
      const float s = cur[j];
      if (s == -CUDART_INF_F) continue;
      const uint32_t m = meta[j];
      if (meta_stamp(m) == e + 1) continue;  // evaluated in round 1
      if (heap_key(s, ix ? ix[j] : j) < m1) continue;
      bool dead;
      const float v = decay_chain(s, bx[j], e - 1, meta_begin(m), sel_box, p, tbl, dead);
      sp[j] = dead ? -CUDART_INF_F : v;
      meta[j] = (m & 0xffffu) | ((uint32_t)(e + 1) << 16);
      worked = 1;
      if (!dead) {
        const unsigned long long k = heap_key(v, ix ? ix[j] : j);
        best = k > best ? k : best;
      }
    }
    unsigned long long mk = m1;
    if (__syncthreads_or(worked)) {
      block_max_push(best, slot2, lane);
      __syncthreads();
      const unsigned long long m2 = *slot2;
      mk = m2 > m1 ? m2 : m1;
    }
    // ---- commit: everything evaluated whose current key precedes the winner's key was popped before it ----
    const int win_idx = (int)(0xffffffffu - (uint32_t)mk);
    for (int j = tid; j < cnt; j += kCtaThreads) {
      const float s = cur[j];
      if (s == -CUDART_INF_F) continue;
      const uint32_t m = meta[j];
      if (meta_stamp(m) != e + 1) continue;
      const int idx = ix ? ix[j] : j;
      if (mk != 0 && idx == win_idx) {  // the selection (its own evaluation left the score unchanged or it re-entered with sp)
        const float v = sp[j];
        sel_box[e] = bx[j];
        out_row[e] = idx;
        out_score[e] = v;
        cur[j] = -CUDART_INF_F;
      } else if (mk == 0 || heap_key(s, idx) > mk) {
        cur[j] = sp[j];                        // (-inf when it died)
        meta[j] = (uint32_t)e | (m & 0xffff0000u);
      }
    }
    if (tid == 0) {
      // next epoch's slots (nobody touches them before the barrier below)
      slots[2 * ((e + 1) & 1)] = 0;
      slots[2 * ((e + 1) & 1) + 1] = 0;
    }
    if (mk == 0) {
      emptied = true;
      break;
    }
    last = key_to_float((uint32_t)(mk >> 32));
    nsel = e + 1;
    __syncthreads();
  }
  nsel_out = nsel;
  last_out = last;
  emptied_out = emptied;
}

// ---------------------------------------------------------------------------------------------------------------------
// one CTA per image: select the `cap` best scores into shared memory, run the epoch loop, flag an unprovable truncation
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) nms_epoch_cta_kernel(const EpochParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n, cap = p.cap;
  float4* s_box = reinterpret_cast<float4*>(smem);                       // [cap]
  float4* sel_box = s_box + cap;                                         // [max_out]
  float* s_cur = reinterpret_cast<float*>(sel_box + p.max_out);          // [cap]
  float* s_sp = s_cur + cap;                                             // [cap]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_sp + cap);               // [cap]
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_idx + cap);           // [cap]
  uint32_t* s_hist = s_meta + cap;                                       // [kHistBins]
  __shared__ double s_tbl[64];
  __shared__ unsigned long long s_slots[4];
  __shared__ uint32_t s_prefix, s_need, s_count, s_above;
  const float* scores = p.scores + (size_t)s * n;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)s * n;
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  const float thr = p.score_thr;
  if (tid < 64) s_tbl[tid] = kExp2Table[tid];
  if (tid < 4) s_slots[tid] = 0;
  if (tid == 0) {
    s_count = 0;
    s_above = 0;
  }
  __syncthreads();

  // ---- how many candidates are there at all ----
  {
    int c = 0;
    for (int j = tid; j < n; j += kCtaThreads) c += scores[j] > thr ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&s_above, (uint32_t)c);
  }
  __syncthreads();
  const int above = (int)s_above;
  // ---- radix select: key of the cap-th best candidate (3 passes: 11 + 11 + 10 bits) ----
  uint32_t cut = 0;        // candidates are the scores with ordered key > cut (and > thr)
  bool truncated = false;
  if (above > cap) {
    truncated = true;
    uint32_t prefix = 0, need = (uint32_t)cap + 1;  // the (cap + 1)-th best key is the best excluded one
    for (int pass = 0; pass < 3; ++pass) {
      const int sh = pass == 0 ? 21 : (pass == 1 ? 10 : 0), width = pass == 2 ? 10 : 11, bins = 1 << width;
      for (int i = tid; i < bins; i += kCtaThreads) s_hist[i] = 0;
      __syncthreads();
      const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (sh + width));
      for (int j0 = 0; j0 < n; j0 += kCtaThreads) {
        const int j = j0 + tid;
        bool ok = false;
        uint32_t bin = 0;
        if (j < n) {
          const float sc = scores[j];
          const uint32_t k = ordered_key(sc);
          ok = sc > thr && (k & himask) == prefix;
          bin = (k >> sh) & (uint32_t)(bins - 1);
        }
        // warp-aggregated histogram update (the scores of a random-init head are nearly flat: most lanes hit one bin)
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const uint32_t peers = __match_any_sync(act, bin);
          if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
        }
      }
      __syncthreads();
      // warp 0: the bin that holds the need-th largest key among the matching ones
      if (warp == 0) {
        const int per = bins / 32;
        uint32_t mine = 0;
        for (int i = 0; i < per; ++i) mine += s_hist[bins - 1 - (lane * per + i)];
        uint32_t incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += o;
        }
        const uint32_t excl = incl - mine;
        if (need > excl && need <= incl) {
          uint32_t acc = excl;
          for (int i = 0; i < per; ++i) {
            const int b = bins - 1 - (lane * per + i);
            const uint32_t h = s_hist[b];
            if (need <= acc + h) {
              s_prefix = prefix | ((uint32_t)b << sh);
              s_need = need - acc;
              break;
            }
            acc += h;
          }
        }
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      __syncthreads();
    }
    cut = prefix;  // the best excluded key: candidates = keys > cut  (at most cap of them)
  }
  // ---- compaction into shared memory (any order: the epoch loop ties by box index) ----
  for (int j0 = 0; j0 < n; j0 += kCtaThreads) {
    const int j = j0 + tid;
    bool ok = false;
    float sc = 0.f;
    if (j < n) {
      sc = scores[j];
      ok = sc > thr && (!truncated || ordered_key(sc) > cut);
    }
    const uint32_t act = __ballot_sync(0xffffffffu, ok);
    if (act) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&s_count, (uint32_t)__popc(act));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (ok) {
        const int slot = (int)(base + (uint32_t)__popc(act & ((1u << lane) - 1u)));
        s_box[slot] = boxes[j];
        s_cur[slot] = sc;
        s_idx[slot] = j;
        s_meta[slot] = 0;
      }
    }
  }
  __syncthreads();
  const int cnt = (int)s_count;

  int nsel;
  float last;
  bool emptied;
  epoch_loop<false>(p, cnt, s_box, s_idx, s_cur, s_sp, s_meta, sel_box, s_slots, s_tbl, out_row, out_score, nsel, last, emptied);
  __syncthreads();
  for (int i = nsel + tid; i < p.max_out; i += kCtaThreads) {
    out_row[i] = 0;
    out_score[i] = 0.f;
  }
  if (tid == 0) {
    p.valid[s] = nsel;
    if (p.flag) {
      const float nx = key_to_float(cut);
      // exact iff nothing excluded could have been popped: every pop scored above the best excluded candidate
      p.flag[s] = (truncated && (emptied || nsel == 0 || !(last > nx))) ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// exact redo of a flagged image over all n candidates: the same loop, state in global memory (L2 resident)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) nms_epoch_full_kernel(const EpochParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int s = blockIdx.x, tid = threadIdx.x;
  if (p.flag && !p.flag[s]) return;
  const int n = p.n;
  float4* sel_box = reinterpret_cast<float4*>(smem);  // [max_out]
  __shared__ double s_tbl[64];
  __shared__ unsigned long long s_slots[4];
  const float* scores = p.scores + (size_t)s * n;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)s * n;
  float* cur = p.g_cur + (size_t)s * n;
  float* sp = p.g_sp + (size_t)s * n;
  uint32_t* meta = p.g_meta + (size_t)s * n;
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  if (tid < 64) s_tbl[tid] = kExp2Table[tid];
  if (tid < 4) s_slots[tid] = 0;
  for (int j = tid; j < n; j += kCtaThreads) {
    const float sc = scores[j];
    cur[j] = sc > p.score_thr ? sc : -CUDART_INF_F;
    meta[j] = 0;
  }
  __syncthreads();
  int nsel;
  float last;
  bool emptied;
  epoch_loop<true>(p, n, boxes, nullptr, cur, sp, meta, sel_box, s_slots, s_tbl, out_row, out_score, nsel, last, emptied);
  __syncthreads();
  for (int i = nsel + tid; i < p.max_out; i += kCtaThreads) {
    out_row[i] = 0;
    out_score[i] = 0.f;
  }
  if (tid == 0) p.valid[s] = nsel;
}

}  // namespace

int udal_nms_cta = 1;  // 0: global NMS through the top-k pre-filter + one-warp-per-image kernels of nms.cu (comparison path)

// Global NMS-V5 over [S,n] boxes / scores (unsorted): one cooperative CTA per image + the exact redo of flagged images.
// Everything is enqueued; no host synchronisation.
int udal_nms_epoch(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n, int32_t* sel_idx,
                   float* sel_scores, int32_t* valid) {
  const udal_config& c = ctx->cfg;
  EpochParams p;
  memset(&p, 0, sizeof(p));
  p.boxes = boxes;
  p.scores = scores;
  p.segments = segments;
  p.n = n;
  p.max_out = c.max_output_size;
  UDAL_REQUIRE(p.max_out < 65535, "max_output_size %d too large", p.max_out);
  int cap = udal_nms_prefilter_k(ctx, n);
  if (cap < 1) cap = 1;
  p.cap = cap;
  p.iou_thr = c.nms_iou_thresh;
  p.score_thr = c.nms_score_thresh;
  p.soft = c.nms_sigma_tf > 0.f;
  p.scale = p.soft ? (-0.5f / c.nms_sigma_tf) : 0.f;
  p.variant_old = c.nms_variant_old;
  p.sel_row = sel_idx;
  p.sel_scores = sel_scores;
  p.valid = valid;
  const bool may_truncate = cap < n;
  if (may_truncate) {
    char* scr;
    const size_t per = (size_t)segments * n;
    UDAL_TRY(udal_scratch_get(ctx, SCR_NMS_B, per * 12 + (size_t)segments * 4, (void**)&scr));
    p.g_cur = (float*)scr;
    p.g_sp = (float*)(scr + per * 4);
    p.g_meta = (uint32_t*)(scr + per * 8);
    p.flag = (int32_t*)(scr + per * 12);
  }
  const size_t smem = (size_t)cap * 32 + (size_t)p.max_out * 16 + kHistBins * 4;
  UDAL_REQUIRE(smem <= 200 * 1024, "nms: %d candidates x max_output_size %d do not fit in shared memory", cap, p.max_out);
  UDAL_CUDA(cudaFuncSetAttribute(nms_epoch_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_epoch_cta_kernel<<<segments, kCtaThreads, smem, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  if (may_truncate) {
    const size_t smem2 = (size_t)p.max_out * 16;
    if (smem2 > 40 * 1024)
      UDAL_CUDA(cudaFuncSetAttribute(nms_epoch_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    nms_epoch_full_kernel<<<segments, kCtaThreads, smem2, ctx->stream>>>(p);
    UDAL_CHECK_LAUNCH(ctx);
  }
  return UDAL_OK;
}
