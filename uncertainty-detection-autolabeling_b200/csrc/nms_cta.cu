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
// survives its own pop unchanged is the maximum of s'.  tests/test_oracle_extra.py holds a NumPy transcription of this
// formulation checked against the heap of nms_v5.c.)  One epoch:
//   round 1  every live candidate tests its box against the newest selection (interval overlap: a few instructions) and
//            records a non-trivial step (IoU != 0) in its 128-bit mask.  A candidate without pending non-trivial steps
//            has s' = s for free; the candidates popped in the previous epoch have a one-step chain.  Block maximum m1
//            of these keys = a lower bound of the winner's key.
//   round 2  a candidate with pending non-trivial steps matters only if its CURRENT key reaches m1: those (typically
//            two per epoch) run their chain over the flagged selections, newest first.  Block maximum m2.
//   commit   the owner of max(m1, m2) appends its candidate to the selection; evaluated candidates whose current key
//            precedes the winner's take their decayed score and begin = e.
// The fp32 product of the decay weights is taken in exactly the oracle's order (steps at IoU 0 multiply by exactly 1):
// results are bit-exact.
//
// Candidates: the K best scores of the image, resident in shared memory (no sort: the epoch loop never looks at an
// order).  The cut is found by an adaptive histogram select over the ordered-uint keys (bins spread over the key range
// of the image, refined while the boundary bin is too full), any cut with at most K keys above it will do.  The
// truncation is provably exact when every selection scored above the cut; otherwise the image is flagged and redone over
// all N candidates by the same loop with its state in global memory (nms_epoch_full_kernel; no work when nothing is
// flagged).
//
// Arithmetic = oracle/nms_v5.c: fp32 IoU without "+1", weight = fp32(exp(fp64((scale*u)*u))).
#include <math_constants.h>

#include "fast_math64.cuh"
#include "udal_common.cuh"

// development counters (udal_nms_debug = 1): cycles of image 0 - select, round 1, round 2, commit; round-2 evaluations
__device__ unsigned long long g_nms_dbg[8];

namespace {

constexpr int kCtaThreads = 1024;
constexpr int kHistBins = 2048;
constexpr int kUnroll = 8;       // independent score loads per thread in the select passes
constexpr int kFullSlots = 32;   // CTAs (= state slots) of the exact redo

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
  int32_t* cursor;       // work cursor of the full kernel (zeroed by the CTA kernel's launch)
  char* g_state;         // full variant: kFullSlots x n x kStateBytes
  int debug;
};

// per-candidate state (structure of arrays over `cap` / n candidates)
constexpr int kStateBytes = 16 + 4 + 4 + 4 + 4 + 4 + 16;
struct State {
  float4* box;      // normalised corners (min y, min x, max y, max x)
  float* area;
  float* cur;       // current heap score, -inf = not in the heap
  float* sp;        // decayed score of this epoch's evaluation (-inf: dies when popped)
  int32_t* idx;     // box index in the image (the heap's tie rule); null = j
  uint32_t* meta;   // begin (bits 0-7) | newest non-trivial selection + 1 (bits 8-15) | epoch stamp of the evaluation + 1 (16-31)
  uint32_t* mask;   // [4] per candidate: selections with a non-trivial step
  __device__ void carve(char* base, size_t count, bool with_idx) {
    box = reinterpret_cast<float4*>(base);
    mask = reinterpret_cast<uint32_t*>(base + count * 16);
    area = reinterpret_cast<float*>(base + count * 32);
    cur = area + count;
    sp = cur + count;
    meta = reinterpret_cast<uint32_t*>(sp + count);
    idx = with_idx ? reinterpret_cast<int32_t*>(meta + count) : nullptr;
  }
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

__device__ __forceinline__ float4 normalise(const float4 a, float& area) {
  const float y0 = fminf(a.x, a.z), x0 = fminf(a.y, a.w), y1 = fmaxf(a.x, a.z), x1 = fmaxf(a.y, a.w);
  area = __fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x1, x0));
  return make_float4(y0, x0, y1, x1);
}
// intersection area of two normalised boxes (0 when either area is not positive: the IoU is 0 then)
__device__ __forceinline__ float intersection(const float4 a, float area_a, const float4 b, float area_b) {
  if (area_a <= 0.f || area_b <= 0.f) return 0.f;
  const float ih = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float iw = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  return __fmul_rn(ih, iw);
}
__device__ __forceinline__ float iou_from(float inter, float area_a, float area_b) {
  if (inter == 0.f) return 0.f;   // (= 0 / union)
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// the decay weight of one step of the oracle's loop at IoU u, and whether the step suppresses the candidate for good
__device__ __forceinline__ float step_weight(float u, const EpochParams& p, const double* tbl, bool& hard) {
  float w = (u == 0.f || p.scale == 0.f) ? 1.f : (float)exp_fast((double)__fmul_rn(__fmul_rn(p.scale, u), u), tbl);
  if (p.variant_old) {
    if (!(u <= p.iou_thr)) w = 0.f;
    hard = u >= p.iou_thr;
  } else {
    if (!(p.soft || u <= p.iou_thr)) w = 0.f;
    hard = !p.soft && u > p.iou_thr;
  }
  return w;
}
// one step of the oracle's decay loop: the candidate (score s) against a selected box at IoU u.  Returns true when the
// loop ends here (hard suppression -> dead, or the score fell to the threshold)
__device__ __forceinline__ bool decay_step(float& s, float u, const EpochParams& p, const double* tbl, bool& dead) {
  bool hard;
  const float w = step_weight(u, p, tbl, hard);
  s = __fmul_rn(s, w);
  if (hard) {
    dead = true;
    return true;
  }
  return s <= p.score_thr;
}

// bits [lo, hi) of a 128-bit mask, word w
__device__ __forceinline__ uint32_t range_word(int w, int lo, int hi) {
  const int a = max(lo - 32 * w, 0), b = min(hi - 32 * w, 32);
  if (a >= b) return 0u;
  const uint32_t upto_b = b == 32 ? 0xffffffffu : ((1u << b) - 1u);
  return upto_b & ~((1u << a) - 1u);   // a < 32 here
}

// The oracle's decay chain of ONE candidate over its pending selections sel[e-1] .. sel[begin], newest first, computed by
// the whole warp: only the selections flagged in the mask can change the score or trigger a rule; lane i takes the i-th
// newest flagged selection (IoU, fp64 exp: the long-latency part, in parallel), then the fp32 product is taken in the
// oracle's order.  All arguments are warp-uniform; every lane returns the same result.
__device__ __forceinline__ float decay_chain_warp(float s, const float4 box, float area, uint32_t m0, uint32_t m1, uint32_t m2,
                                                  uint32_t m3, int begin, int e, const float4* sel_box, const float* sel_area,
                                                  const EpochParams& p, const double* tbl, bool& dead) {
  const int lane = threadIdx.x & 31;
  dead = false;
  uint32_t words[4] = {m0 & range_word(0, begin, e), m1 & range_word(1, begin, e), m2 & range_word(2, begin, e),
                       m3 & range_word(3, begin, e)};
  int left = __popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]);
  while (left > 0) {
    // lane i: the i-th newest flagged selection still pending
    int q = -1, skip = lane;
#pragma unroll
    for (int w = 3; w >= 0; --w) {
      const int c = __popc(words[w]);
      if (q < 0) {
        if (skip < c) q = 32 * w + (int)__fns(words[w], 31, -(skip + 1));
        else skip -= c;
      }
    }
    float wgt = 1.f;
    bool hard = false;
    if (q >= 0) {
      const float a2 = sel_area[q];
      wgt = step_weight(iou_from(intersection(box, area, sel_box[q], a2), area, a2), p, tbl, hard);
    }
    const int take = min(left, 32);
    for (int i = 0; i < take; ++i) {
      const float wi = __shfl_sync(0xffffffffu, wgt, i);
      const int hi = __shfl_sync(0xffffffffu, (int)hard, i);
      s = __fmul_rn(s, wi);
      if (hi) {
        dead = true;
        return s;
      }
      if (s <= p.score_thr) {
        dead = true;
        return s;
      }
    }
    left -= take;
    if (left > 0) {  // drop the 32 newest bits that were just applied
      int drop = 32;
#pragma unroll
      for (int w = 3; w >= 0; --w) {
        const int c = __popc(words[w]);
        if (drop >= c) {
          drop -= c;
          words[w] = 0;
        } else if (drop > 0) {
          const int pos = (int)__fns(words[w], 31, -drop);   // the drop-th newest bit of this word
          words[w] &= (1u << pos) - 1u;
          drop = 0;
        }
      }
    }
  }
  if (!(s > p.score_thr)) dead = true;
  return s;
}

// maximum of a 64-bit key over the warp (two 32-bit redux steps)
__device__ __forceinline__ unsigned long long warp_max(unsigned long long key) {
  const uint32_t hi = (uint32_t)(key >> 32);
  const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mhi ? (uint32_t)key : 0u;
  const uint32_t mlo = __reduce_max_sync(0xffffffffu, lo);
  return ((unsigned long long)mhi << 32) | mlo;
}
// block-wide maximum: one slot per warp, a barrier, every warp reduces the 32 slots again (no atomics: a 64-bit shared
// memory atomicMax is a CAS loop that 32 warps fight over)
__device__ __forceinline__ unsigned long long block_max(unsigned long long key, unsigned long long* wslots) {
  const unsigned long long m = warp_max(key);
  if ((threadIdx.x & 31) == 0) wslots[threadIdx.x >> 5] = m;
  __syncthreads();
  return warp_max(wslots[threadIdx.x & 31]);
}

__device__ __forceinline__ int meta_begin(uint32_t m) { return (int)(m & 0xffu); }
__device__ __forceinline__ int meta_nt(uint32_t m) { return (int)((m >> 8) & 0xffu); }
__device__ __forceinline__ int meta_stamp(uint32_t m) { return (int)(m >> 16); }

// The epoch loop over `cnt` candidates.  wslots: 64 x u64 in shared memory; sel_box / sel_area: [max_out <= 128].
__device__ __forceinline__ void epoch_loop(const EpochParams& p, int cnt, const State& st, float4* sel_box, float* sel_area,
                                           unsigned long long* wslots, const double* tbl, int32_t* out_row, float* out_score,
                                           int& nsel_out, float& last_out, bool& emptied_out) {
  const int tid = threadIdx.x, lane = tid & 31;
  int nsel = 0;
  float last = CUDART_INF_F;
  bool emptied = false;
  // does a step at IoU 0 leave the candidate untouched?  (always, unless a non-positive IoU threshold makes it a hard rule)
  const bool zero_trivial = p.variant_old ? (p.iou_thr > 0.f) : (p.soft || p.iou_thr >= 0.f);
  for (int e = 0; e < p.max_out; ++e) {
    float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
    float narea = 0.f;
    if (e > 0) {
      nb = sel_box[e - 1];
      narea = sel_area[e - 1];
    }
    const long long t0 = p.debug ? clock64() : 0;
    // ---- round 1 ----
    unsigned long long best = 0;
    int bestj = -1;
    bool has_cold = false, mine = false;
    for (int j = tid; j < cnt; j += kCtaThreads) {
      float s = st.cur[j];
      if (s == -CUDART_INF_F) continue;
      uint32_t m = st.meta[j];
      float inter = 0.f, area = 0.f;
      if (e > 0) {
        const float4 box = st.box[j];
        area = st.area[j];
        inter = intersection(box, area, nb, narea);
        if (inter != 0.f || !zero_trivial) {
          st.mask[4 * (size_t)j + ((e - 1) >> 5)] |= 1u << ((e - 1) & 31);
          m = (m & 0xffff00ffu) | ((uint32_t)e << 8);
          st.meta[j] = m;
        }
      }
      const int b = meta_begin(m);
      if (meta_nt(m) > b) {
        if (b != e - 1) {        // pending non-trivial steps from earlier epochs: round 2, if it can matter at all
          has_cold = true;
          continue;
        }
        bool dead = false;       // popped in the previous epoch: one step, against the newest selection
        decay_step(s, iou_from(inter, area, narea), p, tbl, dead);
        if (!(s > p.score_thr)) dead = true;
        st.sp[j] = dead ? -CUDART_INF_F : s;
        st.meta[j] = (m & 0xffffu) | ((uint32_t)(e + 1) << 16);
        mine = true;
        if (dead) continue;
      }
      const unsigned long long k = heap_key(s, st.idx ? st.idx[j] : j);
      if (k > best) {
        best = k;
        bestj = j;
      }
    }
    if (p.debug && blockIdx.x == 0) {   // round-1 loop alone: thread 0 and the slowest warp
      const long long tl = clock64() - t0;
      if (tid == 0) g_nms_dbg[6] += (unsigned long long)tl;
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(wslots + 63), (unsigned int)tl);
    }
    const unsigned long long m1 = block_max(best, wslots);
    const long long t1 = p.debug ? clock64() : 0;
    if (p.debug && blockIdx.x == 0 && tid == 0) {
      g_nms_dbg[7] += (unsigned long long)*reinterpret_cast<unsigned int*>(wslots + 63);
      *reinterpret_cast<unsigned int*>(wslots + 63) = 0;
    }
    // ---- round 2: candidates with a pending chain that would be popped before the round-1 maximum; each is evaluated by
    //      its whole warp (the owner's lanes are otherwise idle here) ----
    int worked = 0;
    unsigned long long best2 = 0;
    if (__any_sync(0xffffffffu, has_cold)) {
      for (int j0 = tid - lane; j0 < cnt; j0 += kCtaThreads) {
        const int j = j0 + lane;
        bool need = false;
        float s = 0.f;
        uint32_t m = 0;
        if (j < cnt && has_cold) {
          s = st.cur[j];
          if (s != -CUDART_INF_F) {
            m = st.meta[j];
            const int b = meta_begin(m);
            need = meta_nt(m) > b && b != e - 1 && heap_key(s, st.idx ? st.idx[j] : j) >= m1;
          }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, need);
        if (!todo) continue;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float area = 0.f;
        uint4 mw = make_uint4(0u, 0u, 0u, 0u);
        if (need) {
          box = st.box[j];
          area = st.area[j];
          mw = *reinterpret_cast<const uint4*>(st.mask + 4 * (size_t)j);
        }
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          float4 bb;
          bb.x = __shfl_sync(0xffffffffu, box.x, src);
          bb.y = __shfl_sync(0xffffffffu, box.y, src);
          bb.z = __shfl_sync(0xffffffffu, box.z, src);
          bb.w = __shfl_sync(0xffffffffu, box.w, src);
          const float ba = __shfl_sync(0xffffffffu, area, src);
          const float bs = __shfl_sync(0xffffffffu, s, src);
          const int bbeg = meta_begin(__shfl_sync(0xffffffffu, m, src));
          const uint32_t w0 = __shfl_sync(0xffffffffu, mw.x, src), w1 = __shfl_sync(0xffffffffu, mw.y, src);
          const uint32_t w2 = __shfl_sync(0xffffffffu, mw.z, src), w3 = __shfl_sync(0xffffffffu, mw.w, src);
          bool dead;
          const float v = decay_chain_warp(bs, bb, ba, w0, w1, w2, w3, bbeg, e, sel_box, sel_area, p, tbl, dead);
          if (lane == src) {
            st.sp[j] = dead ? -CUDART_INF_F : v;
            st.meta[j] = (m & 0xffffu) | ((uint32_t)(e + 1) << 16);
            worked = 1;
            mine = true;
            if (!dead) {
              const unsigned long long k = heap_key(v, st.idx ? st.idx[j] : j);
              if (k > best2) best2 = k;
              if (k > best) {
                best = k;
                bestj = j;
              }
            }
          }
        }
      }
    }
    unsigned long long mk = m1;
    if (__syncthreads_or(worked)) {
      const unsigned long long m2 = block_max(best2, wslots + 32);
      mk = m2 > m1 ? m2 : m1;
    }
    const long long t2 = p.debug ? clock64() : 0;
    if (p.debug && blockIdx.x == 0 && worked) atomicAdd(&g_nms_dbg[4], 1ull);
    // ---- commit ----
    if (mk != 0 && best == mk) {  // this thread owns the winner (box indices are unique)
      const int j = bestj;
      const uint32_t m = st.meta[j];
      const float v = meta_stamp(m) == e + 1 ? st.sp[j] : st.cur[j];
      sel_box[e] = st.box[j];
      sel_area[e] = st.area[j];
      out_row[e] = st.idx ? st.idx[j] : j;
      out_score[e] = v;
      st.cur[j] = -CUDART_INF_F;
    }
    if (mine) {  // evaluated candidates whose current key precedes the winner's were popped before it
      for (int j = tid; j < cnt; j += kCtaThreads) {
        const float s = st.cur[j];
        if (s == -CUDART_INF_F) continue;
        const uint32_t m = st.meta[j];
        if (meta_stamp(m) != e + 1) continue;
        if (mk == 0 || heap_key(s, st.idx ? st.idx[j] : j) > mk) {
          st.cur[j] = st.sp[j];                     // (-inf when it died)
          st.meta[j] = (m & 0xffffff00u) | (uint32_t)e;
        }
      }
    }
    if (mk == 0) {
      emptied = true;
      break;
    }
    last = key_to_float((uint32_t)(mk >> 32));
    nsel = e + 1;
    __syncthreads();
    if (p.debug && blockIdx.x == 0 && tid == 0) {
      const long long t3 = clock64();
      g_nms_dbg[1] += (unsigned long long)(t1 - t0);
      g_nms_dbg[2] += (unsigned long long)(t2 - t1);
      g_nms_dbg[3] += (unsigned long long)(t3 - t2);
    }
  }
  nsel_out = nsel;
  last_out = last;
  emptied_out = emptied;
}

// Per-warp slots of the block-wide arg-max: the warp's best key and the box that goes with it, so that after ONE barrier every
// thread knows the winner AND its box (the next epoch's "newest selection") without a second barrier.
struct WinSlots {
  unsigned long long key[2][2][32];   // [epoch parity][round][warp]
  float4 box[2][2][32];
  float area[2][2][32];
  uint32_t cold[2][32];               // the warp still has candidates with a pending chain
};

// The same loop with the candidate state in REGISTERS (the shared-memory kernel: at most CPT x 1024 candidates, thread t owns
// candidates t, t + 1024, ..): no address arithmetic and no shared-memory traffic per candidate and epoch, two barriers per epoch
// (the memory-based loop above: ~125 instructions per candidate and epoch and four barriers - tools/time_nms.py).
template <int CPT, int THREADS = kCtaThreads>
__device__ __forceinline__ void epoch_loop_regs(const EpochParams& p, int cnt, const State& st, float4* sel_box, float* sel_area,
                                                WinSlots* ws, const double* tbl, int32_t* out_row, float* out_score,
                                                int& nsel_out, float& last_out, bool& emptied_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4 box[CPT];
  float area[CPT], cur[CPT], sp[CPT];
  int idx[CPT], beg[CPT], nt[CPT];
  uint32_t m0[CPT], m1w[CPT], m2[CPT], m3[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int j = tid + c * THREADS;
    const bool have = j < cnt;
    box[c] = have ? st.box[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    area[c] = have ? st.area[j] : 0.f;
    cur[c] = have ? st.cur[j] : -CUDART_INF_F;
    idx[c] = have ? st.idx[j] : 0x7fffffff;
    sp[c] = 0.f;
    beg[c] = nt[c] = 0;
    m0[c] = m1w[c] = m2[c] = m3[c] = 0u;
  }
  int nsel = 0;
  float last = CUDART_INF_F;
  bool emptied = false;
  const bool zero_trivial = p.variant_old ? (p.iou_thr > 0.f) : (p.soft || p.iou_thr >= 0.f);
  float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);   // the newest selection (normalised box, area): carried in registers
  float narea = 0.f;
  // publishes the warp's best (key, box) in its slot
  auto publish = [&](unsigned long long best, int bestc, int par, int round) {
    const unsigned long long wm = warp_max(best);
    if (best == wm && wm != 0) {   // exactly one lane (box indices are unique)
      float4 bx = box[0];
      float ar = area[0];
#pragma unroll
      for (int c = 1; c < CPT; ++c)
        if (bestc == c) {
          bx = box[c];
          ar = area[c];
        }
      ws->box[par][round][warp] = bx;
      ws->area[par][round][warp] = ar;
    }
    if (lane == 0) ws->key[par][round][warp] = wm;
  };
  for (int e = 0; e < p.max_out; ++e) {
    const int par = e & 1;
    const uint32_t newbit = e > 0 ? 1u << ((e - 1) & 31) : 0u;
    const int neww = (e - 1) >> 5;
    const long long t0 = p.debug ? clock64() : 0;
    // ---- round 1 ----
    unsigned long long best = 0;
    int bestc = -1;
    uint32_t evaluated = 0, cold = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float s = cur[c];
      if (s == -CUDART_INF_F) continue;
      float inter = 0.f;
      if (e > 0) {
        inter = intersection(box[c], area[c], nb, narea);
        if (inter != 0.f || !zero_trivial) {
          m0[c] |= neww == 0 ? newbit : 0u;
          m1w[c] |= neww == 1 ? newbit : 0u;
          m2[c] |= neww == 2 ? newbit : 0u;
          m3[c] |= neww == 3 ? newbit : 0u;
          nt[c] = e;
        }
      }
      if (nt[c] > beg[c]) {
        if (beg[c] != e - 1) {
          cold |= 1u << c;
          continue;
        }
        bool dead = false;
        decay_step(s, iou_from(inter, area[c], narea), p, tbl, dead);
        if (!(s > p.score_thr)) dead = true;
        sp[c] = dead ? -CUDART_INF_F : s;
        evaluated |= 1u << c;
        if (dead) continue;
      }
      const unsigned long long k = heap_key(s, idx[c]);
      if (k > best) {
        best = k;
        bestc = c;
      }
    }
    publish(best, bestc, par, 0);
    {
      const uint32_t anyc = __ballot_sync(0xffffffffu, cold != 0u);
      if (lane == 0) ws->cold[par][warp] = anyc;
    }
    if (p.debug && blockIdx.x == 0 && tid == 0) g_nms_dbg[6] += (unsigned long long)(clock64() - t0);
    __syncthreads();
    constexpr int NW = THREADS / 32;   // (slots of warps the CTA does not have read as empty)
    const unsigned long long k1 = lane < NW ? ws->key[par][0][lane] : 0ull;
    const unsigned long long mm1 = warp_max(k1);
    const bool any_cold = __ballot_sync(0xffffffffu, lane < NW && ws->cold[par][lane] != 0u) != 0u;
    const long long t1 = p.debug ? clock64() : 0;
    // ---- round 2 ----
    unsigned long long mk = mm1;
    int wround = 0;
    uint32_t wwarp = __ffs(__ballot_sync(0xffffffffu, k1 == mm1)) - 1;
    if (any_cold) {
      unsigned long long best2 = 0;
      int best2c = -1;
      if (__any_sync(0xffffffffu, cold != 0u)) {   // (warp-uniform)
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const bool need = ((cold >> c) & 1u) && heap_key(cur[c], idx[c]) >= mm1;
          uint32_t todo = __ballot_sync(0xffffffffu, need);
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            float4 bb;
            bb.x = __shfl_sync(0xffffffffu, box[c].x, src);
            bb.y = __shfl_sync(0xffffffffu, box[c].y, src);
            bb.z = __shfl_sync(0xffffffffu, box[c].z, src);
            bb.w = __shfl_sync(0xffffffffu, box[c].w, src);
            const float ba = __shfl_sync(0xffffffffu, area[c], src);
            const float bs = __shfl_sync(0xffffffffu, cur[c], src);
            const int bbeg = __shfl_sync(0xffffffffu, beg[c], src);
            const uint32_t w0 = __shfl_sync(0xffffffffu, m0[c], src), w1 = __shfl_sync(0xffffffffu, m1w[c], src);
            const uint32_t w2 = __shfl_sync(0xffffffffu, m2[c], src), w3 = __shfl_sync(0xffffffffu, m3[c], src);
            bool dead;
            const float v = decay_chain_warp(bs, bb, ba, w0, w1, w2, w3, bbeg, e, sel_box, sel_area, p, tbl, dead);
            if (lane == src) {
              sp[c] = dead ? -CUDART_INF_F : v;
              evaluated |= 1u << c;
              if (p.debug && blockIdx.x == 0) atomicAdd(&g_nms_dbg[4], 1ull);
              if (!dead) {
                const unsigned long long k = heap_key(v, idx[c]);
                if (k > best2) {
                  best2 = k;
                  best2c = c;
                }
                if (k > best) {
                  best = k;
                  bestc = c;
                }
              }
            }
          }
        }
      }
      publish(best2, best2c, par, 1);
      __syncthreads();
      const unsigned long long k2 = lane < NW ? ws->key[par][1][lane] : 0ull;
      const unsigned long long mm2 = warp_max(k2);
      if (mm2 > mm1) {
        mk = mm2;
        wround = 1;
        wwarp = __ffs(__ballot_sync(0xffffffffu, k2 == mm2)) - 1;
      }
    }
    const long long t2 = p.debug ? clock64() : 0;
    if (mk == 0) {
      emptied = true;
      break;
    }
    // ---- commit: the winner's box is the next epoch's newest selection, for everybody, straight from its warp's slot ----
    nb = ws->box[par][wround][wwarp];
    narea = ws->area[par][wround][wwarp];
    const bool winner = best == mk;   // this thread owns the selection (box indices are unique)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (cur[c] == -CUDART_INF_F) continue;
      const bool ev = (evaluated >> c) & 1u;
      if (winner && bestc == c) {
        sel_box[e] = box[c];      // (read by later epochs' chains: ordered by the next epoch's first barrier)
        sel_area[e] = area[c];
        out_row[e] = idx[c];
        out_score[e] = ev ? sp[c] : cur[c];
        cur[c] = -CUDART_INF_F;
      } else if (ev && heap_key(cur[c], idx[c]) > mk) {
        cur[c] = sp[c];   // popped before the winner: its decayed score (-inf when it died), looked at up to here
        beg[c] = e;
      }
    }
    last = key_to_float((uint32_t)(mk >> 32));
    nsel = e + 1;
    if (p.debug && blockIdx.x == 0 && tid == 0) {
      const long long t3 = clock64();
      g_nms_dbg[1] += (unsigned long long)(t1 - t0);
      g_nms_dbg[2] += (unsigned long long)(t2 - t1);
      g_nms_dbg[3] += (unsigned long long)(t3 - t2);
    }
  }
  // (emptied with everything evaluated: every evaluated candidate died - nothing left to commit)
  nsel_out = nsel;
  last_out = last;
  emptied_out = emptied;
}

// ---------------------------------------------------------------------------------------------------------------------
// one CTA per image: select the best scores (at most `cap`) into shared memory, run the epoch loop, flag an unprovable
// truncation
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) nms_epoch_cta_kernel(const EpochParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n, cap = p.cap;
  const size_t cap_al = ((size_t)cap + 3) & ~(size_t)3;   // keeps every array 16-byte aligned
  State st;
  st.carve(reinterpret_cast<char*>(smem), cap_al, true);
  float4* sel_box = reinterpret_cast<float4*>(smem + cap_al * kStateBytes);        // [max_out]
  float* sel_area = reinterpret_cast<float*>(sel_box + p.max_out);                 // [max_out]
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(sel_area + p.max_out);            // [kHistBins]
  __shared__ double s_tbl[64];
  __shared__ unsigned long long s_slots[64];
  if (threadIdx.x == 0) s_slots[63] = 0;
  __shared__ uint32_t s_kmin, s_kmax, s_above, s_count, s_bin, s_acc;
  const float* scores = p.scores + (size_t)s * n;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)s * n;
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  const float thr = p.score_thr;
  const long long t_start = p.debug ? clock64() : 0;
  if (tid < 64) s_tbl[tid] = kExp2Table[tid];
  if (tid == 0) {
    s_kmin = 0xffffffffu;
    s_kmax = 0u;
    s_above = 0u;
    s_count = 0u;
    if (s == 0 && p.cursor) *p.cursor = 0;
  }
  __syncthreads();

  // ---- key range and number of the candidates ----
  {
    uint32_t mn = 0xffffffffu, mx = 0u;
    int c = 0;
    for (int base = 0; base < n; base += kUnroll * kCtaThreads) {   // (kUnroll loads in flight per thread: the pass is L2-latency bound)
      float v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int j = base + u * kCtaThreads + tid;
        v[u] = j < n ? scores[j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int j = base + u * kCtaThreads + tid;
        if (j < n && v[u] > thr) {
          const uint32_t k = ordered_key(v[u]);
          mn = min(mn, k);
          mx = max(mx, k);
          ++c;
        }
      }
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) {
      atomicMin(&s_kmin, mn);
      atomicMax(&s_kmax, mx);
      atomicAdd(&s_above, (uint32_t)c);
    }
  }
  __syncthreads();
  const int above = (int)s_above;
  // ---- the cut: candidates = keys > cut, at most cap of them (adaptive histogram, refined while the boundary bin
  //      holds more than a quarter of the capacity) ----
  uint32_t cut = 0;
  bool truncated = false;
  if (above > cap) {
    truncated = true;
    uint32_t lo = s_kmin, hi = s_kmax;   // inclusive key range still undecided
    uint32_t taken = 0;                  // keys above `hi` (all accepted)
    for (int level = 0; level < 4; ++level) {
      const uint32_t span = hi - lo;
      const int shift = max(0, 32 - __clz(span | 1u) - 11);   // (span >> shift) < 2048
      for (int i = tid; i < kHistBins; i += kCtaThreads) s_hist[i] = 0;
      __syncthreads();
      {
        uint32_t run_bin = 0xffffffffu, run = 0;   // per-thread run-length aggregation (ties / flat score maps)
        for (int base = 0; base < n; base += kUnroll * kCtaThreads) {
          float v[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const int j = base + u * kCtaThreads + tid;
            v[u] = j < n ? scores[j] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const int j = base + u * kCtaThreads + tid;
            const uint32_t k = ordered_key(v[u]);
            if (j < n && v[u] > thr && k >= lo && k <= hi) {
              const uint32_t bin = (k - lo) >> shift;
              if (bin != run_bin) {
                if (run) atomicAdd(&s_hist[run_bin], run);
                run_bin = bin;
                run = 0;
              }
              ++run;
            }
          }
        }
        if (run) atomicAdd(&s_hist[run_bin], run);
      }
      __syncthreads();
      // warp 0: boundary bin = the first bin from the top at which the accepted count would exceed the capacity
      if (warp == 0) {
        const uint32_t room = (uint32_t)cap - taken;
        constexpr int per = kHistBins / 32;
        uint32_t mine = 0;
        for (int i = 0; i < per; ++i) mine += s_hist[kHistBins - 1 - (lane * per + i)];
        uint32_t incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += o;
        }
        const uint32_t excl = incl - mine;
        if (excl <= room && incl > room) {
          uint32_t acc = excl;
          for (int i = 0; i < per; ++i) {
            const int b = kHistBins - 1 - (lane * per + i);
            const uint32_t h = s_hist[b];
            if (acc + h > room) {
              s_bin = (uint32_t)b;
              s_acc = acc;
              break;
            }
            acc += h;
          }
        }
      }
      __syncthreads();
      const uint32_t bin = s_bin, acc = s_acc;
      __syncthreads();
      taken += acc;
      const uint32_t bin_lo = lo + (bin << shift);
      const uint32_t bin_hi = shift == 0 ? bin_lo : min(hi, bin_lo + ((1u << shift) - 1u));
      cut = bin_hi;   // keys in and below the boundary bin are excluded
      if (shift == 0 || taken * 4u >= (uint32_t)cap * 3u) break;
      lo = bin_lo;
      hi = bin_hi;
    }
  }
  // ---- compaction into shared memory (any order: the epoch loop ties by box index) ----
  for (int base = 0; base < n; base += kUnroll * kCtaThreads) {
    float v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int j = base + u * kCtaThreads + tid;
      v[u] = j < n ? scores[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int j = base + u * kCtaThreads + tid;
      const float sc = v[u];
      const bool ok = j < n && sc > thr && (!truncated || ordered_key(sc) > cut);
      const uint32_t act = __ballot_sync(0xffffffffu, ok);
      if (act) {
        uint32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&s_count, (uint32_t)__popc(act));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (ok) {
          const int slot = (int)(slot0 + (uint32_t)__popc(act & ((1u << lane) - 1u)));
          float area;
          st.box[slot] = normalise(boxes[j], area);
          st.area[slot] = area;
          st.cur[slot] = sc;
          st.idx[slot] = j;
          st.meta[slot] = 0;
          reinterpret_cast<uint4*>(st.mask)[slot] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  __syncthreads();
  const int cnt = (int)s_count;
  if (p.debug && s == 0 && tid == 0) {
    g_nms_dbg[0] += (unsigned long long)(clock64() - t_start);
    g_nms_dbg[5] = (unsigned long long)cnt;
  }

  int nsel;
  float last;
  bool emptied;
  __shared__ WinSlots s_win;
  if (cnt <= kCtaThreads) epoch_loop_regs<1>(p, cnt, st, sel_box, sel_area, &s_win, s_tbl, out_row, out_score, nsel, last, emptied);
  else if (cnt <= 2 * kCtaThreads) epoch_loop_regs<2>(p, cnt, st, sel_box, sel_area, &s_win, s_tbl, out_row, out_score, nsel, last, emptied);
  else {  // (more candidates than two per thread: state stays in shared memory)
    for (int j = tid; j < cnt; j += kCtaThreads) {
      st.meta[j] = 0;
      reinterpret_cast<uint4*>(st.mask)[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    epoch_loop(p, cnt, st, sel_box, sel_area, s_slots, s_tbl, out_row, out_score, nsel, last, emptied);
  }
  __syncthreads();
  for (int i = nsel + tid; i < p.max_out; i += kCtaThreads) {
    out_row[i] = 0;
    out_score[i] = 0.f;
  }
  if (tid == 0) {
    p.valid[s] = nsel;
    if (p.flag) {
      // exact iff nothing excluded could have been popped: every selection scored above the cut (every excluded key is
      // at or below it)
      const float nx = key_to_float(cut);
      p.flag[s] = (truncated && (emptied || nsel == 0 || !(last > nx))) ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// exact redo of the flagged images over all n candidates: the same loop, state in global memory (L2 resident).  A few
// persistent CTAs (one state slot each) walk the flag array.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads, 1) nms_epoch_full_kernel(const EpochParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x;
  const int n = p.n;
  float4* sel_box = reinterpret_cast<float4*>(smem);               // [max_out]
  float* sel_area = reinterpret_cast<float*>(sel_box + p.max_out);  // [max_out]
  __shared__ double s_tbl[64];
  __shared__ unsigned long long s_slots[64];
  if (threadIdx.x == 0) s_slots[63] = 0;
  __shared__ int s_seg;
  if (tid < 64) s_tbl[tid] = kExp2Table[tid];
  State st;
  const size_t n_al = ((size_t)n + 3) & ~(size_t)3;
  st.carve(p.g_state + (size_t)blockIdx.x * n_al * kStateBytes, n_al, false);
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      int s = atomicAdd(p.cursor, 1);
      while (s < p.segments && !p.flag[s]) s = atomicAdd(p.cursor, 1);
      s_seg = s;
    }
    __syncthreads();
    const int s = s_seg;
    if (s >= p.segments) break;
    const float* scores = p.scores + (size_t)s * n;
    const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)s * n;
    int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
    float* out_score = p.sel_scores + (size_t)s * p.max_out;
    for (int j = tid; j < n; j += kCtaThreads) {
      const float sc = scores[j];
      float area;
      st.box[j] = normalise(boxes[j], area);
      st.area[j] = area;
      st.cur[j] = sc > p.score_thr ? sc : -CUDART_INF_F;
      st.meta[j] = 0;
      reinterpret_cast<uint4*>(st.mask)[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    int nsel;
    float last;
    bool emptied;
    epoch_loop(p, n, st, sel_box, sel_area, s_slots, s_tbl, out_row, out_score, nsel, last, emptied);
    __syncthreads();
    for (int i = nsel + tid; i < p.max_out; i += kCtaThreads) {
      out_row[i] = 0;
      out_score[i] = 0.f;
    }
    if (tid == 0) p.valid[s] = nsel;
  }
}

}  // namespace

int udal_nms_debug = 0;
extern "C" int udal_nms_debug_read(unsigned long long* out8, int reset) {
  if (cudaMemcpyFromSymbol(out8, g_nms_dbg, sizeof(g_nms_dbg)) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[8] = {0};
    if (cudaMemcpyToSymbol(g_nms_dbg, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
// ---------------------------------------------------------------------------------------------------------------------
// Per-class NMS (postprocess.py:624-716): the same epoch formulation, one CTA of 256 threads per (image, class) segment of
// the top-k candidates (up to four candidates per thread, state in registers).  The candidates of a segment arrive in
// canonical order (score descending), the heap's tie rule is the rank inside the segment - exactly what TF sees when the
// reference calls the op on the class subset.  Segments with more candidates than 4 x 256 are left to the one-warp kernel
// of nms.cu (valid[s] = -1 marks them).  Measured (49 104 anchors x 10 classes, top-k 5000, gaussian): a segment takes
// ~0.28 ms here against ~0.8 ms in the one-warp kernel (one pop after the other, each a warp-wide arg-max plus a decay
// chain) but occupies eight warps instead of one: postprocess_per_class at B = 1 0.91 -> 0.73 ms, at B = 16 (160 segments)
// 1.0 -> 1.7 ms, at B = 64 (640 segments, two waves) 1.55 -> 2.78 ms.  Hard NMS in sorted order is a plain greedy pass in the
// one-warp kernel (0.3 ms against 2.3).  Hence the launcher below takes soft NMS of a few segments (single images: the
// latency case); everything else stays with nms.cu.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSegThreads = 256, kSegCap = 4 * kSegThreads;
struct SegParams {
  EpochParams e;              // thresholds, max_out (boxes / scores: [images, img_stride, ...])
  const int32_t* cand_idx;    // candidate j of segment s -> row in its image: cand_idx[start + j]
  const int32_t* seg_start;   // [S]
  const int32_t* seg_count;   // [S]
  int segs_per_image;
  long long img_stride;
  int32_t* sel_rank;          // [S,max_out] rank inside the segment (nullable); e.sel_row = row in the image
};

__global__ void __launch_bounds__(kSegThreads) nms_epoch_seg_kernel(const SegParams q) {
  const EpochParams& p = q.e;
  extern __shared__ __align__(16) uint8_t smem[];
  const int s = blockIdx.x, tid = threadIdx.x;
  const int cnt = q.seg_count[s];
  if (cnt > kSegCap) {   // (block-uniform)
    if (tid == 0) p.valid[s] = -1;
    return;
  }
  const int image = s / q.segs_per_image;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)image * q.img_stride;
  const float* scores = p.scores + (size_t)image * q.img_stride;
  const int32_t* cidx = q.cand_idx + q.seg_start[s];
  const size_t cap_al = ((size_t)kSegCap + 3) & ~(size_t)3;
  State st;
  st.carve(reinterpret_cast<char*>(smem), cap_al, true);
  float4* sel_box = reinterpret_cast<float4*>(smem + cap_al * kStateBytes);   // [max_out]
  float* sel_area = reinterpret_cast<float*>(sel_box + p.max_out);            // [max_out]
  __shared__ double s_tbl[64];
  __shared__ WinSlots s_win;
  __shared__ int32_t s_rank[128];
  if (tid < 64) s_tbl[tid] = kExp2Table[tid];
  for (int i = tid; i < (int)(sizeof(WinSlots) / 4); i += kSegThreads) reinterpret_cast<uint32_t*>(&s_win)[i] = 0u;
  for (int j = tid; j < cnt; j += kSegThreads) {
    float area;
    const int row = cidx[j];
    const float sc = scores[row];
    st.box[j] = normalise(boxes[row], area);
    st.area[j] = area;
    st.cur[j] = sc > p.score_thr ? sc : -CUDART_INF_F;
    st.idx[j] = j;
  }
  __syncthreads();
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  int nsel;
  float last;
  bool emptied;
  // (the loop writes the winners' ranks; they are translated to image rows below)
  if (cnt <= kSegThreads) epoch_loop_regs<1, kSegThreads>(p, cnt, st, sel_box, sel_area, &s_win, s_tbl, s_rank, out_score, nsel, last, emptied);
  else if (cnt <= 2 * kSegThreads) epoch_loop_regs<2, kSegThreads>(p, cnt, st, sel_box, sel_area, &s_win, s_tbl, s_rank, out_score, nsel, last, emptied);
  else epoch_loop_regs<4, kSegThreads>(p, cnt, st, sel_box, sel_area, &s_win, s_tbl, s_rank, out_score, nsel, last, emptied);
  __syncthreads();
  int32_t* out_rank = q.sel_rank ? q.sel_rank + (size_t)s * p.max_out : nullptr;
  for (int i = tid; i < p.max_out; i += kSegThreads) {
    if (i < nsel) {
      const int r = s_rank[i];
      out_row[i] = cidx[r];
      if (out_rank) out_rank[i] = r;
    } else {
      out_row[i] = 0;
      if (out_rank) out_rank[i] = 0;
      out_score[i] = 0.f;
    }
  }
  if (tid == 0) p.valid[s] = nsel;
}

int udal_nms_seg = 1;  // 0: per-class NMS through the one-warp-per-segment kernel of nms.cu only (comparison path)

// returns 1 in *handled when the segments were enqueued here (segments marked valid = -1 still need the one-warp kernel)
int udal_nms_epoch_segments(udal_ctx* ctx, const float* boxes, const float* scores, const int32_t* cand_idx, const int32_t* seg_start,
                            const int32_t* seg_count, int segments, int segs_per_image, int64_t img_stride, int32_t* sel_row,
                            int32_t* sel_rank, float* sel_scores, int32_t* valid, int* handled) {
  const udal_config& c = ctx->cfg;
  *handled = 0;
  if (!udal_nms_seg || c.max_output_size > 128 || !seg_start || !seg_count) return UDAL_OK;
  int dev = 0, sms = 0;
  UDAL_CUDA(cudaGetDevice(&dev));
  UDAL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!(c.nms_sigma_tf > 0.f) || 4 * segments > sms) return UDAL_OK;   // latency helper: soft NMS of one to three images
  SegParams q;
  memset(&q, 0, sizeof(q));
  q.e.boxes = boxes;
  q.e.scores = scores;
  q.e.segments = segments;
  q.e.max_out = c.max_output_size;
  q.e.iou_thr = c.nms_iou_thresh;
  q.e.score_thr = c.nms_score_thresh;
  q.e.soft = c.nms_sigma_tf > 0.f;
  q.e.scale = q.e.soft ? (-0.5f / c.nms_sigma_tf) : 0.f;
  q.e.variant_old = c.nms_variant_old;
  q.e.sel_row = sel_row;
  q.e.sel_scores = sel_scores;
  q.e.valid = valid;
  q.cand_idx = cand_idx;
  q.seg_start = seg_start;
  q.seg_count = seg_count;
  q.segs_per_image = segs_per_image;
  q.img_stride = img_stride;
  q.sel_rank = sel_rank;
  const size_t smem = (((size_t)kSegCap + 3) & ~(size_t)3) * kStateBytes + (size_t)q.e.max_out * 20;
  UDAL_CUDA(cudaFuncSetAttribute(nms_epoch_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_epoch_seg_kernel<<<segments, kSegThreads, smem, ctx->stream>>>(q);
  UDAL_CHECK_LAUNCH(ctx);
  *handled = 1;
  return UDAL_OK;
}

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
  UDAL_REQUIRE(p.max_out <= 128, "udal_nms_epoch: max_output_size %d > 128", p.max_out);
  int cap = udal_nms_prefilter_k(ctx, n);
  if (cap < 1) cap = 1;
  if (cap > 3072) cap = 3072;
  p.cap = cap;
  p.iou_thr = c.nms_iou_thresh;
  p.score_thr = c.nms_score_thresh;
  p.soft = c.nms_sigma_tf > 0.f;
  p.scale = p.soft ? (-0.5f / c.nms_sigma_tf) : 0.f;
  p.variant_old = c.nms_variant_old;
  p.sel_row = sel_idx;
  p.sel_scores = sel_scores;
  p.valid = valid;
  p.debug = udal_nms_debug;
  const bool may_truncate = cap < n;
  if (may_truncate) {
    char* scr;
    const size_t state = (size_t)kFullSlots * (((size_t)n + 3) & ~(size_t)3) * kStateBytes;
    UDAL_TRY(udal_scratch_get(ctx, SCR_NMS_B, state + (size_t)segments * 4 + 16, (void**)&scr));
    p.g_state = scr;
    p.flag = (int32_t*)(scr + state);
    p.cursor = p.flag + segments;
  }
  const size_t smem = (((size_t)cap + 3) & ~(size_t)3) * kStateBytes + (size_t)p.max_out * 20 + kHistBins * 4;
  UDAL_REQUIRE(smem <= 200 * 1024, "nms: %d candidates x max_output_size %d do not fit in shared memory", cap, p.max_out);
  UDAL_CUDA(cudaFuncSetAttribute(nms_epoch_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_epoch_cta_kernel<<<segments, kCtaThreads, smem, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  if (may_truncate) {
    const size_t smem2 = (size_t)p.max_out * 20;
    const int grid = segments < kFullSlots ? segments : kFullSlots;
    nms_epoch_full_kernel<<<grid, kCtaThreads, smem2, ctx->stream>>>(p);
    UDAL_CHECK_LAUNCH(ctx);
  }
  return UDAL_OK;
}
