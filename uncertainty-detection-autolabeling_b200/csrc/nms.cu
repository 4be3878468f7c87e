// K4: tf.raw_ops.NonMaxSuppressionV5 semantics (hard and gaussian soft NMS) on the device.
//
// Replaces the TF kernel called at reference src/postprocess.py:392-400 (nms), used per image by
// postprocess_global (:472-621) and per (image, class) by per_class_nms (:624-716).
//
// The TF kernel is a lazy max-heap algorithm: pop the best candidate, decay it by the boxes
// selected since it was last looked at (newest first), select it if its score did not change,
// re-insert it otherwise.  Selected scores are non-increasing, so candidates are consumed in the
// order of their ORIGINAL score: the kernels below take the candidates of a segment already
// sorted (score descending, index ascending - the heap's tie rule) and keep only the re-inserted
// ones in an explicit set.  One warp owns one segment:
//   hard mode : 32 candidates at a time; every lane tests its candidate against the selected
//               boxes, a ballot loop resolves the suppression inside the chunk.
//   soft mode : exact emulation of the lazy order (the fp32 product of decay weights is taken in
//               the same order as TF: newest selection first, batch by batch); lanes compute the
//               IoUs / weights of up to 32 selected boxes in parallel, a ballot picks the weights
//               that are not exactly 1.
// A segment may be a prefix of the full candidate list (global NMS over all anchors is run on
// the top-K scores).  The result is provably identical to the untruncated run when every
// candidate that was popped scored above the best excluded one; otherwise the segment is flagged
// and re-done by nms_v5_full_kernel, a one-CTA-per-segment emulation over all candidates.
//
// Arithmetic = oracle/nms_v5.c: fp32 IoU without "+1", weight = fp32(exp(fp64((scale*u)*u))).
#include <math_constants.h>

#include "fast_math64.cuh"
#include "udal_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kStageCap = 2048;  // candidates a staged (shared-memory resident) segment may hold

struct NmsParams {
  const float* boxes;        // [images, img_stride, 4]
  const float* scores;       // [images, img_stride]
  const int32_t* cand_idx;   // candidate j of segment s -> row in its image: cand_idx[start + j]
  const int32_t* seg_start;  // [S] offsets into cand_idx (null: s * seg_n)
  const int32_t* seg_count;  // [S] candidates per segment (null: seg_n)
  const float* next_score;   // [S] best original score among excluded candidates (null: none)
  int next_stride;           // next_score[s * next_stride]
  int segments, seg_n, segs_per_image;
  int64_t img_stride;
  int max_out;
  float iou_thr, score_thr, scale;  // scale = -0.5 / sigma (soft) or 0
  int soft, variant_old;
  // outputs
  int32_t* sel_row;    // [S,max_out] row in the image (cand_idx value), zero padded
  int32_t* sel_rank;   // [S,max_out] rank inside the segment (nullable)
  float* sel_scores;   // [S,max_out]
  int32_t* valid;      // [S]
  int32_t* flag;       // [S] 1 = truncated run not provably exact (nullable)
  int only_marked;     // 1: only the segments the per-segment epoch kernel left (valid[s] == -1, nms_cta.cu)
  // soft-mode re-insertion set; a segment's slice starts at its candidate offset
  float* r_score;
  int32_t* r_rank;
  int32_t* r_begin;
};

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
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__device__ __forceinline__ float soft_weight(float scale, float u) {
  return (float)exp((double)__fmul_rn(__fmul_rn(scale, u), u));
}
// same value through the table-based exp (rel. error ~1e-15: rounds to the same fp32 except ~1e-8 of the
// arguments); exp(0) = 1 exactly, which is what makes non-overlapping boxes free
__device__ __forceinline__ float soft_weight_fast(float scale, float u, const double* tbl) {
  if (u == 0.f) return 1.f;
  return (float)exp_fast((double)__fmul_rn(__fmul_rn(scale, u), u), tbl);
}

__device__ __forceinline__ float4 shfl_box(float4 v, int src) {
  float4 r;
  r.x = __shfl_sync(0xffffffffu, v.x, src);
  r.y = __shfl_sync(0xffffffffu, v.y, src);
  r.z = __shfl_sync(0xffffffffu, v.z, src);
  r.w = __shfl_sync(0xffffffffu, v.w, src);
  return r;
}

// true if (s1, r1) is popped before (s2, r2)
__device__ __forceinline__ bool before(float s1, int r1, float s2, int r2) {
  return s1 > s2 || (s1 == s2 && r1 < r2);
}

// STAGED: one warp per CTA, the segment's candidates (score, box) and the re-insertion set live in
// shared memory, so a pop costs shared-memory latency instead of chains of dependent global loads.
template <bool STAGED>
__global__ void __launch_bounds__((STAGED ? 1 : kWarpsPerBlock) * 32) nms_v5_sorted_kernel(const NmsParams p) {
  constexpr int WPB = STAGED ? 1 : kWarpsPerBlock;
  extern __shared__ float4 smem_boxes[];  // [WPB][max_out] (+ staged candidate arrays)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * WPB + warp;
  if (s >= p.segments) return;
  if (p.only_marked && p.valid[s] != -1) return;   // (warp-uniform; no block barrier follows)
  float4* sel_box = smem_boxes + (size_t)warp * p.max_out;
  float4* s_box = smem_boxes + (size_t)WPB * p.max_out;                    // [kStageCap]
  float* s_score = reinterpret_cast<float*>(s_box + kStageCap);            // [kStageCap]
  float* s_rscore = s_score + kStageCap;                                   // re-insertion set
  int32_t* s_rrank = reinterpret_cast<int32_t*>(s_rscore + kStageCap);
  int32_t* s_rbegin = s_rrank + kStageCap;
  const int image = s / p.segs_per_image;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)image * p.img_stride;
  const float* scores = p.scores + (size_t)image * p.img_stride;
  const int start = p.seg_start ? p.seg_start[s] : s * p.seg_n;
  const int n = p.seg_count ? p.seg_count[s] : p.seg_n;
  const int32_t* cidx = p.cand_idx + start;
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  int32_t* out_rank = p.sel_rank ? p.sel_rank + (size_t)s * p.max_out : nullptr;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  const float thr = p.score_thr;
  int nsel = 0;
  float last_pop = CUDART_INF_F;
  bool emptied = false;
  __shared__ double s_tbl[WPB][64];
  s_tbl[warp][lane] = kExp2Table[lane];
  s_tbl[warp][lane + 32] = kExp2Table[lane + 32];
  __syncwarp();
  if (STAGED) {
    for (int j = lane; j < n; j += 32) {
      const int row = cidx[j];
      s_score[j] = scores[row];
      s_box[j] = boxes[row];
    }
    __syncwarp();
  }
  auto score_at = [&](int j) -> float { return STAGED ? s_score[j] : scores[cidx[j]]; };
  auto box_at = [&](int j) -> float4 { return STAGED ? s_box[j] : boxes[cidx[j]]; };

  if (!p.soft) {
    // ---------------- hard NMS: suppressed iff IoU with a selected box > (>=, old) threshold
    int j0 = 0;
    for (; j0 < n && nsel < p.max_out; j0 += 32) {
      const int j = j0 + lane;
      bool alive = false;
      float4 box = make_float4(0, 0, 0, 0);
      float sc = 0.f;
      int row = 0;
      if (j < n) {
        sc = score_at(j);
        alive = sc > thr;
        box = box_at(j);
      }
      for (int q = 0; q < nsel && alive; ++q) {
        const float u = iou_v5(box, sel_box[q]);
        if (p.variant_old ? (u >= p.iou_thr) : (u > p.iou_thr)) alive = false;
      }
      unsigned int mask = __ballot_sync(0xffffffffu, alive);
      while (mask && nsel < p.max_out) {
        const int i = __ffs(mask) - 1;
        const float4 bi = shfl_box(box, i);
        if (lane == i) {
          row = cidx[j];
          sel_box[nsel] = box;
          out_row[nsel] = row;
          if (out_rank) out_rank[nsel] = j;
          out_score[nsel] = sc;
        }
        if (alive && lane > i) {
          const float u = iou_v5(box, bi);
          if (p.variant_old ? (u >= p.iou_thr) : (u > p.iou_thr)) alive = false;
        }
        ++nsel;
        __syncwarp();
        mask = __ballot_sync(0xffffffffu, alive) & ~((2u << i) - 1u);
        if (i == 31) mask = 0;
      }
      // the list is sorted: once the chunk's last score is at or below the threshold we are done
      if (!(__shfl_sync(0xffffffffu, sc, 31) > thr)) break;
    }
    // hard mode never re-inserts, so a truncated list is exact whenever it filled max_out
    emptied = nsel < p.max_out;
    last_pop = CUDART_INF_F;
  } else {
    // ---------------- soft NMS: exact lazy-heap emulation
    float* r_score = STAGED ? s_rscore : p.r_score + start;
    int32_t* r_rank = STAGED ? s_rrank : p.r_rank + start;
    int32_t* r_begin = STAGED ? s_rbegin : p.r_begin + start;
    int next = 0, rn = 0;
    while (nsel < p.max_out) {
      // best re-inserted candidate
      float bs = -CUDART_INF_F;
      int brank = 0x7fffffff, bslot = -1;
      for (int i = lane; i < rn; i += 32) {
        const float sc = r_score[i];
        const int rk = r_rank[i];
        if (bslot < 0 || before(sc, rk, bs, brank)) {
          bs = sc;
          brank = rk;
          bslot = i;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, off);
        const int ork = __shfl_xor_sync(0xffffffffu, brank, off);
        const int osl = __shfl_xor_sync(0xffffffffu, bslot, off);
        if (osl >= 0 && (bslot < 0 || before(os, ork, bs, brank))) {
          bs = os;
          brank = ork;
          bslot = osl;
        }
      }
      // next never-popped candidate (sorted list)
      float as = -CUDART_INF_F;
      bool have_a = false;
      if (next < n) {
        as = score_at(next);
        have_a = as > thr;
      }
      if (!have_a && bslot < 0) {
        emptied = true;
        break;
      }
      const bool take_a = have_a && (bslot < 0 || before(as, next, bs, brank));
      int rank, begin;
      float s0;
      if (take_a) {
        rank = next;
        s0 = as;
        begin = 0;
        ++next;
      } else {
        rank = brank;
        s0 = bs;
        begin = r_begin[bslot];
        __syncwarp();
        if (lane == 0) {  // swap-remove
          r_score[bslot] = r_score[rn - 1];
          r_rank[bslot] = r_rank[rn - 1];
          r_begin[bslot] = r_begin[rn - 1];
        }
        --rn;
        __syncwarp();
      }
      last_pop = s0;
      const float4 cbox = box_at(rank);
      float sc = s0;
      bool hard = false, stop = false;
      for (int top = nsel - 1; top >= begin && !stop; top -= 32) {
        const int q = top - lane;
        float w = 1.f;
        bool hq = false;
        if (q >= begin) {
          const float u = iou_v5(cbox, sel_box[q]);
          w = soft_weight_fast(p.scale, u, s_tbl[warp]);
          if (p.variant_old) {
            if (!(u <= p.iou_thr)) w = 0.f;
            hq = u >= p.iou_thr;
          }
        }
        unsigned int mask = __ballot_sync(0xffffffffu, (w != 1.f) || hq);
        while (mask) {
          const int i = __ffs(mask) - 1;
          mask &= mask - 1;
          const float wi = __shfl_sync(0xffffffffu, w, i);
          const bool hi = __shfl_sync(0xffffffffu, (int)hq, i) != 0;
          sc = __fmul_rn(sc, wi);
          if (hi) {
            hard = true;
            stop = true;
            break;
          }
          if (sc <= thr) {
            stop = true;
            break;
          }
        }
      }
      if (!hard) {
        if (sc == s0) {
          if (lane == 0) {
            sel_box[nsel] = cbox;
            out_row[nsel] = cidx[rank];
            if (out_rank) out_rank[nsel] = rank;
            out_score[nsel] = sc;
          }
          ++nsel;
          __syncwarp();
        } else if (sc > thr) {
          if (lane == 0) {
            r_score[rn] = sc;
            r_rank[rn] = rank;
            r_begin[rn] = nsel;
          }
          ++rn;
          __syncwarp();
        }
      }
    }
  }

  for (int i = nsel + lane; i < p.max_out; i += 32) {
    out_row[i] = 0;
    if (out_rank) out_rank[i] = 0;
    out_score[i] = 0.f;
  }
  if (lane == 0) {
    p.valid[s] = nsel;
    if (p.flag) {
      const float nx = p.next_score ? p.next_score[(size_t)s * p.next_stride] : -CUDART_INF_F;
      const bool truncated = p.next_score && nx > thr;
      p.flag[s] = (truncated && (emptied || !(last_pop > nx))) ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fallback / cross-check: one CTA per segment, all n candidates, no sorting.  Exact emulation:
// every pop is a block-wide arg-max over the current scores.
// ---------------------------------------------------------------------------------------------
constexpr int kFullThreads = 512;

struct NmsFullParams {
  const float* boxes;   // [S,n,4]
  const float* scores;  // [S,n]
  int segments, n, max_out;
  float iou_thr, score_thr, scale;
  int soft, variant_old;
  const int32_t* flag;  // run only where flag[s] != 0 (null: run everywhere)
  float* cur;           // [S,n] scratch
  int32_t* begin;       // [S,n] scratch
  int32_t* sel_row;
  float* sel_scores;
  int32_t* valid;
};

__global__ void __launch_bounds__(kFullThreads) nms_v5_full_kernel(const NmsFullParams p) {
  extern __shared__ float4 sel_box[];  // [max_out]
  __shared__ float red_s[kFullThreads / 32];
  __shared__ int red_i[kFullThreads / 32];
  __shared__ float best_s;
  __shared__ int best_i;
  __shared__ int sh_nsel;
  const int s = blockIdx.x;
  if (p.flag && !p.flag[s]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)s * p.n;
  const float* scores = p.scores + (size_t)s * p.n;
  float* cur = p.cur + (size_t)s * p.n;
  int32_t* beg = p.begin + (size_t)s * p.n;
  int32_t* out_row = p.sel_row + (size_t)s * p.max_out;
  float* out_score = p.sel_scores + (size_t)s * p.max_out;
  const float thr = p.score_thr;
  for (int i = tid; i < p.n; i += kFullThreads) {
    const float sc = scores[i];
    cur[i] = sc > thr ? sc : -CUDART_INF_F;
    beg[i] = 0;
  }
  if (tid == 0) sh_nsel = 0;
  __syncthreads();
  while (true) {
    const int nsel = sh_nsel;
    if (nsel >= p.max_out) break;
    float bs = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (int i = tid; i < p.n; i += kFullThreads) {
      const float sc = cur[i];
      if (sc > bs) {  // ascending i per thread: strict > keeps the smallest index on ties
        bs = sc;
        bi = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (os > bs || (os == bs && oi < bi)) {
        bs = os;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_s[warp] = bs;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kFullThreads / 32; ++w)
        if (red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi)) {
          bs = red_s[w];
          bi = red_i[w];
        }
      best_s = bs;
      best_i = bi;
      if (bs > -CUDART_INF_F) {
        const float4 cbox = boxes[bi];
        float sc = bs;
        bool hard = false;
        for (int q = nsel - 1; q >= beg[bi]; --q) {
          const float u = iou_v5(cbox, sel_box[q]);
          float w = soft_weight(p.scale, u);
          if (p.variant_old) {
            if (!(u <= p.iou_thr)) w = 0.f;
            sc = __fmul_rn(sc, w);
            if (u >= p.iou_thr) {
              hard = true;
              break;
            }
          } else {
            if (!(p.soft || u <= p.iou_thr)) w = 0.f;
            sc = __fmul_rn(sc, w);
            if (!p.soft && u > p.iou_thr) {
              hard = true;
              break;
            }
          }
          if (sc <= thr) break;
        }
        beg[bi] = nsel;
        cur[bi] = -CUDART_INF_F;
        if (!hard) {
          if (sc == bs) {
            sel_box[nsel] = cbox;
            out_row[nsel] = bi;
            out_score[nsel] = sc;
            sh_nsel = nsel + 1;
          } else if (sc > thr) {
            cur[bi] = sc;
          }
        }
      }
    }
    __syncthreads();
    if (!(best_s > -CUDART_INF_F)) break;
  }
  __syncthreads();
  const int nsel = sh_nsel;
  for (int i = nsel + tid; i < p.max_out; i += kFullThreads) {
    out_row[i] = 0;
    out_score[i] = 0.f;
  }
  if (tid == 0) p.valid[s] = nsel;
}

__global__ void fill_segments_kernel(int32_t* starts, int32_t* counts, int segments, int stride, int count) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < segments) {
    starts[s] = s * stride;
    counts[s] = count;
  }
}

}  // namespace

int udal_nms_post_unstaged = 0;  // 1: on the post stream of back-to-back udal_run calls use the kernel without the 64 KB staging area
                                 // (it can co-reside with the persistent head kernels; measured slower since those claim their
                                 // work items dynamically: 3.21 vs 3.09 ms per step)

// sorted-candidate NMS over generic segments (internal)
int udal_nms_sorted(udal_ctx* ctx, const float* boxes, const float* scores, const int32_t* cand_idx,
                    const int32_t* seg_start, const int32_t* seg_count, const float* next_score,
                    int next_stride, int segments, int seg_n, int segs_per_image, int64_t img_stride,
                    int64_t total_cand, int32_t* sel_row, int32_t* sel_rank, float* sel_scores,
                    int32_t* valid, int32_t* flag) {
  const udal_config& c = ctx->cfg;
  NmsParams p;
  memset(&p, 0, sizeof(p));
  p.boxes = boxes;
  p.scores = scores;
  p.cand_idx = cand_idx;
  p.seg_start = seg_start;
  p.seg_count = seg_count;
  p.next_score = next_score;
  p.next_stride = next_stride;
  p.segments = segments;
  p.seg_n = seg_n;
  p.segs_per_image = segs_per_image;
  p.img_stride = img_stride;
  p.max_out = c.max_output_size;
  p.iou_thr = c.nms_iou_thresh;
  p.score_thr = c.nms_score_thresh;
  p.soft = c.nms_sigma_tf > 0.f;
  p.scale = p.soft ? (-0.5f / c.nms_sigma_tf) : 0.f;
  p.variant_old = c.nms_variant_old;
  p.sel_row = sel_row;
  p.sel_rank = sel_rank;
  p.sel_scores = sel_scores;
  p.valid = valid;
  p.flag = flag;
  if (cand_idx && seg_start && seg_count && !next_score && !flag) {
    // per-class segments: one cooperative CTA per segment (nms_cta.cu); what it leaves (valid = -1) runs below
    int handled = 0;
    UDAL_TRY(udal_nms_epoch_segments(ctx, boxes, scores, cand_idx, seg_start, seg_count, segments, segs_per_image, img_stride, sel_row,
                                     sel_rank, sel_scores, valid, &handled));
    p.only_marked = handled;
  }
  if (p.soft) {
    char* scr;
    const size_t per = (size_t)total_cand;
    UDAL_TRY(udal_scratch_get(ctx, SCR_NMS_A, per * 12, (void**)&scr));
    p.r_score = (float*)scr;
    p.r_rank = (int32_t*)(scr + per * 4);
    p.r_begin = (int32_t*)(scr + per * 8);
  }
  // The staged variant (64 KB of shared memory per segment) is the fastest stand-alone; on the post stream of
  // back-to-back udal_run calls the kernel has to co-reside with the persistent head kernels of the next run,
  // which leave < 20 KB per SM: there the candidates stay in global memory (L1 / L2 hits) and the latency
  // hides behind the heads.
  const bool staged = seg_n <= kStageCap &&
                      !(ctx->in_run && ctx->run_pipelined && ctx->stream == ctx->post_stream && udal_nms_post_unstaged);
  if (staged) {
    const size_t smem = (size_t)p.max_out * sizeof(float4) + (size_t)kStageCap * (16 + 4 + 12);
    UDAL_REQUIRE(smem <= 200 * 1024, "max_output_size %d too large", p.max_out);
    UDAL_CUDA(cudaFuncSetAttribute(nms_v5_sorted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_v5_sorted_kernel<true><<<segments, 32, smem, ctx->stream>>>(p);
  } else {
    const size_t smem = (size_t)kWarpsPerBlock * p.max_out * sizeof(float4);
    UDAL_REQUIRE(smem <= 200 * 1024, "max_output_size %d too large", p.max_out);
    if (smem > 48 * 1024)
      UDAL_CUDA(cudaFuncSetAttribute(nms_v5_sorted_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (segments + kWarpsPerBlock - 1) / kWarpsPerBlock;
    nms_v5_sorted_kernel<false><<<blocks, kWarpsPerBlock * 32, smem, ctx->stream>>>(p);
  }
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

// full (unsorted, exact) NMS on [S,n]; runs only where flag != 0 when flag is given
int udal_nms_full(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n,
                  const int32_t* flag, int32_t* sel_row, float* sel_scores, int32_t* valid) {
  const udal_config& c = ctx->cfg;
  NmsFullParams p;
  memset(&p, 0, sizeof(p));
  p.boxes = boxes;
  p.scores = scores;
  p.segments = segments;
  p.n = n;
  p.max_out = c.max_output_size;
  p.iou_thr = c.nms_iou_thresh;
  p.score_thr = c.nms_score_thresh;
  p.soft = c.nms_sigma_tf > 0.f;
  p.scale = p.soft ? (-0.5f / c.nms_sigma_tf) : 0.f;
  p.variant_old = c.nms_variant_old;
  p.flag = flag;
  char* scr;
  const size_t per = (size_t)segments * n;
  UDAL_TRY(udal_scratch_get(ctx, SCR_NMS_B, per * 8, (void**)&scr));
  p.cur = (float*)scr;
  p.begin = (int32_t*)(scr + per * 4);
  p.sel_row = sel_row;
  p.sel_scores = sel_scores;
  p.valid = valid;
  const size_t smem = (size_t)p.max_out * sizeof(float4);
  if (smem > 40 * 1024)
    UDAL_CUDA(cudaFuncSetAttribute(nms_v5_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_v5_full_kernel<<<segments, kFullThreads, smem, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_nms_prefilter_k(const udal_ctx* ctx, int n) {
  int k = ctx->cfg.prefilter_k > 0 ? ctx->cfg.prefilter_k : 2047;
  if (k > 8191) k = 8191;
  if (k >= n) k = n;  // no truncation
  return k;
}

// Global NMS over [S,n] boxes/scores (unsorted) in two steps so that a pipelined udal_run can take the
// bandwidth-bound pre-filter on its main stream as soon as the scores exist and leave only the latency-bound
// selection to the post stream:
//   udal_nms_prefilter : exact top-K of the scores (+ the best excluded one) and the segment table
//   udal_nms_select    : sorted-candidate kernel + exact redo of flagged segments
// Everything is enqueued; no host synchronisation.
int udal_nms_prefilter(udal_ctx* ctx, const float* scores, int segments, int n, udal_nms_plan* plan) {
  UDAL_REQUIRE(segments > 0 && n > 0, "udal_nms_prefilter: bad sizes");
  const int kk = udal_nms_prefilter_k(ctx, n);   // candidates handed to the sorted kernel
  const int kq = kk < n ? kk + 1 : kk;          // top-k query size (one extra = best excluded)
  char* scr;
  const size_t per = (size_t)segments * kq;
  UDAL_TRY(udal_scratch_get(ctx, SCR_PRE_B, per * 8 + (size_t)segments * 4, (void**)&scr));
  plan->kk = kk;
  plan->kq = kq;
  plan->tk_idx = (int32_t*)scr;
  plan->tk_val = (float*)(scr + per * 4);
  plan->flag = (int32_t*)(scr + per * 8);
  UDAL_TRY(udal_launch_topk(ctx, scores, segments, n, kq, plan->tk_idx, plan->tk_val));
  // segment s: candidates tk_idx[s*kq .. s*kq+kk), rows index image s
  UDAL_TRY(udal_scratch_get(ctx, SCR_MISC, (size_t)segments * 8, (void**)&plan->starts));
  fill_segments_kernel<<<(segments + 255) / 256, 256, 0, ctx->stream>>>(plan->starts, plan->starts + segments, segments, kq, kk);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}

int udal_nms_select(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n, const udal_nms_plan& plan,
                    int32_t* sel_idx, float* sel_scores, int32_t* valid) {
  UDAL_REQUIRE(((uintptr_t)boxes & 15) == 0, "boxes must be 16-byte aligned");
  const int kk = plan.kk, kq = plan.kq;
  const float* next = kk < n ? plan.tk_val + kk : nullptr;
  UDAL_TRY(udal_nms_sorted(ctx, boxes, scores, plan.tk_idx, plan.starts, plan.starts + segments, next, kq, segments, kq,
                           1, n, (int64_t)segments * kq, sel_idx, nullptr, sel_scores, valid, plan.flag));
  if (kk < n) UDAL_TRY(udal_nms_full(ctx, boxes, scores, segments, n, plan.flag, sel_idx, sel_scores, valid));
  return UDAL_OK;
}

int udal_launch_nms_v5(udal_ctx* ctx, const float* boxes, const float* scores, int segments, int n,
                       int32_t* sel_idx, float* sel_scores, int32_t* valid) {
  const int max_out = ctx->cfg.max_output_size;
  UDAL_REQUIRE(segments > 0 && n >= 0, "udal_nms_v5: bad sizes");
  UDAL_REQUIRE(((uintptr_t)boxes & 15) == 0, "boxes must be 16-byte aligned");
  if (n == 0) {
    UDAL_CUDA(cudaMemsetAsync(sel_idx, 0, (size_t)segments * max_out * 4, ctx->stream));
    UDAL_CUDA(cudaMemsetAsync(sel_scores, 0, (size_t)segments * max_out * 4, ctx->stream));
    UDAL_CUDA(cudaMemsetAsync(valid, 0, (size_t)segments * 4, ctx->stream));
    return UDAL_OK;
  }
  // default: one cooperative CTA per image (nms_cta.cu); udal_nms_cta = 0 keeps the round-1 path (top-k pre-filter +
  // one warp per image) for comparison
  if (udal_nms_cta && max_out <= 128) return udal_nms_epoch(ctx, boxes, scores, segments, n, sel_idx, sel_scores, valid);
  udal_nms_plan plan;
  UDAL_TRY(udal_nms_prefilter(ctx, scores, segments, n, &plan));
  return udal_nms_select(ctx, boxes, scores, segments, n, plan, sel_idx, sel_scores, valid);
}
