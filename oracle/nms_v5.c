/* TEST INFRASTRUCTURE ONLY - CPU oracle, never linked into the product library.
 *
 * Restatement of the CPU kernel behind tf.raw_ops.NonMaxSuppressionV5
 * (tensorflow==2.10.0, core/kernels/image/non_max_suppression_op.cc; third-party, not
 * vendored in the reference and not installable offline => PARITY UNPINNED).  Reference call
 * site: src/postprocess.py:392-400.
 *
 * Published algorithm: candidates with score > score_threshold go into a max-heap ordered by
 * (score desc, index asc).  The top candidate is decayed by every already-selected box it has
 * not yet been compared with, newest selection first; it is selected if its score did not
 * change, re-inserted if it changed but is still above the threshold, dropped otherwise.
 *
 * Arithmetic pinned by this oracle (and matched bit for bit by the CUDA kernels):
 *   iou      fp32, corner order normalised, 0 if either area <= 0, no "+1" pixel convention
 *   weight   fp32( exp( fp64( (scale * iou) * iou ) ) ), scale = -0.5 / soft_nms_sigma in fp32
 *            (TF calls std::exp(float); rounding the fp64 exp is the correctly rounded fp32
 *            value, which glibc expf returns in all but ~0.4% of arguments, 1 ulp off there)
 *   variant  0 = "new" (TF >= ~2.4, the pinned 2.10): weight applies for every iou in soft
 *                mode, hard break only when sigma == 0 and iou >  threshold
 *            1 = "old" (TF <= ~2.3): weight = 0 above the threshold and a hard break when
 *                iou >= threshold in both modes
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float score;
  int idx;
  int begin;
} cand_t;

static inline int cand_before(const cand_t* a, const cand_t* b) {
  /* 1 if a must be popped before b */
  if (a->score != b->score) return a->score > b->score;
  return a->idx < b->idx;
}

static void sift_up(cand_t* h, int i) {
  cand_t x = h[i];
  while (i > 0) {
    int p = (i - 1) >> 1;
    if (!cand_before(&x, &h[p])) break;
    h[i] = h[p];
    i = p;
  }
  h[i] = x;
}

static void sift_down(cand_t* h, int n, int i) {
  cand_t x = h[i];
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m;
    if (l >= n) break;
    m = (r < n && cand_before(&h[r], &h[l])) ? r : l;
    if (!cand_before(&h[m], &x)) break;
    h[i] = h[m];
    i = m;
  }
  h[i] = x;
}

float udal_oracle_iou(const float* a, const float* b) {
  float ay0 = fminf(a[0], a[2]), ax0 = fminf(a[1], a[3]);
  float ay1 = fmaxf(a[0], a[2]), ax1 = fmaxf(a[1], a[3]);
  float by0 = fminf(b[0], b[2]), bx0 = fminf(b[1], b[3]);
  float by1 = fmaxf(b[0], b[2]), bx1 = fmaxf(b[1], b[3]);
  float area_a = (ay1 - ay0) * (ax1 - ax0);
  float area_b = (by1 - by0) * (bx1 - bx0);
  if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f;
  float iy0 = fmaxf(ay0, by0), ix0 = fmaxf(ax0, bx0);
  float iy1 = fminf(ay1, by1), ix1 = fminf(ax1, bx1);
  float ih = fmaxf(iy1 - iy0, 0.0f), iw = fmaxf(ix1 - ix0, 0.0f);
  float inter = ih * iw;
  return inter / (area_a + area_b - inter);
}

/* 0 (default, what the CUDA kernels match): fp32(exp(fp64)), the correctly rounded value.
 * 1: the C library's expf, i.e. what TF's std::exp(float) executes on a glibc host.  Only the
 *    differential test (tests/test_oracle_extra.py) sets it, to put a number on the unpinned rounding. */
int udal_oracle_weight_mode = 0;

float udal_oracle_soft_weight(float scale, float iou) {
  float arg = (scale * iou) * iou;
  if (udal_oracle_weight_mode == 1) return expf(arg);
  return (float)exp((double)arg);
}

/* returns the number of valid selections; sel_idx / sel_scores must hold max_out entries and are
 * zero padded past the valid count (callers slice when pad_to_max_output_size is false). */
int udal_oracle_nms_v5(const float* boxes, const float* scores, int n, int max_out, float iou_thr,
                       float score_thr, float soft_sigma, int variant_old, int* sel_idx,
                       float* sel_scores) {
  int heap_n = 0, nsel = 0;
  cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
  const int soft = soft_sigma > 0.0f;
  const float scale = soft ? (-0.5f / soft_sigma) : 0.0f;
  for (int i = 0; i < n; ++i) {
    if (scores[i] > score_thr) {
      heap[heap_n].score = scores[i];
      heap[heap_n].idx = i;
      heap[heap_n].begin = 0;
      sift_up(heap, heap_n);
      ++heap_n;
    }
  }
  memset(sel_idx, 0, sizeof(int) * (size_t)max_out);
  memset(sel_scores, 0, sizeof(float) * (size_t)max_out);
  while (nsel < max_out && heap_n > 0) {
    cand_t c = heap[0];
    const float original = c.score;
    heap[0] = heap[--heap_n];
    if (heap_n > 0) sift_down(heap, heap_n, 0);
    int hard = 0;
    for (int j = nsel - 1; j >= c.begin; --j) {
      const float u = udal_oracle_iou(boxes + 4 * (size_t)c.idx, boxes + 4 * (size_t)sel_idx[j]);
      float w = udal_oracle_soft_weight(scale, u);
      if (variant_old) {
        if (!(u <= iou_thr)) w = 0.0f;
        c.score *= w;
        if (u >= iou_thr) { hard = 1; break; }
      } else {
        if (!(soft || u <= iou_thr)) w = 0.0f;
        c.score *= w;
        if (!soft && u > iou_thr) { hard = 1; break; }
      }
      if (c.score <= score_thr) break;
    }
    c.begin = nsel;
    if (!hard) {
      if (c.score == original) {
        sel_idx[nsel] = c.idx;
        sel_scores[nsel] = c.score;
        ++nsel;
      } else if (c.score > score_thr) {
        heap[heap_n] = c;
        sift_up(heap, heap_n);
        ++heap_n;
      }
    }
  }
  free(heap);
  return nsel;
}

/* batch helper used by the CPU baseline: images are independent */
void udal_oracle_nms_v5_batch(const float* boxes, const float* scores, int batch, int n, int max_out,
                              float iou_thr, float score_thr, float soft_sigma, int variant_old,
                              int* sel_idx, float* sel_scores, int* valid) {
  for (int b = 0; b < batch; ++b) {
    valid[b] = udal_oracle_nms_v5(boxes + (size_t)b * n * 4, scores + (size_t)b * n, n, max_out,
                                  iou_thr, score_thr, soft_sigma, variant_old,
                                  sel_idx + (size_t)b * max_out, sel_scores + (size_t)b * max_out);
  }
}
