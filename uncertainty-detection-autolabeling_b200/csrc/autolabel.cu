// calibrated-uncertainty application + auto-label threshold pass on the detections of
// postprocess_global, as an epilogue kernel (one CTA per image, one thread per detection row).
//
// Replaces (reference src/):
//   infer_model.py:585-595   entropy of stable_softmax(logits)              (utils_class.py:36-41)
//   utils_box.py:404-524     CalibrateBoxUncert.calibrate_boxuncert: temperature scaling / isotonic
//                            regression applied to the aleatoric box std; an sklearn IsotonicRegression
//                            (increasing, out_of_bounds="clip") at inference is a clipped piece-wise linear
//                            interpolation over its X_thresholds_ / y_thresholds_ table
//   utils_box.py:279-292     relativize_uncert (std / [h, w, h, w])
//   infer_model.py:688-691, 742-764   weighted sum of the selected uncertainties, threshold decision over
//                            the detections with score > min_score
#include <cuda_fp16.h>

#include "udal_common.cuh"

namespace {

struct AutoParams {
  const float* boxes;    // [B,M,box_stride]: box at 0..3, aleatoric std at albox_col..+3
  const float* scores;   // [B,M]
  const float* classes;  // [B,M,class_stride]: class id (1-based, as float) at column 0
  const float* logits;   // [B,M,C]
  int box_stride, albox_col, class_stride, C, M;
  udal_autolabel_params prm;
  float* entropy;      // [B,M]
  float* calib_albox;  // [B,M,4]
  float* rel_albox;    // [B,M,4]
  float* opt_uncert;   // [B,M]
  int32_t* decision;   // [B]
};

// sklearn IsotonicRegression.predict with out_of_bounds="clip": interp1d(kind="linear") over the knots, float64
__device__ double iso_predict(const float* __restrict__ tx, const float* __restrict__ ty, int n, double x) {
  if (n <= 0) return 0.0;
  if (n == 1) return (double)ty[0];
  const double lo = tx[0], hi = tx[n - 1];
  x = fmin(fmax(x, lo), hi);
  int a = 0, b = n - 1;  // invariant: tx[a] <= x <= tx[b]
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    if ((double)tx[mid] <= x) a = mid;
    else b = mid;
  }
  const double x0 = tx[a], x1 = tx[b], y0 = ty[a], y1 = ty[b];
  if (x1 == x0) return y0;
  return (y1 - y0) / (x1 - x0) * (x - x0) + y0;
}

__device__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

__global__ void __launch_bounds__(128) autolabel_kernel(const AutoParams p) {
  const int b = blockIdx.x;
  const udal_autolabel_params& q = p.prm;
  __shared__ float first_sigma[4];
  __shared__ int all_below;
  if (threadIdx.x == 0) all_below = 1;
  const bool calibrated = q.calib_method_box != UDAL_CALIB_NONE;
  __syncthreads();
  // pass 1: entropy and calibrated aleatoric std of every row
  for (int i = threadIdx.x; i < p.M; i += blockDim.x) {
    const size_t row = (size_t)b * p.M + i;
    // ---- entropy of the (temperature scaled) softmax, fp32 like NumPy on float32 logits ----
    const float* lg = p.logits + row * p.C;
    float mx = -INFINITY;
    for (int c = 0; c < p.C; ++c) mx = fmaxf(mx, lg[c] / q.class_temp);
    float sum = 0.f;
    for (int c = 0; c < p.C; ++c) sum += expf(lg[c] / q.class_temp - mx);
    float ent = 0.f;
    for (int c = 0; c < p.C; ++c) {
      const float pr = expf(lg[c] / q.class_temp - mx) / sum;
      ent += pr * log2f(fmaxf(pr, 1e-7f));
    }
    p.entropy[row] = -ent;
    // ---- calibration of the aleatoric std ----
    const float* bx = p.boxes + row * p.box_stride;
    const int cls = (int)p.classes[row * p.class_stride];
    const float h = bx[2] - bx[0], w = bx[3] - bx[1];
    for (int j = 0; j < 4; ++j) {
      // np.nan_to_num is part of calibrate_boxuncert (utils_box.py:417): without a calibrator the raw std
      // (NaNs included) flows on
      const float raw = p.albox_col >= 0 ? bx[p.albox_col + j] : 0.f;
      const float u = calibrated ? nan_to_num(raw) : raw;
      float v = u;
      switch (q.calib_method_box) {
        case UDAL_CALIB_TS_ALL: v = u / q.temps[0]; break;
        case UDAL_CALIB_TS_PERCOO: v = u / q.temps[j]; break;
        case UDAL_CALIB_ISO_ALL:
        case UDAL_CALIB_ISO_PERCOO:
        case UDAL_CALIB_ISO_PERCLSCOO:
        case UDAL_CALIB_REL_ISO_PERCLSCOO: {
          int t = 0;
          if (q.calib_method_box == UDAL_CALIB_ISO_PERCOO) t = j;
          if (q.calib_method_box == UDAL_CALIB_ISO_PERCLSCOO || q.calib_method_box == UDAL_CALIB_REL_ISO_PERCLSCOO)
            t = (cls >= 1 && cls * 4 <= q.num_tables) ? (cls - 1) * 4 + j : -1;
          if (t < 0) {
            v = 0.f;  // classes without a calibrator keep the zeros of np.zeros_like (utils_box.py:456-466)
            break;
          }
          const int o = q.table_off[t], n = q.table_off[t + 1] - o;
          if (q.calib_method_box == UDAL_CALIB_REL_ISO_PERCLSCOO) {
            // np.divide(uncert, norm, where=norm != 0, dtype=np.float16), calibrate, multiply back
            const float norm = (j & 1) ? w : h;
            const __half r = norm != 0.f ? __float2half(__half2float(__float2half(u)) / __half2float(__float2half(norm)))
                                         : __float2half(0.f);
            v = (float)iso_predict(q.table_x + o, q.table_y + o, n, (double)__half2float(r)) * norm;
          } else {
            v = (float)iso_predict(q.table_x + o, q.table_y + o, n, (double)u);
          }
          break;
        }
        default: break;
      }
      p.calib_albox[row * 4 + j] = v;
      if (i == 0) first_sigma[j] = v;
    }
  }
  __syncthreads();
  // pass 2: relative std, weighted sum, decision
  for (int i = threadIdx.x; i < p.M; i += blockDim.x) {
    const size_t row = (size_t)b * p.M + i;
    const float* bx = p.boxes + row * p.box_stride;
    const float h = bx[2] - bx[0], w = bx[3] - bx[1];
    float acc = 0.f;
    for (int j = 0; j < 4; ++j) {
      // infer_model.py:688-691: after calibration the reference indexes the FIRST detection's std
      const float s = (q.strict_reference && calibrated) ? first_sigma[j] : p.calib_albox[row * 4 + j];
      const float r = s / ((j & 1) ? w : h);
      p.rel_albox[row * 4 + j] = r;
      acc += r;
    }
    // an uncertainty that is not selected (weight 0) is not part of the sum at all: no 0 * NaN
    float opt = 0.f;
    if (q.w_entropy != 0.f) opt += q.w_entropy * p.entropy[row];
    if (q.w_albox != 0.f) opt += q.w_albox * (acc / 4.f);
    p.opt_uncert[row] = opt;
    if (p.scores[row] > q.min_score && !(opt < q.threshold)) all_below = 0;  // benign race: only zeros are written
  }
  __syncthreads();
  if (threadIdx.x == 0) p.decision[b] = all_below;
}

}  // namespace

extern "C" int udal_autolabel(udal_ctx* ctx, const float* boxes, int box_stride, int albox_col, const float* scores,
                              const float* classes, int class_stride, const float* logits, int num_classes, int batch,
                              int max_out, const udal_autolabel_params* prm, float* entropy, float* calib_albox,
                              float* rel_albox, float* opt_uncert, int32_t* decision) {
  UDAL_REQUIRE(ctx && boxes && scores && classes && logits && prm && entropy && calib_albox && rel_albox && opt_uncert && decision,
               "NULL argument");
  UDAL_TRY(udal_join(ctx));
  UDAL_REQUIRE(box_stride >= 4 && (albox_col < 0 || albox_col + 4 <= box_stride), "bad box layout");
  UDAL_REQUIRE(prm->calib_method_box >= UDAL_CALIB_NONE && prm->calib_method_box <= UDAL_CALIB_REL_ISO_PERCLSCOO,
               "Unknown calibration method %d", prm->calib_method_box);
  if (prm->calib_method_box >= UDAL_CALIB_ISO_ALL) {
    const int need = prm->calib_method_box == UDAL_CALIB_ISO_ALL ? 1 : prm->calib_method_box == UDAL_CALIB_ISO_PERCOO ? 4 : 4;
    UDAL_REQUIRE(prm->table_x && prm->table_y && prm->table_off && prm->num_tables >= need,
                 "isotonic calibration needs %d table(s), got %d", need, prm->num_tables);
  }
  UDAL_REQUIRE(prm->class_temp != 0.f, "class temperature must be non-zero");
  if (batch == 0) return UDAL_OK;
  AutoParams p;
  p.boxes = boxes;
  p.scores = scores;
  p.classes = classes;
  p.logits = logits;
  p.box_stride = box_stride;
  p.albox_col = albox_col;
  p.class_stride = class_stride;
  p.C = num_classes;
  p.M = max_out;
  p.prm = *prm;
  p.entropy = entropy;
  p.calib_albox = calib_albox;
  p.rel_albox = rel_albox;
  p.opt_uncert = opt_uncert;
  p.decision = decision;
  autolabel_kernel<<<batch, 128, 0, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// CalibrateClass._perform_class_calib without the MC class uncertainty (reference src/utils_class.py:116-187):
//   ts_all / ts_percls   probab = stable_softmax(logits / T)            (T scalar, or one per class)
//   iso_all / iso_percls p = stable_softmax(logits); q = iso.predict(p) (one table, or one per class; float64, clipped
//                        piece-wise linear); probab = q / sum(q)
//   entropy = -sum(probab * nan_to_num(log2(max(probab, 1e-7))))
// One thread per detection row.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct ClassCalParams {
  const float* logits;   // [rows, C]
  long long rows;
  int C, method;
  const float* temps;    // [1] | [C]
  const float* tx;       // isotonic knots, tables back to back
  const float* ty;
  const int32_t* off;    // [ntables + 1]
  float* probab;         // [rows, C]
  float* entropy;        // [rows]
};

__global__ void class_calibrate_kernel(const ClassCalParams p) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  const float* x = p.logits + r * p.C;
  float* out = p.probab + r * p.C;
  const bool ts = p.method == UDAL_CLASSCAL_TS_ALL || p.method == UDAL_CLASSCAL_TS_PERCLS;
  const bool per = p.method == UDAL_CLASSCAL_TS_PERCLS || p.method == UDAL_CLASSCAL_ISO_PERCLS;
  // stable softmax of the (temperature-scaled) logits, fp32 like NumPy on float32 inputs
  float mx = -3.4028234663852886e38f;
  for (int c = 0; c < p.C; ++c) {
    const float v = ts ? __fdiv_rn(x[c], p.temps[per ? c : 0]) : x[c];
    mx = fmaxf(mx, v);
  }
  float sum = 0.f;
  for (int c = 0; c < p.C; ++c) {
    const float v = ts ? __fdiv_rn(x[c], p.temps[per ? c : 0]) : x[c];
    const float e = expf(v - mx);
    out[c] = e;
    sum += e;
  }
  double ent = 0.0;
  if (ts) {
    float entf = 0.f;
    for (int c = 0; c < p.C; ++c) {
      const float pr = __fdiv_rn(out[c], sum);
      out[c] = pr;
      entf += pr * log2f(fmaxf(pr, 1e-7f));
    }
    p.entropy[r] = -entf;
    return;
  }
  double qs = 0.0;
  for (int c = 0; c < p.C; ++c) {
    const int t = per ? c : 0;
    const double q = iso_predict(p.tx + p.off[t], p.ty + p.off[t], p.off[t + 1] - p.off[t], (double)__fdiv_rn(out[c], sum));
    qs += q;
  }
  for (int c = 0; c < p.C; ++c) {
    const int t = per ? c : 0;
    const double q = iso_predict(p.tx + p.off[t], p.ty + p.off[t], p.off[t + 1] - p.off[t], (double)__fdiv_rn(out[c], sum));
    const double pr = q / qs;
    out[c] = (float)pr;
    ent += pr * log2(fmax(pr, 1e-7));
  }
  p.entropy[r] = (float)(-ent);
}

}  // namespace

extern "C" int udal_calibrate_class(udal_ctx* ctx, const float* logits, long long rows, int C, int method, const float* temps,
                                    const float* tx, const float* ty, const int32_t* off, float* probab, float* entropy) {
  UDAL_REQUIRE(ctx && logits && probab && entropy, "NULL argument");
  UDAL_REQUIRE(rows >= 0 && C >= 1, "udal_calibrate_class: bad sizes");
  UDAL_REQUIRE(method >= UDAL_CLASSCAL_TS_ALL && method <= UDAL_CLASSCAL_ISO_PERCLS, "Unknown calibration method");
  const bool ts = method == UDAL_CLASSCAL_TS_ALL || method == UDAL_CLASSCAL_TS_PERCLS;
  UDAL_REQUIRE(ts ? temps != nullptr : (tx && ty && off), "udal_calibrate_class: calibrator tables missing");
  UDAL_TRY(udal_join(ctx));
  if (rows == 0) return UDAL_OK;
  ClassCalParams p = {logits, rows, C, method, temps, tx, ty, off, probab, entropy};
  class_calibrate_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(p);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
