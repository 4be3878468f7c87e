// K3: exact top-k of a [B,M] fp32 array in canonical order (value descending, index ascending).
//
// Replaces tf.math.top_k as used at reference src/postprocess.py:100-105 (k = max_nms_inputs
// over the flattened [N*C] mean logits) and is the candidate pre-filter of the global soft-NMS.
//
// Every element gets a distinct 64-bit composite key  (orderable(value) << 32) | ~index  so the
// k largest composites ARE the canonical top-k, ties included.  Three launches:
//   1. hist    : 12-bit radix histogram of the top key bits (one read of the input)
//   2. filter  : bins above the threshold bin -> selected list, threshold bin -> candidate list
//                (second read of the input; typically L2 resident)
//   3. finalize: one CTA per row narrows the candidate list with 11-bit radix passes until
//                selected + candidates fit the shared-memory sorter, then bitonic-sorts them.
// If a threshold bin overflows the candidate buffer (degenerate inputs, e.g. all values equal)
// finalize falls back to radix passes over the full input row - slower, still exact.
#include "udal_common.cuh"

namespace {

constexpr int kBins1 = 4096;          // first pass: 12 bits
constexpr int kHistThreads = 256;
constexpr int kChunk = 8192;          // elements per CTA in hist / filter
constexpr int kFinalThreads = 1024;
constexpr int kSortCap = 8192;        // entries the shared-memory sorter holds (64 KB)

struct TopkMeta {  // per row
  unsigned int n_sel;
  unsigned int n_cand;
};

__device__ __forceinline__ unsigned long long composite(float v, unsigned int idx) {
  // v + 0.0f maps -0.0 to +0.0 so both zeros compare equal (as in NumPy / TF) and tie on index
  return ((unsigned long long)udal_float_key(v + 0.0f) << 32) | (unsigned long long)(~idx);
}

__global__ void __launch_bounds__(kHistThreads) topk_hist_kernel(const float* __restrict__ values,
                                                                 int64_t m, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[kBins1];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kBins1; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  const float* row = values + (size_t)b * m;
  const int64_t start = (int64_t)blockIdx.x * kChunk;
  const int64_t end = min(start + (int64_t)kChunk, m);
  for (int64_t i = start + threadIdx.x; i < end; i += kHistThreads)
    atomicAdd(&sh[udal_float_key(__ldg(row + i) + 0.0f) >> 20], 1u);
  __syncthreads();
  unsigned int* gh = hist + (size_t)b * kBins1;
  for (int i = threadIdx.x; i < kBins1; i += kHistThreads)
    if (sh[i]) atomicAdd(&gh[i], sh[i]);
}

// threshold bin of a 4096-bin histogram: largest tb with count(bins > tb) < k.
// Block-wide (any block size >= 256; the first 256 threads work, everyone synchronises).
// Writes tb to *tb_out and count(bins > tb) to *above_out (both in shared memory).
__device__ void find_threshold_bin(const unsigned int* __restrict__ gh, unsigned int k, unsigned int* sh_part,
                                   int* tb_out, unsigned int* above_out) {
  constexpr int kWorkers = 256;
  constexpr int per = kBins1 / kWorkers;
  const int t = threadIdx.x;
  const bool active = t < kWorkers;
  // worker t owns bins [per*(255-t), per*(255-t)+per): worker 0 holds the highest bins
  const int base = per * (kWorkers - 1 - (active ? t : 0));
  unsigned int loc[per];
  unsigned int s = 0;
  if (active) {
#pragma unroll
    for (int i = 0; i < per; ++i) {
      loc[i] = gh[base + i];
      s += loc[i];
    }
    sh_part[t] = s;
  }
  __syncthreads();
  for (int off = 1; off < kWorkers; off <<= 1) {
    unsigned int v = (active && t >= off) ? sh_part[t - off] : 0;
    __syncthreads();
    if (active) sh_part[t] += v;
    __syncthreads();
  }
  if (active) {
    const unsigned int incl = sh_part[t];
    const unsigned int excl = incl - s;
    if (excl < k && incl >= k) {
      unsigned int above = excl;
      for (int i = per - 1; i >= 0; --i) {
        if (above + loc[i] >= k) {
          *tb_out = base + i;
          *above_out = above;
          break;
        }
        above += loc[i];
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kHistThreads) topk_filter_kernel(
    const float* __restrict__ values, int64_t m, unsigned int k, const unsigned int* __restrict__ hist,
    unsigned long long* __restrict__ sel, unsigned long long* __restrict__ cand, unsigned int cand_cap,
    TopkMeta* __restrict__ meta) {
  __shared__ unsigned int sh_part[kHistThreads];
  __shared__ int sh_tb;
  __shared__ unsigned int sh_above;
  const int b = blockIdx.y;
  find_threshold_bin(hist + (size_t)b * kBins1, k, sh_part, &sh_tb, &sh_above);
  const unsigned int tb = (unsigned int)sh_tb;
  const float* row = values + (size_t)b * m;
  unsigned long long* sel_row = sel + (size_t)b * k;
  unsigned long long* cand_row = cand + (size_t)b * 2 * cand_cap;
  const int64_t start = (int64_t)blockIdx.x * kChunk;
  const int64_t end = min(start + (int64_t)kChunk, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Two passes over the CTA's chunk (the second one hits L1 / L2): count, then ONE atomic per CTA and list, then write at
  // offsets that follow from the counts.  With one warp-aggregated atomic per 32 elements the ~3.5 k atomics per row on the
  // same two counters serialised in L2 (ncu: 85 % of the samples waited for them; 265 us for a second read of 126 MB).
  __shared__ unsigned int sh_cnt[2][kHistThreads / 32];
  __shared__ unsigned int sh_base[2][kHistThreads / 32];
  unsigned int n_s = 0, n_c = 0;
  for (int64_t i0 = start; i0 < end; i0 += kHistThreads) {
    const int64_t i = i0 + threadIdx.x;
    bool is_sel = false, is_cand = false;
    if (i < end) {
      const unsigned int bin = (unsigned int)(composite(__ldg(row + i), (unsigned int)i) >> 52);
      is_sel = bin > tb;
      is_cand = bin == tb;
    }
    n_s += __popc(__ballot_sync(0xffffffffu, is_sel));
    n_c += __popc(__ballot_sync(0xffffffffu, is_cand));
  }
  if (lane == 0) {
    sh_cnt[0][warp] = n_s;
    sh_cnt[1][warp] = n_c;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned int tot = 0;
    for (int w = 0; w < kHistThreads / 32; ++w) tot += sh_cnt[threadIdx.x][w];
    unsigned int base = 0;
    if (tot) base = atomicAdd(threadIdx.x == 0 ? &meta[b].n_sel : &meta[b].n_cand, tot);
    for (int w = 0; w < kHistThreads / 32; ++w) {
      sh_base[threadIdx.x][w] = base;
      base += sh_cnt[threadIdx.x][w];
    }
  }
  __syncthreads();
  unsigned int run_s = sh_base[0][warp], run_c = sh_base[1][warp];
  if (n_s == 0 && n_c == 0) return;   // (warp-uniform; no barrier follows)
  const unsigned int lt = (1u << lane) - 1;
  for (int64_t i0 = start; i0 < end; i0 += kHistThreads) {
    const int64_t i = i0 + threadIdx.x;
    bool is_sel = false, is_cand = false;
    unsigned long long comp = 0;
    if (i < end) {
      comp = composite(__ldg(row + i), (unsigned int)i);
      const unsigned int bin = (unsigned int)(comp >> 52);
      is_sel = bin > tb;
      is_cand = bin == tb;
    }
    const unsigned int ms = __ballot_sync(0xffffffffu, is_sel);
    const unsigned int mc = __ballot_sync(0xffffffffu, is_cand);
    if (is_sel) sel_row[run_s + __popc(ms & lt)] = comp;
    if (is_cand) {
      const unsigned int pos = run_c + __popc(mc & lt);
      if (pos < cand_cap) cand_row[pos] = comp;
    }
    run_s += __popc(ms);
    run_c += __popc(mc);
  }
}

// block-wide: find digit d* of an `nb`-bin histogram (shared memory, nb a power of two >= 32)
// such that count(d > d*) < need <= count(d >= d*).  Warp 0 scans, everyone synchronises.
__device__ void pick_digit(const unsigned int* sh_hist, int nb, unsigned int need, int* d_out,
                           unsigned int* above_out) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int per = nb >> 5;
    const int base = nb - (lane + 1) * per;  // lane 0 owns the highest digits
    unsigned int s = 0;
    for (int i = 0; i < per; ++i) s += sh_hist[base + i];
    unsigned int incl = s;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const unsigned int excl = incl - s;
    if (excl < need && incl >= need) {
      unsigned int above = excl;
      for (int i = per - 1; i >= 0; --i) {
        const unsigned int c = sh_hist[base + i];
        if (above + c >= need) {
          *d_out = base + i;
          *above_out = above;
          break;
        }
        above += c;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kFinalThreads) topk_finalize_kernel(
    const float* __restrict__ values, int64_t m, unsigned int k, const unsigned int* __restrict__ hist,
    unsigned long long* __restrict__ sel, unsigned long long* __restrict__ cand, unsigned int cand_cap,
    TopkMeta* __restrict__ meta, int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  extern __shared__ unsigned long long sort_buf[];  // kSortCap entries
  __shared__ unsigned int sh_hist[2048];
  __shared__ unsigned int sh_cnt[2];
  __shared__ int sh_d;
  __shared__ unsigned int sh_above;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* row = values + (size_t)b * m;
  unsigned long long* sel_row = sel + (size_t)b * k;
  unsigned long long* cbuf[2] = {cand + (size_t)b * 2 * cand_cap, cand + (size_t)b * 2 * cand_cap + cand_cap};
  unsigned int nsel = meta[b].n_sel;
  unsigned int ncand = meta[b].n_cand;
  unsigned int need = k - nsel;
  // decided prefix of the composite key: bits [63:lowbit] equal `prefix`; starts as the
  // threshold bin of the first pass
  find_threshold_bin(hist + (size_t)b * kBins1, k, sh_hist, &sh_d, &sh_above);
  unsigned long long prefix = (unsigned long long)sh_d;  // value of bits [63:lowbit]
  int lowbit = 52;
  bool from_input = ncand > cand_cap;  // overflow: candidates were not all materialised
  int cur = 0;
  __syncthreads();

  while (nsel + ncand > (unsigned int)kSortCap || from_input) {
    if (lowbit == 0) break;  // cannot happen: composites are distinct
    const int bits = lowbit >= 11 ? 11 : lowbit;
    const int shift = lowbit - bits;
    const int nb = 1 << bits;
    for (int i = tid; i < nb; i += kFinalThreads) sh_hist[i] = 0;
    if (tid < 2) sh_cnt[tid] = 0;
    __syncthreads();
    if (from_input) {
      for (int64_t i = tid; i < m; i += kFinalThreads) {
        const unsigned long long c = composite(__ldg(row + i), (unsigned int)i);
        if ((c >> lowbit) == prefix) atomicAdd(&sh_hist[(unsigned int)(c >> shift) & (nb - 1)], 1u);
      }
    } else {
      for (unsigned int i = tid; i < ncand; i += kFinalThreads)
        atomicAdd(&sh_hist[(unsigned int)(cbuf[cur][i] >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    pick_digit(sh_hist, nb, need, &sh_d, &sh_above);
    const unsigned int dstar = (unsigned int)sh_d;
    const unsigned int n_eq = sh_hist[dstar];
    const bool materialise = !from_input || n_eq <= cand_cap;
    // scatter: digit > d* -> selected; digit == d* -> next candidate buffer
    if (from_input) {
      for (int64_t i = tid; i < m; i += kFinalThreads) {
        const unsigned long long c = composite(__ldg(row + i), (unsigned int)i);
        if ((c >> lowbit) != prefix) continue;
        const unsigned int d = (unsigned int)(c >> shift) & (nb - 1);
        if (d > dstar) sel_row[nsel + atomicAdd(&sh_cnt[0], 1u)] = c;
        else if (d == dstar && materialise) cbuf[cur ^ 1][atomicAdd(&sh_cnt[1], 1u)] = c;
      }
    } else {
      for (unsigned int i = tid; i < ncand; i += kFinalThreads) {
        const unsigned long long c = cbuf[cur][i];
        const unsigned int d = (unsigned int)(c >> shift) & (nb - 1);
        if (d > dstar) sel_row[nsel + atomicAdd(&sh_cnt[0], 1u)] = c;
        else if (d == dstar) cbuf[cur ^ 1][atomicAdd(&sh_cnt[1], 1u)] = c;
      }
    }
    __syncthreads();
    nsel += sh_above;
    need -= sh_above;
    ncand = n_eq;
    prefix = (prefix << bits) | dstar;
    lowbit = shift;
    if (materialise) {
      from_input = false;
      cur ^= 1;
    }
    __syncthreads();
  }

  // ---- sort selected + remaining candidates (descending composite) ------------------------
  const unsigned int total = nsel + ncand;
  unsigned int p2 = 1;
  while (p2 < total) p2 <<= 1;
  for (unsigned int i = tid; i < p2; i += kFinalThreads) {
    unsigned long long v = 0ull;  // sorts last (a real composite is never 0: ~idx != 0 for idx < 2^32-1)
    if (i < nsel) v = sel_row[i];
    else if (i < total) v = cbuf[cur][i - nsel];
    sort_buf[i] = v;
  }
  __syncthreads();
  for (unsigned int size = 2; size <= p2; size <<= 1) {
    for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned int i = tid; i < (p2 >> 1); i += kFinalThreads) {
        const unsigned int lo = 2 * i - (i & (stride - 1));
        const unsigned int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sort_buf[lo], c = sort_buf[hi];
        if ((a < c) == desc) {
          sort_buf[lo] = c;
          sort_buf[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (unsigned int i = tid; i < k; i += kFinalThreads) {
    const unsigned long long c = sort_buf[i];
    idx_out[(size_t)b * k + i] = (int32_t)(~(unsigned int)(c & 0xffffffffull));
    if (val_out) val_out[(size_t)b * k + i] = udal_key_float((unsigned int)(c >> 32));
  }
}

}  // namespace

int udal_topk_cand_cap_override = 0;  // tests shrink the candidate buffer to exercise the fallback

int udal_launch_topk(udal_ctx* ctx, const float* values, int batch, int64_t m, int k, int32_t* idx_out,
                     float* val_out) {
  UDAL_REQUIRE(values && idx_out, "udal_topk: NULL pointer");
  UDAL_REQUIRE(batch > 0 && m > 0, "udal_topk: empty input");
  UDAL_REQUIRE(k > 0 && k <= kSortCap, "udal_topk: k=%d outside [1,%d]", k, kSortCap);
  UDAL_REQUIRE((int64_t)k <= m, "udal_topk: k=%d exceeds the row length %lld", k, (long long)m);
  UDAL_REQUIRE(m < (int64_t)0xffffffffll, "udal_topk: row too long");
  unsigned int cap = 1u << 17;
  if (udal_topk_cand_cap_override > 0) cap = (unsigned int)udal_topk_cand_cap_override;
  if ((int64_t)cap > m) cap = (unsigned int)m;
  unsigned int* hist;
  unsigned long long* bufs;
  TopkMeta* meta;
  UDAL_TRY(udal_scratch_get(ctx, SCR_TOPK_HIST, (size_t)batch * kBins1 * sizeof(unsigned int) + (size_t)batch * sizeof(TopkMeta), (void**)&hist));
  meta = (TopkMeta*)(hist + (size_t)batch * kBins1);
  UDAL_TRY(udal_scratch_get(ctx, SCR_TOPK_CAND, ((size_t)batch * k + (size_t)batch * 2 * cap) * sizeof(unsigned long long), (void**)&bufs));
  unsigned long long* sel = bufs;
  unsigned long long* cand = bufs + (size_t)batch * k;
  UDAL_CUDA(cudaMemsetAsync(hist, 0, (size_t)batch * kBins1 * sizeof(unsigned int) + (size_t)batch * sizeof(TopkMeta), ctx->stream));
  const int chunks = (int)((m + kChunk - 1) / kChunk);
  dim3 grid(chunks, batch);
  topk_hist_kernel<<<grid, kHistThreads, 0, ctx->stream>>>(values, m, hist);
  UDAL_CHECK_LAUNCH(ctx);
  topk_filter_kernel<<<grid, kHistThreads, 0, ctx->stream>>>(values, m, (unsigned int)k, hist, sel, cand, cap, meta);
  UDAL_CHECK_LAUNCH(ctx);
  UDAL_CUDA(cudaFuncSetAttribute(topk_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSortCap * (int)sizeof(unsigned long long)));
  topk_finalize_kernel<<<batch, kFinalThreads, kSortCap * sizeof(unsigned long long), ctx->stream>>>(
      values, m, (unsigned int)k, hist, sel, cand, cap, meta, idx_out, val_out);
  UDAL_CHECK_LAUNCH(ctx);
  return UDAL_OK;
}
