// Per-label ranking metrics of a split on the GPU: AUROC, area under the precision-recall curve, recall at a
// false-discovery-rate cutoff and average precision -- what utils/evals.py:86-90 asks sklearn for through
// utils/metrics.py:238-253 (roc_auc_score), :168-183 (precision_recall_curve + auc), :148-165 (recall at the first
// point with 1 - precision <= 0.5) and :25-26 (average_precision_score), once per split per epoch on the
// [sum N, nclass] prediction matrix (runner.py:41,45,51).  SURVEY.md section 8(f) rank 4.
//
// All four are functions of the confusion counts (tp_j, fp_j) at the distinct score thresholds of one label:
//   1. one key per (window, label): (label << 33) | (descending-order image of the fp32 score << 1) | (1 - target);
//   2. one stable LSD radix sort of all keys (radix_sort.cuh; 33 + log2(labels) bits): label segments, scores descending;
//   3. one CTA per label walks its segment tile by tile: block prefix sum of the target bits (tp), tie-group ends
//      (score differs from the next one) are the thresholds sklearn's _binary_clf_curve keeps, a block max-scan carries
//      the previous threshold's (tp, fp) to the next one, and the trapezoid / step sums accumulate per thread;
//      integer sums are exact (uint64), fp64 sums are combined in a fixed order (deterministic).
// Pure integer / byte work bound by the sort's HBM passes.
#include <math_constants.h>

#include "common.cuh"
#include "radix_sort.cuh"

namespace cgcn {

constexpr int MT_THREADS = 1024;
constexpr int MT_ITEMS = 4;
constexpr int MT_TILE = MT_THREADS * MT_ITEMS;

// fp32 score -> 32-bit key whose ascending unsigned order is the DESCENDING order of the scores (-0.0 == +0.0)
__device__ __forceinline__ uint32_t desc_key(float s) {
  if (s == 0.0f) s = 0.0f;
  uint32_t b = __float_as_uint(s);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ~b;
}

__global__ void __launch_bounds__(256) metrics_keys_kernel(const float* __restrict__ preds, int64_t pred_ld,
                                                           const float* __restrict__ targets, int64_t tgt_ld,
                                                           const uint32_t* __restrict__ tbits, int wpr, int c0, int G, int64_t n,
                                                           uint64_t* __restrict__ keys) {
  const int64_t total = n * G;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = e / G;
    const int c = static_cast<int>(e - i * G);
    const int cl = c0 + c;
    const float s = __ldg(preds + i * pred_ld + cl);
    const bool t = (tbits != nullptr) ? ((__ldg(tbits + i * wpr + (cl >> 5)) >> (cl & 31)) & 1u) != 0
                                      : __ldg(targets + i * tgt_ld + cl) != 0.0f;
    keys[e] = (static_cast<uint64_t>(c) << 33) | (static_cast<uint64_t>(desc_key(s)) << 1) | (t ? 0ull : 1ull);
  }
}

// ---- block scans over MT_THREADS threads (one value per thread); sh: 32 entries per scan
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* sh, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = sh[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    sh[lane] = winc - w;                       // exclusive prefix of the warp totals
    if (lane == 31) sh[32] = winc;             // block total
  }
  __syncthreads();
  const uint32_t r = sh[warp] + inc - v;
  *total = sh[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ uint64_t block_excl_max(uint64_t v, uint64_t* sh, uint64_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o && t > inc) inc = t;
  }
  uint64_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = 0;
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint64_t w = sh[lane];
    uint64_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o && t > winc) winc = t;
    }
    uint64_t wex = __shfl_up_sync(0xffffffffu, winc, 1);
    if (lane == 0) wex = 0;
    sh[lane] = wex;
    if (lane == 31) sh[32] = winc;
  }
  __syncthreads();
  const uint64_t pre = sh[warp];
  *total = sh[32];
  __syncthreads();
  return pre > excl ? pre : excl;
}

struct LabelMetricsOut {
  double* auroc;   // NaN when the label has one class only (roc_auc_score raises; utils/metrics.py:245-246 skips it)
  double* aupr;
  double* fdr;     // recall at the lowest threshold with 1 - precision <= cutoff (0 if none)
  double* ap;
  double* npos;
};

__global__ void __launch_bounds__(MT_THREADS) metrics_label_kernel(const uint64_t* __restrict__ keys, int64_t n, int c0,
                                                                   double fdr_cutoff, LabelMetricsOut out) {
  __shared__ uint32_t sh_sum[33];
  __shared__ uint64_t sh_max[33];
  __shared__ unsigned long long red_u[32];
  __shared__ double red_a[32], red_b[32];
  __shared__ uint32_t red_f[32];
  const uint64_t* seg = keys + static_cast<int64_t>(blockIdx.x) * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t carry_tp = 0;                       // positives before this tile
  uint64_t carry_pt = 0;                       // previous threshold point packed (pos + 1) << 32 | tp ; 0 = the origin
  unsigned long long roc = 0;                  // sum (fp - fp') (tp + tp')      == 2 P N AUROC, exact
  double pr = 0.0, ap = 0.0;                   // sum (tp - tp') (p + p') ; sum (tp - tp') p
  uint32_t fdr_tp = 0;
  for (int64_t base = 0; base < n; base += MT_TILE) {
    const int64_t p0 = base + static_cast<int64_t>(tid) * MT_ITEMS;
    uint64_t k[MT_ITEMS + 1];
#pragma unroll
    for (int u = 0; u <= MT_ITEMS; ++u) k[u] = (p0 + u < n) ? __ldg(seg + p0 + u) : ~0ull;
    uint32_t t[MT_ITEMS], tsum = 0;
    bool isb[MT_ITEMS];
    uint64_t last_b = 0;
#pragma unroll
    for (int u = 0; u < MT_ITEMS; ++u) {
      const bool valid = p0 + u < n;
      t[u] = (valid && !(k[u] & 1ull)) ? 1u : 0u;
      tsum += t[u];
      isb[u] = valid && ((k[u] >> 1) != (k[u + 1] >> 1));      // the sentinel ends the last tie group of the segment
    }
    uint32_t tile_tp;
    const uint32_t excl = block_excl_sum(tsum, sh_sum, &tile_tp);
    uint32_t tp = carry_tp + excl;
    // this thread's last threshold point (if any), for the threads after it
    {
      uint32_t run = tp;
#pragma unroll
      for (int u = 0; u < MT_ITEMS; ++u) {
        run += t[u];
        if (isb[u]) last_b = (static_cast<uint64_t>(p0 + u + 1) << 32) | run;
      }
    }
    uint64_t tile_last;
    uint64_t prev = block_excl_max(last_b, sh_max, &tile_last);
    if (prev == 0) prev = carry_pt;
#pragma unroll
    for (int u = 0; u < MT_ITEMS; ++u) {
      tp += t[u];
      if (isb[u]) {
        const uint32_t tp0 = static_cast<uint32_t>(prev), cnt0 = static_cast<uint32_t>(prev >> 32);
        const uint32_t cnt = static_cast<uint32_t>(p0 + u + 1);
        const uint32_t fp = cnt - tp, fp0 = cnt0 - tp0;
        roc += static_cast<unsigned long long>(fp - fp0) * static_cast<unsigned long long>(tp + tp0);
        const double p = static_cast<double>(tp) / static_cast<double>(cnt);
        const double pp = cnt0 ? static_cast<double>(tp0) / static_cast<double>(cnt0) : 1.0;
        const double dtp = static_cast<double>(tp - tp0);
        pr += dtp * (p + pp);
        ap += dtp * p;
        if (1.0 - p <= fdr_cutoff && tp > fdr_tp) fdr_tp = tp;
        prev = (static_cast<uint64_t>(cnt) << 32) | tp;
      }
    }
    carry_tp += tile_tp;
    if (tile_last != 0) carry_pt = tile_last;
  }
  // fixed-order block reduction: lanes by shuffle tree, warps in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    roc += __shfl_down_sync(0xffffffffu, roc, o);
    pr += __shfl_down_sync(0xffffffffu, pr, o);
    ap += __shfl_down_sync(0xffffffffu, ap, o);
    const uint32_t f = __shfl_down_sync(0xffffffffu, fdr_tp, o);
    if (f > fdr_tp) fdr_tp = f;
  }
  if (lane == 0) {
    red_u[warp] = roc;
    red_a[warp] = pr;
    red_b[warp] = ap;
    red_f[warp] = fdr_tp;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long r = 0;
    double a = 0.0, b = 0.0;
    uint32_t f = 0;
    for (int w = 0; w < MT_THREADS / 32; ++w) {
      r += red_u[w];
      a += red_a[w];
      b += red_b[w];
      if (red_f[w] > f) f = red_f[w];
    }
    const double P = static_cast<double>(carry_tp), N = static_cast<double>(n) - P;
    const int c = c0 + blockIdx.x;
    out.npos[c] = P;
    out.auroc[c] = (P > 0.0 && N > 0.0) ? static_cast<double>(r) / (2.0 * P * N) : CUDART_NAN;
    // no positives: sklearn sets recall to 1 at every threshold (precision 0) and appends (precision 1, recall 0):
    // the one trapezoid left has area 1/2; average precision and the recall at the cutoff are 0
    out.aupr[c] = P > 0.0 ? a / (2.0 * P) : 0.5;
    out.ap[c] = P > 0.0 ? b / P : 0.0;
    out.fdr[c] = P > 0.0 ? static_cast<double>(f) / P : 0.0;
  }
}

static int label_bits(int G) {
  int b = 0;
  while ((1 << b) < G) ++b;
  return b;
}

// labels per sort: bounded so that one chunk's keys stay below 2^27 elements (1 GiB of keys per buffer)
static int labels_per_chunk(int64_t n, int nclass) {
  int64_t g = (1ll << 27) / (n < 1 ? 1 : n);
  if (g < 1) g = 1;
  if (g > nclass) g = nclass;
  if (g > 128) g = 128;
  return static_cast<int>(g);
}

}  // namespace cgcn

using namespace cgcn;

extern "C" size_t cgcn_label_metrics_workspace_bytes(int64_t n, int32_t nclass) {
  if (n < 1 || nclass < 1) return 0;
  const int G = labels_per_chunk(n, nclass);
  const int64_t items = n * G;
  return 2 * align_up(static_cast<size_t>(items) * sizeof(uint64_t), 256) + rsort::sort_temp_bytes(items) + 1024;
}

extern "C" int cgcn_label_metrics(const float* preds, int64_t pred_ld, const float* targets, int64_t target_ld,
                                  const uint32_t* target_bits, int64_t n, int32_t nclass, double fdr_cutoff, double* out,
                                  void* workspace, size_t workspace_bytes, cgcn_stream_t stream) {
  CGCN_REQUIRE(preds && out && (targets != nullptr) != (target_bits != nullptr),
               "cgcn_label_metrics: needs preds, out and exactly one of targets / target_bits");
  CGCN_REQUIRE(n >= 1 && n < (1ll << 31) && nclass >= 1 && pred_ld >= nclass && (targets == nullptr || target_ld >= nclass),
               "cgcn_label_metrics: bad shape n=%lld nclass=%d", static_cast<long long>(n), nclass);
  const size_t need = cgcn_label_metrics_workspace_bytes(n, nclass);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("cgcn_label_metrics: workspace %zu < %zu bytes", workspace_bytes, need);
    return CGCN_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int G = labels_per_chunk(n, nclass);
  Arena a(workspace, workspace_bytes);
  uint64_t* ka = a.take<uint64_t>(static_cast<size_t>(n) * G);
  uint64_t* kb = a.take<uint64_t>(static_cast<size_t>(n) * G);
  void* temp = a.take<char>(rsort::sort_temp_bytes(n * G));
  CGCN_REQUIRE(a.ok(), "cgcn_label_metrics: workspace layout overflow");
  LabelMetricsOut o{out, out + nclass, out + 2 * nclass, out + 3 * nclass, out + 4 * nclass};
  const int wpr = (nclass + 31) / 32;
  for (int c0 = 0; c0 < nclass; c0 += G) {
    const int g = (nclass - c0) < G ? (nclass - c0) : G;
    const int64_t items = n * g;
    int grid = static_cast<int>((items + 255) / 256);
    if (grid > sm_count() * 16) grid = sm_count() * 16;
    metrics_keys_kernel<<<grid, 256, 0, st>>>(preds, pred_ld, targets, target_ld, target_bits, wpr, c0, g, n, ka);
    CGCN_TRY(check_launch("metrics_keys_kernel"));
    bool in_b = false;
    CGCN_TRY(rsort::radix_sort<uint64_t>(ka, kb, nullptr, nullptr, items, 0, 33 + label_bits(g), temp, st, &in_b));
    metrics_label_kernel<<<g, MT_THREADS, 0, st>>>(in_b ? kb : ka, n, c0, fdr_cutoff, o);
    CGCN_TRY(check_launch("metrics_label_kernel"));
  }
  return CGCN_OK;
}
