// Device-wide primitives of the adjacency build, hand-written: an inclusive prefix sum and a stable
// LSD radix sort (8-bit digits) over 32/64-bit keys with an optional 32-bit payload.
//
// All of it is integer work bound by HBM traffic: each radix pass reads the keys twice (histogram, scatter) and
// writes them once.  Stability is what the contract needs (dict insertion order / "sorted(...) is stable",
// data/7create_graph_new.py:86,94): inside a CTA tile, items are ranked in index order with warp-level
// match_any + popc multi-split and per-warp digit counters; across CTAs by a digit-major scan of the per-CTA
// histograms.
#pragma once

#include "common.cuh"

namespace cgcn {
namespace rsort {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                         // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 per CTA

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;       // 4096 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

// ------------------------------------------------------------------------------------ prefix sum
__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_sums /*[8]*/, unsigned* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned incl = warp_inclusive_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const unsigned s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + incl - v;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T* __restrict__ in, int64_t n, unsigned* __restrict__ block_sums) {
  __shared__ unsigned ws[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += static_cast<unsigned>(in[base + i]);
  unsigned total;
  block_exclusive_scan(s, ws, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the block sums in place (sequential over chunks of 256 with a carry)
static __global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(unsigned* __restrict__ block_sums, int64_t nb) {
  __shared__ unsigned ws[8];
  unsigned carry = 0;
  for (int64_t c = 0; c < nb; c += SCAN_THREADS) {
    const int64_t i = c + threadIdx.x;
    const unsigned v = i < nb ? block_sums[i] : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan(v, ws, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
}

template <typename T, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const T* __restrict__ in, int64_t n, const unsigned* __restrict__ block_offsets,
                                                                  T* __restrict__ out) {
  __shared__ unsigned ws[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS];
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? static_cast<unsigned>(in[base + i]) : 0u;
    s += v[i];
  }
  unsigned total;
  unsigned run = block_exclusive_scan(s, ws, &total) + block_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const unsigned incl = run + v[i];
    if (base + i < n) out[base + i] = static_cast<T>(INCLUSIVE ? incl : run);
    run = incl;
  }
}

inline size_t scan_temp_bytes(int64_t n) {
  const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  return align_up(static_cast<size_t>(nb < 1 ? 1 : nb) * sizeof(unsigned), 256);
}

// out[i] = sum_{j<=i} in[j] (INCLUSIVE) or sum_{j<i} in[j]; totals must fit 32 bits.  in may alias out.
template <typename T, bool INCLUSIVE>
int prefix_sum(const T* in, T* out, int64_t n, void* temp, cudaStream_t stream) {
  if (n <= 0) return CGCN_OK;
  const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  unsigned* sums = static_cast<unsigned*>(temp);
  scan_reduce_kernel<T><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, stream>>>(in, n, sums);
  CGCN_TRY(check_launch("scan_reduce_kernel"));
  scan_spine_kernel<<<1, SCAN_THREADS, 0, stream>>>(sums, nb);
  CGCN_TRY(check_launch("scan_spine_kernel"));
  scan_apply_kernel<T, INCLUSIVE><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, stream>>>(in, n, sums, out);
  return check_launch("scan_apply_kernel");
}

// ------------------------------------------------------------------------------------ radix sort
template <typename KeyT>
__device__ __forceinline__ unsigned digit_of(KeyT k, int shift) {
  return static_cast<unsigned>((k >> shift) & static_cast<KeyT>(0xFF));
}

// per-CTA digit histogram, written digit-major: hist[digit * nblocks + block]
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift, unsigned nblocks,
                                                                unsigned* __restrict__ hist) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[digit_of(keys[i], shift)], 1u);
  }
  __syncthreads();
  hist[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: items of the tile in index order (round-major, then thread) keep their relative order per digit
template <typename KeyT, bool HAS_VALUES>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                                                                   const unsigned* __restrict__ vals_in, unsigned* __restrict__ vals_out,
                                                                   int64_t n, int shift, unsigned nblocks,
                                                                   const unsigned* __restrict__ offsets /* scanned hist */) {
  __shared__ unsigned running[256];                 // next output slot per digit for this CTA
  __shared__ unsigned warp_cnt[RS_WARPS][256];      // per round: items of each digit held by each warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = offsets[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; ++r) {
    if (base + static_cast<int64_t>(r) * RS_THREADS >= n) break;            // uniform
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = base + static_cast<int64_t>(r) * RS_THREADS + threadIdx.x;
    const bool valid = i < n;
    KeyT k = 0;
    unsigned d = 0xFFFFFFFFu;                        // invalid lanes form their own match group
    if (valid) {
      k = keys_in[i];
      d = digit_of(k, shift);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned before = __popc(peers & ((1u << lane) - 1u));
    if (valid && before == 0) warp_cnt[warp][d] = __popc(peers);            // one writer per (warp, digit)
    __syncthreads();
    unsigned pos = 0;
    if (valid) {
      unsigned prior = 0;
      for (int w = 0; w < warp; ++w) prior += warp_cnt[w][d];
      pos = running[d] + prior + before;
    }
    __syncthreads();
    {                                                // thread t advances digit t by this round's total
      unsigned tot = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) tot += warp_cnt[w][threadIdx.x];
      running[threadIdx.x] += tot;
    }
    if (valid) {
      keys_out[pos] = k;
      if (HAS_VALUES) vals_out[pos] = vals_in[i];
    }
    __syncthreads();
  }
}

inline size_t sort_temp_bytes(int64_t n) {
  const int64_t nb = (n + RS_TILE - 1) / RS_TILE;
  const size_t hist = align_up(static_cast<size_t>(nb < 1 ? 1 : nb) * 256 * sizeof(unsigned), 256);
  return hist + scan_temp_bytes(static_cast<int64_t>(nb < 1 ? 1 : nb) * 256) + 256;
}

// Stable ascending sort on key bits [begin_bit, end_bit).  Ping-pongs between (a) and (b); *result_in_b tells
// where the sorted data ended up.  vals may be NULL (keys only).
template <typename KeyT>
int radix_sort(KeyT* keys_a, KeyT* keys_b, unsigned* vals_a, unsigned* vals_b, int64_t n, int begin_bit, int end_bit,
               void* temp, cudaStream_t stream, bool* result_in_b) {
  *result_in_b = false;
  if (n <= 1) return CGCN_OK;
  const unsigned nb = static_cast<unsigned>((n + RS_TILE - 1) / RS_TILE);
  unsigned* hist = static_cast<unsigned*>(temp);
  void* scan_tmp = static_cast<char*>(temp) + align_up(static_cast<size_t>(nb) * 256 * sizeof(unsigned), 256);
  bool in_b = false;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    KeyT* kin = in_b ? keys_b : keys_a;
    KeyT* kout = in_b ? keys_a : keys_b;
    unsigned* vin = in_b ? vals_b : vals_a;
    unsigned* vout = in_b ? vals_a : vals_b;
    radix_hist_kernel<KeyT><<<nb, RS_THREADS, 0, stream>>>(kin, n, shift, nb, hist);
    CGCN_TRY(check_launch("radix_hist_kernel"));
    CGCN_TRY((prefix_sum<unsigned, false>(hist, hist, static_cast<int64_t>(nb) * 256, scan_tmp, stream)));
    if (vals_a != nullptr)
      radix_scatter_kernel<KeyT, true><<<nb, RS_THREADS, 0, stream>>>(kin, kout, vin, vout, n, shift, nb, hist);
    else
      radix_scatter_kernel<KeyT, false><<<nb, RS_THREADS, 0, stream>>>(kin, kout, nullptr, nullptr, n, shift, nb, hist);
    CGCN_TRY(check_launch("radix_scatter_kernel"));
    in_b = !in_b;
  }
  *result_in_b = in_b;
  return CGCN_OK;
}

}  // namespace rsort
}  // namespace cgcn
