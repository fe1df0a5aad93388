// Hi-C contact list -> window adjacency on the GPU, bit-exact with the reference's
// data/7create_graph_new.py:67-120 (see include/chromegcn.h for the contract), plus the
// structural half of process_graph('hic') (utils/util_methods.py:152-165) and the COO -> CSR
// pattern conversion for callers that still hold the reference's torch sparse tensor.
//
// All of it is integer / fp64 compare work bound by HBM traffic of the sort passes.  The
// device-wide primitives (stable LSD radix sort, prefix sum) are hand-written too (radix_sort.cuh);
// the contract-specific stages are the kernels below:
//   filter      binary-search both bins in the sorted window starts, fp64 normalise,
//               accept flag                                                    (:78-86)
//   dedup       stable sort by (i,j) key with the accepted-order index as payload; per key
//               group: position = first index, value = value at the last index (dict semantics, :86)
//   rank        stable sort by position, then stable sort by the order-preserving bit pattern
//               of -value  => "sorted(items, key=value, reverse=True)" with ties in insertion
//               order (:94); first K
//   symmetrise  (i,j),(j,i) keys, sort, unique, rowptr by binary search          (:108-120)
#include <math_constants.h>

#include "common.cuh"
#include "radix_sort.cuh"

namespace cgcn {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 KEY_NONE = ~0ull;

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* __restrict__ a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int64_t lower_bound_u64(const u64* __restrict__ a, int64_t n, u64 v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// err bits: 1 = NaN value among accepted rows, 2 = bin beyond the norm vector
__global__ void adj_filter_kernel(const int64_t* __restrict__ bin1, const int64_t* __restrict__ bin2,
                                  const double* __restrict__ val, int64_t m, const int64_t* __restrict__ starts, int64_t n,
                                  const double* __restrict__ norm, int64_t norm_len, int64_t res_bp, int use_norm,
                                  u64* __restrict__ key, double* __restrict__ v_out, int* __restrict__ flag,
                                  int* __restrict__ err) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < m;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b1 = __ldg(bin1 + r), b2 = __ldg(bin2 + r);
    int ok = 0;
    u64 k = KEY_NONE;
    double v = 0.0;
    if (b1 != b2) {
      const int64_t i = lower_bound_i64(starts, n, b1);
      const int64_t j = lower_bound_i64(starts, n, b2);
      if (i < n && j < n && __ldg(starts + i) == b1 && __ldg(starts + j) == b2) {
        ok = 1;
        k = (static_cast<u64>(i) << 32) | static_cast<u64>(j);
        v = __ldg(val + r);
        if (use_norm) {
          const int64_t q1 = b1 / res_bp, q2 = b2 / res_bp;
          if (q1 < 0 || q2 < 0 || q1 >= norm_len || q2 >= norm_len) {
            atomicOr(err, 2);
          } else {
            double n1 = __ldg(norm + q1), n2 = __ldg(norm + q2);
            if (n1 != n1 || n1 == 0.0) n1 = CUDART_INF;
            if (n2 != n2 || n2 == 0.0) n2 = CUDART_INF;
            v = __ddiv_rn(v, __dmul_rn(n1, n2));
          }
        }
      }
    }
    key[r] = k;
    v_out[r] = v;
    flag[r] = ok;
  }
}

// stable compaction of accepted rows; limit < 0: no limit, else keep only the first `limit`
__global__ void adj_compact_kernel(const u64* __restrict__ key, const double* __restrict__ v, const int* __restrict__ flag,
                                   const int* __restrict__ incl, int64_t m, int64_t limit, u64* __restrict__ key_c,
                                   double* __restrict__ v_c, u32* __restrict__ idx_c, int* __restrict__ err) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < m;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (!flag[r]) continue;
    const int64_t p = incl[r] - 1;
    if (limit >= 0 && p >= limit) continue;
    const double x = v[r];
    if (x != x) atomicOr(err, 1);             // only rows the reference would rank can poison the sort
    key_c[p] = key[r];
    v_c[p] = x;
    idx_c[p] = static_cast<u32>(p);
  }
}

__global__ void adj_heads_kernel(const u64* __restrict__ key_s, int64_t ma, int* __restrict__ head) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < ma;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x)
    head[t] = (t == 0 || key_s[t] != key_s[t - 1]) ? 1 : 0;
}

__device__ __forceinline__ u64 desc_value_key(double v) {
  if (v == 0.0) v = 0.0;                       // -0.0 compares equal to 0.0 in Python: one key for both
  u64 b = static_cast<u64>(__double_as_longlong(v));
  b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // ascending order-preserving map
  return ~b;                                   // descending
}

// one entry per distinct key: position = first accepted index, value = value at the last one
__global__ void adj_groups_kernel(const u64* __restrict__ key_s, const u32* __restrict__ idx_s, const int* __restrict__ head,
                                  const int* __restrict__ incl, const double* __restrict__ v_c, int64_t ma,
                                  u64* __restrict__ ukey, u32* __restrict__ upos, u64* __restrict__ uvkey,
                                  u32* __restrict__ iota) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < ma;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t g = incl[t] - 1;
    if (head[t]) {
      ukey[g] = key_s[t];
      upos[g] = idx_s[t];                      // stable sort: first of the group = smallest accepted index
      iota[g] = static_cast<u32>(g);
    }
    if (t == ma - 1 || key_s[t + 1] != key_s[t]) uvkey[g] = desc_value_key(v_c[idx_s[t]]);   // last assignment wins
  }
}

__global__ void adj_gather_vkey_kernel(const u64* __restrict__ uvkey, const u32* __restrict__ perm, int64_t u,
                                       u64* __restrict__ out) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < u;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[t] = uvkey[perm[t]];
}

__global__ void adj_symmetrise_kernel(const u64* __restrict__ ukey, const u32* __restrict__ perm, int64_t ksel,
                                      u64* __restrict__ sym) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < ksel;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const u64 k = ukey[perm[t]];
    sym[2 * t] = k;
    sym[2 * t + 1] = (k << 32) | (k >> 32);
  }
}

__global__ void adj_emit_cols_kernel(const u64* __restrict__ sym_s, const int* __restrict__ head, const int* __restrict__ incl,
                                     int64_t cnt, int64_t cap, int32_t* __restrict__ colidx, u64* __restrict__ uniq) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < cnt;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (!head[t]) continue;
    const int64_t q = incl[t] - 1;
    uniq[q] = sym_s[t];
    if (q < cap) colidx[q] = static_cast<int32_t>(sym_s[t] & 0xffffffffull);
  }
}

__global__ void adj_rowptr_kernel(const u64* __restrict__ uniq, int64_t nnz, int64_t n, int32_t* __restrict__ rowptr) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r <= n;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x)
    rowptr[r] = static_cast<int32_t>(lower_bound_u64(uniq, nnz, static_cast<u64>(r) << 32));
}

static int blocks_for(int64_t items) {
  int64_t b = (items + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

static int bits_for(int64_t count) {          // bits needed to represent values in [0, count)
  int b = 1;
  while (b < 63 && (1ll << b) < count) ++b;
  return b;
}

static size_t cub_temp_bytes(int64_t items) {      // scratch of the sort / scan primitives for `items` elements
  if (items < 1) items = 1;
  const size_t a = rsort::sort_temp_bytes(items), b = rsort::scan_temp_bytes(items);
  return (a > b ? a : b) + 256;
}

template <typename T>
static void swap_ptr(T*& a, T*& b) {
  T* t = a;
  a = b;
  b = t;
}

struct AdjWs {
  u64 *key, *key_c, *key_s, *ukey, *uvkey, *vk_a, *vk_b, *sym, *sym_s, *uniq;
  double *v, *v_c;
  int *flag, *incl, *head, *err_count;   // err_count: [0] err bits, [1..] scratch counts
  u32 *idx_c, *idx_s, *upos, *upos_s, *iota, *perm1, *perm2;
  void* cub_temp;
  size_t cub_bytes;
};

static size_t carve(AdjWs& w, void* base, size_t bytes, int64_t m, int64_t sym_cap) {
  Arena a(base, bytes);
  const size_t mm = static_cast<size_t>(m < 1 ? 1 : m);
  const size_t sc = static_cast<size_t>(sym_cap < 2 ? 2 : sym_cap);
  w.key = a.take<u64>(mm);    w.key_c = a.take<u64>(mm);  w.key_s = a.take<u64>(mm);
  w.ukey = a.take<u64>(mm);   w.uvkey = a.take<u64>(mm);  w.vk_a = a.take<u64>(mm);   w.vk_b = a.take<u64>(mm);
  w.v = a.take<double>(mm);   w.v_c = a.take<double>(mm);
  w.flag = a.take<int>(mm);   w.incl = a.take<int>(mm > sc ? mm : sc);  w.head = a.take<int>(mm > sc ? mm : sc);
  w.err_count = a.take<int>(16);
  w.idx_c = a.take<u32>(mm);  w.idx_s = a.take<u32>(mm);  w.upos = a.take<u32>(mm);  w.upos_s = a.take<u32>(mm);
  w.iota = a.take<u32>(mm);   w.perm1 = a.take<u32>(mm);  w.perm2 = a.take<u32>(mm);
  w.sym = a.take<u64>(sc);    w.sym_s = a.take<u64>(sc);  w.uniq = a.take<u64>(sc);
  w.cub_bytes = cub_temp_bytes(static_cast<int64_t>(mm > sc ? mm : sc));
  w.cub_temp = a.take<char>(w.cub_bytes);
  return a.off + 256;
}

static int64_t sym_capacity(int64_t m, int64_t k_pairs) {
  const int64_t pairs = (k_pairs > 0 && k_pairs < m) ? k_pairs : m;
  return 2 * pairs;
}

}  // namespace cgcn

using namespace cgcn;

extern "C" int cgcn_adj_build_workspace_bytes(int64_t m, int64_t n, int64_t k_pairs, size_t* bytes_host) {
  CGCN_REQUIRE(bytes_host != nullptr && m >= 0 && n >= 0 && k_pairs >= 0, "cgcn_adj_build_workspace_bytes: bad argument");
  CGCN_REQUIRE(m < 2147483647LL, "cgcn_adj_build: at most 2^31-1 contact rows per call");
  AdjWs w;
  *bytes_host = carve(w, nullptr, 0, m, sym_capacity(m, k_pairs));
  return CGCN_OK;
}

extern "C" int cgcn_adj_build(const int64_t* bin1, const int64_t* bin2, const double* val, int64_t m,
                              const int64_t* window_starts, int64_t n, const double* norm, int64_t norm_len,
                              int64_t res_bp, int64_t k_pairs, int32_t use_norm, int32_t* rowptr, int32_t* colidx,
                              int64_t colidx_cap, int64_t* nnz_host, void* workspace, size_t workspace_bytes,
                              cgcn_stream_t stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CGCN_REQUIRE(m >= 0 && n >= 0 && k_pairs >= 0 && res_bp > 0, "cgcn_adj_build: bad size argument");
  CGCN_REQUIRE(m < 2147483647LL && n < 2147483647LL, "cgcn_adj_build: sizes must fit int32");
  CGCN_REQUIRE(rowptr && nnz_host && (m == 0 || (bin1 && bin2 && val)) && (n == 0 || window_starts),
               "cgcn_adj_build: null argument");
  CGCN_REQUIRE(!use_norm || norm != nullptr, "cgcn_adj_build: use_norm without a norm vector");
  size_t need = 0;
  CGCN_TRY(cgcn_adj_build_workspace_bytes(m, n, k_pairs, &need));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("cgcn_adj_build: workspace %zu < %zu bytes", workspace_bytes, need);
    return CGCN_ERR_WORKSPACE;
  }
  AdjWs w;
  carve(w, workspace, workspace_bytes, m, sym_capacity(m, k_pairs));
  *nnz_host = 0;
  CGCN_CUDA(cudaMemsetAsync(w.err_count, 0, 16 * sizeof(int), st));
  if (m == 0 || n == 0) {
    CGCN_CUDA(cudaMemsetAsync(rowptr, 0, static_cast<size_t>(n + 1) * sizeof(int32_t), st));
    CGCN_CUDA(cudaStreamSynchronize(st));
    return CGCN_OK;
  }
  bool in_b = false;

  // 1. filter + normalise
  adj_filter_kernel<<<blocks_for(m), 256, 0, st>>>(bin1, bin2, val, m, window_starts, n, norm, norm_len, res_bp, use_norm,
                                                   w.key, w.v, w.flag, w.err_count);
  CGCN_TRY(check_launch("adj_filter_kernel"));
  CGCN_TRY((rsort::prefix_sum<int, true>(w.flag, w.incl, m, w.cub_temp, st)));
  const int64_t limit = (!use_norm && k_pairs > 0) ? k_pairs : -1;
  adj_compact_kernel<<<blocks_for(m), 256, 0, st>>>(w.key, w.v, w.flag, w.incl, m, limit, w.key_c, w.v_c, w.idx_c,
                                                    w.err_count);
  CGCN_TRY(check_launch("adj_compact_kernel"));
  int host_counts[2] = {0, 0};
  CGCN_CUDA(cudaMemcpyAsync(&host_counts[0], w.incl + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  CGCN_CUDA(cudaMemcpyAsync(&host_counts[1], w.err_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  CGCN_CUDA(cudaStreamSynchronize(st));
  int64_t ma = host_counts[0];
  if (limit >= 0 && ma > limit) ma = limit;
  if (host_counts[1] & 2) {
    set_error("cgcn_adj_build: a contact bin lies beyond the norm vector (IndexError in data/7create_graph_new.py:81-82)");
    return CGCN_ERR_DATA;
  }
  if (host_counts[1] & 1) {
    set_error("cgcn_adj_build: NaN contact value (sort order undefined in data/7create_graph_new.py:94)");
    return CGCN_ERR_DATA;
  }
  if (ma == 0) {
    CGCN_CUDA(cudaMemsetAsync(rowptr, 0, static_cast<size_t>(n + 1) * sizeof(int32_t), st));
    CGCN_CUDA(cudaStreamSynchronize(st));
    return CGCN_OK;
  }

  // 2. dict semantics: group by key
  const int key_bits = 32 + bits_for(n);
  CGCN_TRY(rsort::radix_sort<u64>(w.key_c, w.key_s, w.idx_c, w.idx_s, ma, 0, key_bits, w.cub_temp, st, &in_b));
  if (!in_b) {                                   // sorted data sits in the (key_c, idx_c) pair: make *_s the sorted one
    swap_ptr(w.key_c, w.key_s);
    swap_ptr(w.idx_c, w.idx_s);
  }
  adj_heads_kernel<<<blocks_for(ma), 256, 0, st>>>(w.key_s, ma, w.head);
  CGCN_TRY(check_launch("adj_heads_kernel"));
  CGCN_TRY((rsort::prefix_sum<int, true>(w.head, w.incl, ma, w.cub_temp, st)));
  adj_groups_kernel<<<blocks_for(ma), 256, 0, st>>>(w.key_s, w.idx_s, w.head, w.incl, w.v_c, ma, w.ukey, w.upos, w.uvkey,
                                                    w.iota);
  CGCN_TRY(check_launch("adj_groups_kernel"));
  int u_host = 0;
  CGCN_CUDA(cudaMemcpyAsync(&u_host, w.incl + (ma - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  CGCN_CUDA(cudaStreamSynchronize(st));
  const int64_t u = u_host;

  // 3. stable descending rank: by insertion position, then by value key
  CGCN_TRY(rsort::radix_sort<u32>(w.upos, w.upos_s, w.iota, w.perm1, u, 0, bits_for(ma), w.cub_temp, st, &in_b));
  if (!in_b) {
    swap_ptr(w.upos, w.upos_s);
    swap_ptr(w.iota, w.perm1);
  }
  adj_gather_vkey_kernel<<<blocks_for(u), 256, 0, st>>>(w.uvkey, w.perm1, u, w.vk_a);
  CGCN_TRY(check_launch("adj_gather_vkey_kernel"));
  CGCN_TRY(rsort::radix_sort<u64>(w.vk_a, w.vk_b, w.perm1, w.perm2, u, 0, 64, w.cub_temp, st, &in_b));
  if (!in_b) {
    swap_ptr(w.vk_a, w.vk_b);
    swap_ptr(w.perm1, w.perm2);
  }
  const int64_t ksel = (k_pairs > 0 && k_pairs < u) ? k_pairs : u;

  // 4. symmetrise, unique, CSR
  adj_symmetrise_kernel<<<blocks_for(ksel), 256, 0, st>>>(w.ukey, w.perm2, ksel, w.sym);
  CGCN_TRY(check_launch("adj_symmetrise_kernel"));
  CGCN_TRY(rsort::radix_sort<u64>(w.sym, w.sym_s, nullptr, nullptr, 2 * ksel, 0, key_bits, w.cub_temp, st, &in_b));
  if (!in_b) swap_ptr(w.sym, w.sym_s);
  adj_heads_kernel<<<blocks_for(2 * ksel), 256, 0, st>>>(w.sym_s, 2 * ksel, w.head);
  CGCN_TRY(check_launch("adj_heads_kernel"));
  CGCN_TRY((rsort::prefix_sum<int, true>(w.head, w.incl, 2 * ksel, w.cub_temp, st)));
  int nnz_h = 0;
  CGCN_CUDA(cudaMemcpyAsync(&nnz_h, w.incl + (2 * ksel - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  CGCN_CUDA(cudaStreamSynchronize(st));
  if (nnz_h > colidx_cap || (nnz_h > 0 && colidx == nullptr)) {
    set_error("cgcn_adj_build: colidx capacity %lld < %d entries", static_cast<long long>(colidx_cap), nnz_h);
    return CGCN_ERR_CAPACITY;
  }
  adj_emit_cols_kernel<<<blocks_for(2 * ksel), 256, 0, st>>>(w.sym_s, w.head, w.incl, 2 * ksel, colidx_cap, colidx, w.uniq);
  CGCN_TRY(check_launch("adj_emit_cols_kernel"));
  adj_rowptr_kernel<<<blocks_for(n + 1), 256, 0, st>>>(w.uniq, nnz_h, n, rowptr);
  CGCN_TRY(check_launch("adj_rowptr_kernel"));
  CGCN_CUDA(cudaStreamSynchronize(st));
  *nnz_host = nnz_h;
  return CGCN_OK;
}

// ----------------------------------------------------------------------------- A -> bin(A + I)
namespace cgcn {

__global__ void selfloop_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                                      int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = rowptr[i], e = rowptr[i + 1];
  int lo = s, hi = e;                        // columns are sorted: binary search for the diagonal
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (colidx[mid] < i) lo = mid + 1; else hi = mid;
  }
  const int has = (lo < e && colidx[lo] == i) ? 1 : 0;
  cnt[i] = (e - s) + 1 - has;
}

__global__ void selfloop_rowptr_kernel(const int* __restrict__ incl, int n, int32_t* __restrict__ rowptr_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  rowptr_out[i] = (i == 0) ? 0 : incl[i - 1];
}

// one warp per row: copy the row, inserting the diagonal at its sorted position
__global__ void selfloop_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                                     const int32_t* __restrict__ rowptr_out, int32_t* __restrict__ colidx_out) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int s = rowptr[i], e = rowptr[i + 1];
  const int os = rowptr_out[i];
  const int has = (rowptr_out[i + 1] - os) == (e - s);          // diagonal already stored
  for (int t = s + lane; t < e; t += 32) {
    const int c = colidx[t];
    colidx_out[os + (t - s) + ((!has && c > i) ? 1 : 0)] = c;
  }
  if (!has && lane == 0) {
    int lo = s, hi = e;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (colidx[mid] < i) lo = mid + 1; else hi = mid;
    }
    colidx_out[os + (lo - s)] = i;
  }
}

}  // namespace cgcn

extern "C" int cgcn_adj_add_selfloops_workspace_bytes(int32_t n, size_t* bytes_host) {
  CGCN_REQUIRE(bytes_host != nullptr && n >= 0, "cgcn_adj_add_selfloops_workspace_bytes: bad argument");
  const size_t t = rsort::scan_temp_bytes(n < 1 ? 1 : n);
  *bytes_host = align_up(static_cast<size_t>(n < 1 ? 1 : n) * sizeof(int), 256) * 2 + t + 1024;
  return CGCN_OK;
}

extern "C" int cgcn_adj_add_selfloops(const int32_t* rowptr, const int32_t* colidx, int32_t n, int32_t* rowptr_out,
                                      int32_t* colidx_out, void* workspace, size_t workspace_bytes,
                                      cgcn_stream_t stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CGCN_REQUIRE(rowptr && rowptr_out && colidx_out && n >= 1, "cgcn_adj_add_selfloops: bad argument");
  size_t need = 0;
  CGCN_TRY(cgcn_adj_add_selfloops_workspace_bytes(n, &need));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("cgcn_adj_add_selfloops: workspace %zu < %zu bytes", workspace_bytes, need);
    return CGCN_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes);
  int* cnt = a.take<int>(n);
  int* incl = a.take<int>(n);
  void* temp = a.take<char>(rsort::scan_temp_bytes(n));
  selfloop_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(rowptr, colidx, n, cnt);
  CGCN_TRY(check_launch("selfloop_count_kernel"));
  CGCN_TRY((rsort::prefix_sum<int, true>(cnt, incl, n, temp, st)));
  selfloop_rowptr_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(incl, n, rowptr_out);
  CGCN_TRY(check_launch("selfloop_rowptr_kernel"));
  const int64_t threads = static_cast<int64_t>(n) * 32;
  selfloop_fill_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(rowptr, colidx, n, rowptr_out, colidx_out);
  return check_launch("selfloop_fill_kernel");
}

// ----------------------------------------------------------------------------- COO -> CSR pattern
namespace cgcn {

__global__ void coo_keys_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, int64_t nnz, int n,
                                u64* __restrict__ key, u32* __restrict__ iota, int* __restrict__ flags) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < nnz;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = rows[t], c = cols[t];
    if (r < 0 || c < 0 || r >= n || c >= n) atomicOr(flags, 4);          // out of range
    key[t] = (static_cast<u64>(r) << 32) | static_cast<u64>(c & 0xffffffffll);
    iota[t] = static_cast<u32>(t);
  }
}

// flags bit 0 (cleared on violation): value == 1/deg(row);  bit 1 (cleared): symmetric;  bit 3: duplicates
__global__ void coo_check_kernel(const u64* __restrict__ key_s, const u32* __restrict__ perm, const float* __restrict__ vals,
                                 int64_t nnz, const int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                                 int* __restrict__ viol) {
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < nnz;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const u64 k = key_s[t];
    const int r = static_cast<int>(k >> 32), c = static_cast<int>(k & 0xffffffffull);
    colidx[t] = c;
    if (t > 0 && key_s[t - 1] == k) atomicOr(viol, 8);
    const int deg = rowptr[r + 1] - rowptr[r];
    if (vals[perm[t]] != __fdiv_rn(1.0f, static_cast<float>(deg))) atomicOr(viol, 1);
    const u64 tk = (static_cast<u64>(c) << 32) | static_cast<u64>(r);
    const int64_t p = lower_bound_u64(key_s, nnz, tk);
    if (p >= nnz || key_s[p] != tk) atomicOr(viol, 2);
  }
}

}  // namespace cgcn

extern "C" int cgcn_coo_to_pattern_workspace_bytes(int64_t nnz, int64_t n, size_t* bytes_host) {
  CGCN_REQUIRE(bytes_host != nullptr && nnz >= 0 && n >= 0 && nnz < 2147483647LL, "cgcn_coo_to_pattern_workspace_bytes: bad argument");
  const size_t c = static_cast<size_t>(nnz < 1 ? 1 : nnz);
  *bytes_host = align_up(c * 8, 256) * 2 + align_up(c * 4, 256) * 2 + cub_temp_bytes(static_cast<int64_t>(c)) + 4096;
  return CGCN_OK;
}

extern "C" int cgcn_coo_to_pattern(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz, int32_t n,
                                   int32_t* rowptr, int32_t* colidx, int32_t* flags_host, void* workspace,
                                   size_t workspace_bytes, cgcn_stream_t stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CGCN_REQUIRE(rows && cols && vals && rowptr && colidx && flags_host && nnz >= 1 && n >= 1, "cgcn_coo_to_pattern: bad argument");
  size_t need = 0;
  CGCN_TRY(cgcn_coo_to_pattern_workspace_bytes(nnz, n, &need));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("cgcn_coo_to_pattern: workspace %zu < %zu bytes", workspace_bytes, need);
    return CGCN_ERR_WORKSPACE;
  }
  Arena a(workspace, workspace_bytes);
  u64* key = a.take<u64>(nnz);
  u64* key_s = a.take<u64>(nnz);
  u32* iota = a.take<u32>(nnz);
  u32* perm = a.take<u32>(nnz);
  int* viol = a.take<int>(4);
  void* temp = a.take<char>(cub_temp_bytes(nnz));
  bool in_b = false;
  CGCN_CUDA(cudaMemsetAsync(viol, 0, 4 * sizeof(int), st));
  coo_keys_kernel<<<blocks_for(nnz), 256, 0, st>>>(rows, cols, nnz, n, key, iota, viol);
  CGCN_TRY(check_launch("coo_keys_kernel"));
  CGCN_TRY(rsort::radix_sort<u64>(key, key_s, iota, perm, nnz, 0, 32 + bits_for(n), temp, st, &in_b));
  if (!in_b) {
    swap_ptr(key, key_s);
    swap_ptr(iota, perm);
  }
  adj_rowptr_kernel<<<blocks_for(n + 1), 256, 0, st>>>(key_s, nnz, n, rowptr);
  CGCN_TRY(check_launch("adj_rowptr_kernel"));
  coo_check_kernel<<<blocks_for(nnz), 256, 0, st>>>(key_s, perm, vals, nnz, rowptr, colidx, viol);
  CGCN_TRY(check_launch("coo_check_kernel"));
  int v = 0;
  CGCN_CUDA(cudaMemcpyAsync(&v, viol, sizeof(int), cudaMemcpyDeviceToHost, st));
  CGCN_CUDA(cudaStreamSynchronize(st));
  if (v & 4) {
    set_error("cgcn_coo_to_pattern: index out of range");
    return CGCN_ERR_DATA;
  }
  if (v & 8) {
    set_error("cgcn_coo_to_pattern: duplicate (row, col) entries; coalesce the tensor first");
    return CGCN_ERR_DATA;
  }
  *flags_host = ((v & 1) ? 0 : 1) | ((v & 2) ? 0 : 2);
  return CGCN_OK;
}
