// fp32 CUDA-core (FFMA) dense contractions for the GCN path: the always-available, exactly-fp32
// implementation of the two shapes the model needs.  gemm_tc.cu provides the tcgen05 (3xTF32)
// versions of the same entry points; this file is the shape-generic one (any k, n <= 128, any
// leading dimension) and the accuracy yardstick for the tensor-core kernels.
//
//   row panel :  C[m x n]   = rowscale * (A[m x k] * op(B)) + bias     m huge, k,n <= 128
//                 torch.mm(input, weight)  models/SubLayers.py:43 ; self.out(x)  models/ChromeModels.py:51
//                 and their input gradients G W^T
//   gram      :  C[ka x nb] = sum_r A[r,:]^T (x) B[r,:]                 reduction over m rows
//                 weight gradients X^T G (autograd of the above)
#include "common.cuh"

namespace cgcn {

constexpr int GB = 128;          // tile edge (rows of the panel per CTA tile, and max n / k)
constexpr int GBK = 16;          // k-chunk
constexpr int GTHREADS = 256;    // 16 x 16 threads, 8 x 8 outputs each

// -------------------------------------------------------------------------------- row panel
struct RowPanelArgs {
  const float* A;
  int64_t lda;
  const float* B;
  int b_transposed;
  const float* bias;
  float* C;
  int64_t ldc;
  int64_t m;
  int n, k;
  const int32_t* rowscale_rowptr;
  const float* rowscale_inv;
  int rowscale_group;
  int a_vec, c_vec;              // 128-bit access legal on A rows / C rows
  int64_t ldb;                   // floats between consecutive rows of B as stored ([k][n], or [n][k] when transposed)
  int accumulate;                // C += ... (k-blocked contractions wider than 128)
};

__device__ __forceinline__ void fma_8x8(float (&acc)[8][8], const float4 a0, const float4 a1, const float4 b0,
                                        const float4 b1) {
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
}

__global__ void __launch_bounds__(GTHREADS, 2) gemm_rowpanel_kernel(const RowPanelArgs p) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float smem[];
  float* Bs = smem;                       // [k_pad][GB]
  const int k_pad = (p.k + GBK - 1) / GBK * GBK;
  float* As = smem + k_pad * GB;          // [2][GBK][GB]   (transposed: As[kk][row])
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // B resident for the whole CTA lifetime, zero padded to [k_pad][128]
  for (int idx = tid; idx < k_pad * GB; idx += GTHREADS) {
    int kk, j;
    if (p.b_transposed) {
      j = idx / k_pad;
      kk = idx - j * k_pad;
    } else {
      kk = idx / GB;
      j = idx - kk * GB;
    }
    float v = 0.f;
    if (kk < p.k && j < p.n) v = p.b_transposed ? __ldg(p.B + static_cast<size_t>(j) * p.ldb + kk)
                                                : __ldg(p.B + static_cast<size_t>(kk) * p.ldb + j);
    Bs[kk * GB + j] = v;
  }

  const int64_t tiles = (p.m + GB - 1) / GB;
  const int chunks = k_pad / GBK;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row0 = tile * GB;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // A chunk loader: 128 rows x 16 floats = 512 float4, two per thread
    float4 pre[2];
    auto load_chunk = [&](int c) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int f = tid + q * GTHREADS;
        const int r = f >> 2, c4 = f & 3;
        const int64_t grow = row0 + r;
        const int kcol = c * GBK + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < p.m) {
          const float* src = p.A + grow * p.lda + kcol;
          if (p.a_vec && kcol + 3 < p.k) {
            v = ldg4(src);
          } else {
            if (kcol + 0 < p.k) v.x = __ldg(src + 0);
            if (kcol + 1 < p.k) v.y = __ldg(src + 1);
            if (kcol + 2 < p.k) v.z = __ldg(src + 2);
            if (kcol + 3 < p.k) v.w = __ldg(src + 3);
          }
        }
        pre[q] = v;
      }
    };
    auto store_chunk = [&](int buf) {
      float* dst = As + buf * GBK * GB;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int f = tid + q * GTHREADS;
        const int r = f >> 2, c4 = f & 3;
        dst[(c4 * 4 + 0) * GB + r] = pre[q].x;
        dst[(c4 * 4 + 1) * GB + r] = pre[q].y;
        dst[(c4 * 4 + 2) * GB + r] = pre[q].z;
        dst[(c4 * 4 + 3) * GB + r] = pre[q].w;
      }
    };

    load_chunk(0);
    __syncthreads();                      // previous tile finished reading As (and Bs is written)
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
      const int buf = c & 1;
      if (c + 1 < chunks) load_chunk(c + 1);
      const float* a_s = As + buf * GBK * GB;
      const float* b_s = Bs + c * GBK * GB;
#pragma unroll
      for (int kk = 0; kk < GBK; ++kk) {
        const float4 a0 = ld4(a_s + kk * GB + ty * 4);
        const float4 a1 = ld4(a_s + kk * GB + 64 + ty * 4);
        const float4 b0 = ld4(b_s + kk * GB + tx * 4);
        const float4 b1 = ld4(b_s + kk * GB + 64 + tx * 4);
        fma_8x8(acc, a0, a1, b0, b1);
      }
      if (c + 1 < chunks) {
        store_chunk(buf ^ 1);             // the other buffer was last read in iteration c-1
        __syncthreads();
      }
    }

    // epilogue
    float bias_r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
      bias_r[j] = (p.bias != nullptr && col < p.n) ? __ldg(p.bias + col) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
      const int64_t grow = row0 + r;
      if (grow >= p.m) continue;
      float s = 1.0f;
      if (p.rowscale_rowptr != nullptr || p.rowscale_inv != nullptr)
        s = row_scale(p.rowscale_rowptr, p.rowscale_inv, static_cast<int>(grow / p.rowscale_group));
      float* dst = p.C + grow * p.ldc;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = h * 64 + tx * 4;
        float4 v = make_float4(fmaf(acc[i][h * 4 + 0], s, bias_r[h * 4 + 0]), fmaf(acc[i][h * 4 + 1], s, bias_r[h * 4 + 1]),
                               fmaf(acc[i][h * 4 + 2], s, bias_r[h * 4 + 2]), fmaf(acc[i][h * 4 + 3], s, bias_r[h * 4 + 3]));
        if (p.c_vec && col + 3 < p.n) {
          if (p.accumulate) {
            const float4 o = ld4(dst + col);
            v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
          }
          st4(dst + col, v);
        } else if (p.accumulate) {
          if (col + 0 < p.n) dst[col + 0] += v.x;
          if (col + 1 < p.n) dst[col + 1] += v.y;
          if (col + 2 < p.n) dst[col + 2] += v.z;
          if (col + 3 < p.n) dst[col + 3] += v.w;
        } else {
          if (col + 0 < p.n) dst[col + 0] = v.x;
          if (col + 1 < p.n) dst[col + 1] = v.y;
          if (col + 2 < p.n) dst[col + 2] = v.z;
          if (col + 3 < p.n) dst[col + 3] = v.w;
        }
      }
    }
  }
}

int gemm_rowpanel_ffma(const float* A, int64_t lda, const float* B, int b_transposed, const float* bias, float* C,
                       int64_t ldc, int64_t m, int n, int k, const int32_t* rowscale_rowptr, const float* rowscale_inv, int rowscale_group,
                       cudaStream_t stream, int64_t ldb, int accumulate) {
  if (ldb <= 0) ldb = b_transposed ? k : n;
  CGCN_REQUIRE(A && B && C, "cgcn_gemm_rowpanel: null operand");
  CGCN_REQUIRE(n >= 1 && n <= GB && k >= 1 && k <= GB, "cgcn_gemm_rowpanel: n=%d k=%d must be in [1,128]", n, k);
  CGCN_REQUIRE(lda >= k && ldc >= n, "cgcn_gemm_rowpanel: leading dimension too small");
  CGCN_REQUIRE((rowscale_rowptr == nullptr && rowscale_inv == nullptr) || rowscale_group >= 1, "cgcn_gemm_rowpanel: rowscale_group");
  if (m <= 0) return CGCN_OK;
  RowPanelArgs p{A, lda, B, b_transposed, bias, C, ldc, m, n, k, rowscale_rowptr, rowscale_inv, rowscale_group, 0, 0, ldb, accumulate};
  p.a_vec = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  p.c_vec = (ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0);
  const int k_pad = (k + GBK - 1) / GBK * GBK;
  const size_t smem = static_cast<size_t>(k_pad * GB + 2 * GBK * GB) * sizeof(float);
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    CGCN_CUDA(cudaFuncSetAttribute(gemm_rowpanel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (GB * GB + 2 * GBK * GB) * static_cast<int>(sizeof(float))));
  }
  const int64_t tiles = (m + GB - 1) / GB;
  const int grid = static_cast<int>(tiles < 2LL * sm_count() ? tiles : 2LL * sm_count());
  CGCN_CUDA(launch_k(gemm_rowpanel_kernel, dim3(grid), dim3(GTHREADS), smem, stream, p));
  return check_launch("gemm_rowpanel_kernel");
}

// -------------------------------------------------------------------------------- gram
constexpr int GRAM_MIN_ROWS = 256;    // rows per CTA (lower bound); the upper bound 2*SMs CTAs caps the partial tiles

struct GramArgs {
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  int64_t m;
  int ka, nb;
  int64_t rows_per_cta;
  float* partial;                // [gridDim.x][ka][nb]
  int a_vec, b_vec;
};

__global__ void __launch_bounds__(GTHREADS, 2) gemm_gram_kernel(const GramArgs p) {
  pdl_grid_sync();
  __shared__ __align__(16) float As[2][GBK][GB];
  __shared__ __align__(16) float Bs[2][GBK][GB];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * p.rows_per_cta;
  const int64_t r_end = min(r_begin + p.rows_per_cta, p.m);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // chunk = 16 rows x 128 cols of A and of B: 512 float4 each, two + two per thread
  float4 pa[2], pb[2];
  auto load_rows = [&](const float* M, int64_t ld, int width, int vec_ok, int64_t r0, float4 (&dst)[2]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * GTHREADS;
      const int r = f >> 5, c = (f & 31) * 4;
      const int64_t grow = r0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < r_end && c < width) {
        const float* src = M + grow * ld + c;
        if (vec_ok && c + 3 < width) {
          v = ldg4(src);
        } else {
          v.x = __ldg(src);
          if (c + 1 < width) v.y = __ldg(src + 1);
          if (c + 2 < width) v.z = __ldg(src + 2);
          if (c + 3 < width) v.w = __ldg(src + 3);
        }
      }
      dst[q] = v;
    }
  };
  auto store_rows = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = tid + q * GTHREADS;
      const int r = f >> 5, c = (f & 31) * 4;
      st4(&As[buf][r][c], pa[q]);
      st4(&Bs[buf][r][c], pb[q]);
    }
  };

  const int64_t nchunks = (r_end - r_begin + GBK - 1) / GBK;
  if (nchunks > 0) {
    load_rows(p.A, p.lda, p.ka, p.a_vec, r_begin, pa);
    load_rows(p.B, p.ldb, p.nb, p.b_vec, r_begin, pb);
    store_rows(0);
    __syncthreads();
    for (int64_t c = 0; c < nchunks; ++c) {
      const int buf = static_cast<int>(c & 1);
      if (c + 1 < nchunks) {
        load_rows(p.A, p.lda, p.ka, p.a_vec, r_begin + (c + 1) * GBK, pa);
        load_rows(p.B, p.ldb, p.nb, p.b_vec, r_begin + (c + 1) * GBK, pb);
      }
#pragma unroll
      for (int kk = 0; kk < GBK; ++kk) {
        const float4 a0 = ld4(&As[buf][kk][ty * 4]);
        const float4 a1 = ld4(&As[buf][kk][64 + ty * 4]);
        const float4 b0 = ld4(&Bs[buf][kk][tx * 4]);
        const float4 b1 = ld4(&Bs[buf][kk][64 + tx * 4]);
        fma_8x8(acc, a0, a1, b0, b1);
      }
      if (c + 1 < nchunks) {
        store_rows(buf ^ 1);
        __syncthreads();
      }
    }
  }
  float* dst = p.partial + static_cast<size_t>(blockIdx.x) * p.ka * p.nb;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    if (r >= p.ka) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
      if (col < p.nb) dst[r * p.nb + col] = acc[i][j];
    }
  }
}

// C (+)= sum over parts, fixed order: CTA = 64 output elements x 8 slices of the parts axis.
// Part q holds its tile at partial + q*part_stride, element (i, j) at i*pld + j.
__global__ void __launch_bounds__(512) gram_finalize_kernel(const float* __restrict__ partial, int parts, int part_stride,
                                                            int pld, int ka, int nb, float* __restrict__ C, int64_t ldc,
                                                            int accumulate) {
  pdl_grid_sync();
  __shared__ double sh[8][64];
  const int idx = blockIdx.x * 64 + threadIdx.x;
  const int count = ka * nb;
  const int i = idx / nb, j = idx - i * nb;
  const size_t off = static_cast<size_t>(i) * pld + j;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (idx < count) {
    int q = threadIdx.y;
    for (; q + 24 < parts; q += 32) {
      a0 += static_cast<double>(partial[static_cast<size_t>(q) * part_stride + off]);
      a1 += static_cast<double>(partial[static_cast<size_t>(q + 8) * part_stride + off]);
      a2 += static_cast<double>(partial[static_cast<size_t>(q + 16) * part_stride + off]);
      a3 += static_cast<double>(partial[static_cast<size_t>(q + 24) * part_stride + off]);
    }
    for (; q < parts; q += 8) a0 += static_cast<double>(partial[static_cast<size_t>(q) * part_stride + off]);
  }
  sh[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y == 0 && idx < count) {
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sh[y][threadIdx.x];
    float* dst = C + static_cast<int64_t>(i) * ldc + j;
    *dst = accumulate ? static_cast<float>(static_cast<double>(*dst) + s) : static_cast<float>(s);
  }
}

void gram_finalize_launch(const float* partial, int parts, int part_stride, int pld, int ka, int nb, float* C,
                          int64_t ldc, int accumulate, cudaStream_t stream) {
  launch_k(gram_finalize_kernel, dim3((ka * nb + 63) / 64), dim3(64, 8), 0, stream, partial, parts, part_stride, pld, ka, nb, C, ldc,
           accumulate);
}

int64_t gram_rows_per_cta(int64_t m) {
  const int64_t target = (m + 2LL * sm_count() - 1) / (2LL * sm_count());
  int64_t rows = target > GRAM_MIN_ROWS ? target : GRAM_MIN_ROWS;
  return (rows + GBK - 1) / GBK * GBK;
}

size_t gram_workspace_bytes(int64_t m) {
  if (m <= 0) return 256;
  const int64_t rows = gram_rows_per_cta(m);
  const int64_t parts = (m + rows - 1) / rows;
  return static_cast<size_t>(parts) * GB * GB * sizeof(float) + 256;
}

int gemm_gram_ffma(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t m, int ka,
                   int nb, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  CGCN_REQUIRE(A && B && C, "cgcn_gemm_gram: null operand");
  CGCN_REQUIRE(ka >= 1 && ka <= GB && nb >= 1 && nb <= GB, "cgcn_gemm_gram: ka=%d nb=%d must be in [1,128]", ka, nb);
  CGCN_REQUIRE(lda >= ka && ldb >= nb && ldc >= nb, "cgcn_gemm_gram: leading dimension too small");
  CGCN_REQUIRE(m >= 1, "cgcn_gemm_gram: m must be positive");
  if (workspace == nullptr || workspace_bytes < gram_workspace_bytes(m)) {
    set_error("cgcn_gemm_gram: workspace %zu < %zu bytes", workspace_bytes, gram_workspace_bytes(m));
    return CGCN_ERR_WORKSPACE;
  }
  GramArgs p{A, lda, B, ldb, m, ka, nb, gram_rows_per_cta(m), static_cast<float*>(workspace), 0, 0};
  p.a_vec = (lda % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  p.b_vec = (ldb % 4 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0);
  const int parts = static_cast<int>((m + p.rows_per_cta - 1) / p.rows_per_cta);
  CGCN_CUDA(launch_k(gemm_gram_kernel, dim3(parts), dim3(GTHREADS), 0, stream, p));
  CGCN_TRY(check_launch("gemm_gram_kernel"));
  gram_finalize_launch(p.partial, parts, ka * nb, nb, ka, nb, C, ldc, accumulate, stream);
  return check_launch("gram_finalize_kernel");
}

}  // namespace cgcn
