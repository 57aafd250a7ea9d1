// Several right-hand sides at once ("SpMM-lite"): Y[k] = alpha * A x_k + beta * bias for k < nv <= 8, one pass over A.
//
// The reference runs the vectors of a batch one after another (runLinear's loop, pyhispmv/src/fpga_handle.cpp:336,
// 366-379, called per layer by apps/fpga_layer_manager.py:58-67), streaming the matrix once per vector.  Here the
// vectors are interleaved, xi[c * K + k] = x_k[c], so that
//   * col/val are read once for all K vectors, and
//   * the K values a nonzero needs are one contiguous 4K-byte read: with K = 8 exactly the 32-byte sector that a
//     single-vector gather fetches for 4 useful bytes (the L2 sector rate is what binds).
// Sparse rows are walked by sub-warps of LANES lanes (the csr_vector shape), so the path is offered for matrices whose
// longest row a sub-warp can walk (DNN layers); others keep running vector by vector.  The dense overlay has the same
// treatment (gemm_lite_kernel): A is streamed once for up to eight vectors, x panels staged in shared memory.
#include <limits.h>

#include "device_utils.cuh"
#include "internal.h"

namespace hispmv {
namespace {

// x [nv][n] (row-major, nv <= K) -> xi [n_pad][K], zero for k >= nv and for rows n <= i < n_pad
template <int K>
__global__ void __launch_bounds__(256) interleave_kernel(const float* __restrict__ x, int nv, int64_t n, int64_t n_pad,
                                                         float* __restrict__ xi) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = (k < nv && i < n) ? x[(int64_t)k * n + i] : 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) xi[i * K + k] = v[k];
  }
}

template <int K>
__device__ __forceinline__ void load_xk(const float* __restrict__ xi, int c, float (&xv)[K]) {
  static_assert(K == 2 || K == 4 || K == 8, "batch width");
  const float* p = xi + (int64_t)c * K;  // 8 * K-byte aligned: one 8-, 16- or 32-byte read
  if constexpr (K == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
    xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
  } else if constexpr (K == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
  } else {
    const float2 a = __ldg(reinterpret_cast<const float2*>(p));
    xv[0] = a.x; xv[1] = a.y;
  }
}

template <int K, int LANES>
__global__ void __launch_bounds__(256) spmm_csr_kernel(CsrDev A, int skip_len, const float* __restrict__ xi, float* __restrict__ y,
                                                       int nv, Epilogue ep) {
  const uint64_t ps = policy_evict_first();
  const int lane = threadIdx.x & (LANES - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  const int64_t groups_per_warp = 32 / LANES;
  const int64_t warp_first = (group / groups_per_warp) * groups_per_warp;
  for (int64_t base = warp_first; base < A.rows; base += ngroups) {  // whole warps iterate together (shuffles below)
    const int64_t r = base + (group - warp_first);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0f;
    bool mine = r < A.rows;
    if (mine) {
      const int b = A.row_ptr[r], e = A.row_ptr[r + 1];
      mine = e - b <= skip_len;   // longer rows are left to one CTA each (spmm_csr_cta_kernel over the long-row list)
      int j = mine ? b + lane : e;
      for (; j + LANES < e; j += 2 * LANES) {  // two nonzeros, 2K products in flight per lane
        const int c0 = ld_stream_i1(A.col + j, ps), c1 = ld_stream_i1(A.col + j + LANES, ps);
        const float v0 = ld_stream_f1(A.val + j, ps), v1 = ld_stream_f1(A.val + j + LANES, ps);
        float x0[K], x1[K];
        load_xk<K>(xi, c0, x0);
        load_xk<K>(xi, c1, x1);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(v0, x0[k], acc[k]);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(v1, x1[k], acc[k]);
      }
      if (j < e) {
        const int c0 = ld_stream_i1(A.col + j, ps);
        const float v0 = ld_stream_f1(A.val + j, ps);
        float x0[K];
        load_xk<K>(xi, c0, x0);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(v0, x0[k], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = subwarp_sum<LANES>(acc[k]);
    if (lane == 0 && mine) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < nv) y[(int64_t)k * A.rows + r] = finish(acc[k], ep.alpha, ep.beta, ep.bias, r, ep.relu);
    }
  }
}

// Dense overlay with several vectors ("GeMM-lite").  Columns are cut into panels of kPanel4 * 4 = 512; the x values of
// a panel for all K vectors (16 KB at K = 8) are staged in shared memory once per CTA and reused by every row the CTA
// owns -- through L1 alone the 128-256 KB of interleaved x does not stay resident next to the stream of A, and the
// first version of this kernel ran at the speed of its L2 x reads (profiles/r1_batch_probe.txt).  x comes in the layout
// xp[panel][j * K + k][c4] (column 4 * (panel * kPanel4 + c4) + j, vector k): staging is a straight coalesced copy
// and the compute loop's shared-memory reads are conflict-free (consecutive lanes, consecutive c4).
constexpr int kPanel4 = 128;

template <int K>
__global__ void __launch_bounds__(256) interleave_panel_kernel(const float* __restrict__ x, int nv, int64_t n,
                                                               int64_t total, float* __restrict__ xp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
    const int c4l = (int)(o % kPanel4);
    const int64_t t = o / kPanel4;
    const int jk = (int)(t % (4 * K));
    const int64_t p = t / (4 * K);
    const int j = jk / K, k = jk % K;
    const int64_t c = 4 * (p * kPanel4 + c4l) + j;
    xp[o] = (k < nv && c < n) ? x[(int64_t)k * n + c] : 0.0f;
  }
}

template <int K, int R>  // a warp owns R consecutive rows, a CTA 8 * R
__global__ void __launch_bounds__(256) gemm_lite_kernel(DenseDev A, const float* __restrict__ xp, float* __restrict__ y,
                                                        int nv, Epilogue ep) {
  __shared__ __align__(16) float xs[4 * K][kPanel4];
  const uint64_t ps = policy_evict_first();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * R;
  const int nr = (int)max((int64_t)0, min((int64_t)R, (int64_t)A.rows - r0));
  const int ncol4 = (int)(A.ld >> 2);
  float acc[R][K];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[q][k] = 0.0f;
  const float* arow = A.a + r0 * A.ld;
  constexpr int T = kPanel4 / 32;  // column groups per lane and panel
  for (int p0 = 0, p = 0; p0 < ncol4; p0 += kPanel4, ++p) {
    // this panel's rows of A are requested first: their DRAM round trip overlaps the staging of x
    float4 a[T][R];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int c4 = p0 + lane + 32 * t;
        a[t][q] = (c4 < ncol4 && q < nr) ? ld_stream_f4(arow + (int64_t)q * A.ld + 4 * c4, ps)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    __syncthreads();  // every warp is done with the previous panel
    {
      const float4* src = reinterpret_cast<const float4*>(xp + (int64_t)p * (4 * K * kPanel4));
      float4* dst = reinterpret_cast<float4*>(&xs[0][0]);
      for (int i = tid; i < K * kPanel4; i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int c4l = lane + 32 * t;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float x0 = xs[0 * K + k][c4l], x1 = xs[1 * K + k][c4l], x2 = xs[2 * K + k][c4l], x3 = xs[3 * K + k][c4l];
#pragma unroll
        for (int q = 0; q < R; ++q) {
          acc[q][k] = fmaf(a[t][q].x, x0, acc[q][k]);
          acc[q][k] = fmaf(a[t][q].y, x1, acc[q][k]);
          acc[q][k] = fmaf(a[t][q].z, x2, acc[q][k]);
          acc[q][k] = fmaf(a[t][q].w, x3, acc[q][k]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[q][k] = warp_sum(acc[q][k]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (q < nr && k < nv)
          y[(int64_t)k * A.rows + r0 + q] = finish(acc[q][k], ep.alpha, ep.beta, ep.bias, r0 + q, ep.relu);
  }
}

template <int K>
int launch_gemm_lite_k(const DenseDev& A, const float* xp, float* y, int nv, Epilogue ep, cudaStream_t s) {
  // rows per warp: more rows share every staged panel, fewer keep small matrices spread over the SMs
  if (A.rows >= 148 * 32) {
    gemm_lite_kernel<K, 4><<<(int)(((int64_t)A.rows + 31) / 32), 256, 0, s>>>(A, xp, y, nv, ep);
  } else if (A.rows >= 148 * 16) {
    gemm_lite_kernel<K, 2><<<(int)(((int64_t)A.rows + 15) / 16), 256, 0, s>>>(A, xp, y, nv, ep);
  } else {
    gemm_lite_kernel<K, 1><<<(int)(((int64_t)A.rows + 7) / 8), 256, 0, s>>>(A, xp, y, nv, ep);
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

// Few, long rows (a 1024 x 8192 layer at density 0.25 has 1024 rows of ~2048 nonzeros): one CTA per row keeps all SMs
// busy where one warp per row would leave most of them idle.
template <int K>
__global__ void __launch_bounds__(256) spmm_csr_cta_kernel(CsrDev A, const int32_t* __restrict__ row_list, int64_t n_list,
                                                           const float* __restrict__ xi, float* __restrict__ y, int nv,
                                                           Epilogue ep) {
  __shared__ float s_red[8][K];
  const uint64_t ps = policy_evict_first();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t count = row_list ? n_list : (int64_t)A.rows;   // every row, or only the listed (long) ones
  for (int64_t i = blockIdx.x; i < count; i += gridDim.x) {
    const int64_t r = row_list ? (int64_t)row_list[i] : i;
    const int b = A.row_ptr[r], e = A.row_ptr[r + 1];
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0f;
    int j = b + tid;
    for (; j + 256 < e; j += 512) {
      const int c0 = ld_stream_i1(A.col + j, ps), c1 = ld_stream_i1(A.col + j + 256, ps);
      const float v0 = ld_stream_f1(A.val + j, ps), v1 = ld_stream_f1(A.val + j + 256, ps);
      float x0[K], x1[K];
      load_xk<K>(xi, c0, x0);
      load_xk<K>(xi, c1, x1);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(v0, x0[k], acc[k]);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(v1, x1[k], acc[k]);
    }
    if (j < e) {
      const int c0 = ld_stream_i1(A.col + j, ps);
      const float v0 = ld_stream_f1(A.val + j, ps);
      float x0[K];
      load_xk<K>(xi, c0, x0);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(v0, x0[k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = warp_sum(acc[k]);
    __syncthreads();  // the previous row's s_red has been read
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) s_red[warp][k] = acc[k];
    }
    __syncthreads();
    if (tid < K && tid < nv) {
      float t = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_red[w][tid];
      y[(int64_t)tid * A.rows + r] = finish(t, ep.alpha, ep.beta, ep.bias, r, ep.relu);
    }
  }
}

template <int K>
int launch_spmm_k(const CsrDev& A, int lanes, const int32_t* long_rows, int64_t n_long, int long_len, const float* xi,
                  float* y, int nv, Epilogue ep, cudaStream_t s) {
  if (lanes > 32) {  // one CTA per row
    spmm_csr_cta_kernel<K><<<(int)std::min<int64_t>(A.rows, 148 * 16), 256, 0, s>>>(A, nullptr, 0, xi, y, nv, ep);
    HISPMV_CUDA(cudaGetLastError());
    return HISPMV_OK;
  }
  const int64_t threads = (int64_t)A.rows * lanes;
  const int grid = (int)std::min<int64_t>((threads + 255) / 256, 148 * 64);
  const int skip = n_long > 0 ? long_len : INT_MAX;
  switch (lanes) {
    case 2: spmm_csr_kernel<K, 2><<<grid, 256, 0, s>>>(A, skip, xi, y, nv, ep); break;
    case 4: spmm_csr_kernel<K, 4><<<grid, 256, 0, s>>>(A, skip, xi, y, nv, ep); break;
    case 8: spmm_csr_kernel<K, 8><<<grid, 256, 0, s>>>(A, skip, xi, y, nv, ep); break;
    case 16: spmm_csr_kernel<K, 16><<<grid, 256, 0, s>>>(A, skip, xi, y, nv, ep); break;
    default: spmm_csr_kernel<K, 32><<<grid, 256, 0, s>>>(A, skip, xi, y, nv, ep); break;
  }
  HISPMV_CUDA(cudaGetLastError());
  if (n_long > 0) {  // the rows a sub-warp should not walk alone: one CTA each
    spmm_csr_cta_kernel<K><<<(int)std::min<int64_t>(n_long, 148 * 16), 256, 0, s>>>(A, long_rows, n_long, xi, y, nv, ep);
    HISPMV_CUDA(cudaGetLastError());
  }
  return HISPMV_OK;
}

}  // namespace

int batch_width(int nv) { return nv > 4 ? 8 : nv > 2 ? 4 : 2; }

int launch_interleave(const float* x, int nv, int64_t n, int64_t n_pad, float* xi, cudaStream_t s) {
  if (n_pad < n) n_pad = n;
  if (n_pad <= 0) return HISPMV_OK;
  const int grid = (int)std::min<int64_t>((n_pad + 255) / 256, 148 * 32);
  switch (batch_width(nv)) {
    case 8: interleave_kernel<8><<<grid, 256, 0, s>>>(x, nv, n, n_pad, xi); break;
    case 4: interleave_kernel<4><<<grid, 256, 0, s>>>(x, nv, n, n_pad, xi); break;
    default: interleave_kernel<2><<<grid, 256, 0, s>>>(x, nv, n, n_pad, xi); break;
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int64_t panel_floats(int64_t ld) {  // floats of the panel layout for kBatchMax = 8 vectors
  const int64_t panels = ((ld >> 2) + kPanel4 - 1) / kPanel4;
  return std::max<int64_t>(panels, 1) * 4 * 8 * kPanel4;
}

int launch_interleave_panels(const float* x, int nv, int64_t n, int64_t ld, float* xp, cudaStream_t s) {
  const int K = batch_width(nv);
  const int64_t panels = ((ld >> 2) + kPanel4 - 1) / kPanel4;
  const int64_t total = panels * 4 * K * kPanel4;
  if (total <= 0) return HISPMV_OK;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  switch (K) {
    case 8: interleave_panel_kernel<8><<<grid, 256, 0, s>>>(x, nv, n, total, xp); break;
    case 4: interleave_panel_kernel<4><<<grid, 256, 0, s>>>(x, nv, n, total, xp); break;
    default: interleave_panel_kernel<2><<<grid, 256, 0, s>>>(x, nv, n, total, xp); break;
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

namespace {
__global__ void long_rows_kernel(const int32_t* __restrict__ rp, int32_t rows, int32_t min_len, int32_t* __restrict__ out,
                                 int* __restrict__ count) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows && rp[r + 1] - rp[r] > min_len) out[atomicAdd(count, 1)] = (int32_t)r;
}
}  // namespace

// Rows with more than min_len nonzeros (any order: every row is independent); d_out cudaMalloc'ed by the callee.
int batch_long_rows_device(const int32_t* d_row_ptr, int32_t rows, int32_t min_len, int64_t max_count, int32_t** d_out,
                           int64_t* count, cudaStream_t stream) {
  *d_out = nullptr;
  *count = 0;
  if (rows <= 0 || max_count <= 0) return HISPMV_OK;
  int* d_cnt = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)d_out, (size_t)max_count * 4));
  HISPMV_CUDA(cudaMalloc((void**)&d_cnt, sizeof(int)));
  HISPMV_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), stream));
  long_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(d_row_ptr, rows, min_len, *d_out, d_cnt);
  int h = 0;
  int st = check_cuda(cudaMemcpyAsync(&h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, stream), "D2H", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamSynchronize(stream), "sync", __FILE__, __LINE__);
  cudaFree(d_cnt);
  if (st != HISPMV_OK) return st;
  *count = h;
  return HISPMV_OK;
}

int launch_spmm_csr(const CsrDev& A, int lanes, const int32_t* long_rows, int64_t n_long, int long_len, const float* xi,
                    float* y, int nv, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  if (ep.y_mc) {
    set_error("spmm: a multicast y is not supported for batches");
    return HISPMV_ERR_STATE;
  }
  switch (batch_width(nv)) {
    case 8: return launch_spmm_k<8>(A, lanes, long_rows, n_long, long_len, xi, y, nv, ep, s);
    case 4: return launch_spmm_k<4>(A, lanes, long_rows, n_long, long_len, xi, y, nv, ep, s);
    default: return launch_spmm_k<2>(A, lanes, long_rows, n_long, long_len, xi, y, nv, ep, s);
  }
}

int launch_gemm_lite(const DenseDev& A, const float* xi, float* y, int nv, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  if (ep.y_mc) {
    set_error("gemm_lite: a multicast y is not supported for batches");
    return HISPMV_ERR_STATE;
  }
  switch (batch_width(nv)) {
    case 8: return launch_gemm_lite_k<8>(A, xi, y, nv, ep, s);
    case 4: return launch_gemm_lite_k<4>(A, xi, y, nv, ep, s);
    default: return launch_gemm_lite_k<2>(A, xi, y, nv, ep, s);
  }
}

}  // namespace hispmv
