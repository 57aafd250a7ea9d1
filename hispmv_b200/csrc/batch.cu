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
// treatment (gemm_lite_kernel): A is streamed once for up to eight vectors.
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
__global__ void __launch_bounds__(256) spmm_csr_kernel(CsrDev A, const float* __restrict__ xi, float* __restrict__ y,
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
    if (r < A.rows) {
      const int b = A.row_ptr[r], e = A.row_ptr[r + 1];
      int j = b + lane;
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
    if (lane == 0 && r < A.rows) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < nv) y[(int64_t)k * A.rows + r] = finish(acc[k], ep.alpha, ep.beta, ep.bias, r, ep.relu);
    }
  }
}

// Dense overlay with several vectors ("GeMM-lite"): a warp owns R consecutive rows and sweeps the columns four at a
// time; the 4K interleaved x values of those columns come through L1 once and feed all R rows and K vectors.
template <int K, int R>
__global__ void __launch_bounds__(256) gemm_lite_kernel(DenseDev A, const float* __restrict__ xi, float* __restrict__ y,
                                                        int nv, Epilogue ep) {
  const uint64_t ps = policy_evict_first();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * R;
  if (r0 >= A.rows) return;
  const int nr = (int)min((int64_t)R, (int64_t)A.rows - r0);
  const int ncol4 = (int)(A.ld >> 2);
  float acc[R][K];
#pragma unroll
  for (int q = 0; q < R; ++q)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[q][k] = 0.0f;
  const float* arow = A.a + r0 * A.ld;
  for (int c4 = lane; c4 < ncol4; c4 += 32) {
    float4 a[R];
#pragma unroll
    for (int q = 0; q < R; ++q)
      a[q] = q < nr ? ld_stream_f4(arow + (int64_t)q * A.ld + 4 * c4, ps) : make_float4(0.f, 0.f, 0.f, 0.f);
    float xv[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j) load_xk<K>(xi, 4 * c4 + j, xv[j]);
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc[q][k] = fmaf(a[q].x, xv[0][k], acc[q][k]);
        acc[q][k] = fmaf(a[q].y, xv[1][k], acc[q][k]);
        acc[q][k] = fmaf(a[q].z, xv[2][k], acc[q][k]);
        acc[q][k] = fmaf(a[q].w, xv[3][k], acc[q][k]);
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
int launch_gemm_lite_k(const DenseDev& A, const float* xi, float* y, int nv, Epilogue ep, cudaStream_t s) {
  constexpr int R = 2;
  const int64_t warps = (A.rows + R - 1) / R;
  const int grid = (int)((warps + 7) / 8);
  gemm_lite_kernel<K, R><<<grid, 256, 0, s>>>(A, xi, y, nv, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

template <int K>
int launch_spmm_k(const CsrDev& A, int lanes, const float* xi, float* y, int nv, Epilogue ep, cudaStream_t s) {
  const int64_t threads = (int64_t)A.rows * lanes;
  const int grid = (int)std::min<int64_t>((threads + 255) / 256, 148 * 64);
  switch (lanes) {
    case 2: spmm_csr_kernel<K, 2><<<grid, 256, 0, s>>>(A, xi, y, nv, ep); break;
    case 4: spmm_csr_kernel<K, 4><<<grid, 256, 0, s>>>(A, xi, y, nv, ep); break;
    case 8: spmm_csr_kernel<K, 8><<<grid, 256, 0, s>>>(A, xi, y, nv, ep); break;
    case 16: spmm_csr_kernel<K, 16><<<grid, 256, 0, s>>>(A, xi, y, nv, ep); break;
    default: spmm_csr_kernel<K, 32><<<grid, 256, 0, s>>>(A, xi, y, nv, ep); break;
  }
  HISPMV_CUDA(cudaGetLastError());
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

int launch_spmm_csr(const CsrDev& A, int lanes, const float* xi, float* y, int nv, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  if (ep.y_mc) {
    set_error("spmm: a multicast y is not supported for batches");
    return HISPMV_ERR_STATE;
  }
  switch (batch_width(nv)) {
    case 8: return launch_spmm_k<8>(A, lanes, xi, y, nv, ep, s);
    case 4: return launch_spmm_k<4>(A, lanes, xi, y, nv, ep, s);
    default: return launch_spmm_k<2>(A, lanes, xi, y, nv, ep, s);
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
