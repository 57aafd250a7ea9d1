// SpMV kernels for sm_100a: y = alpha * A x + beta * bias on a CSR row block.
//
// What they replace in the reference (semantics only -- the FPGA dataflow is not ported):
//   ComputeAB      val * x[col]                       automation_tool/assets/base_functions.cpp:228-241
//   PreAccumulator adder chain over same-row products  base_functions.cpp:307-327
//   ADD/SWB/SSW    shared-row partial sums + routing   base_functions.cpp:356-436
//   AccumBuffer    y_Ax[row] += product                base_functions.cpp:475-488
//   Compute_C      y = beta*c_in + alpha*(A x)         base_functions.cpp:535
//
// Three strategies, chosen per matrix by the runtime selector (partition.cu: select_kernel):
//   csr_scalar  one thread per row            -- very short regular rows
//   csr_vector  LANES-wide sub-warp per row    -- regular rows, shuffle reduction
//   (adaptive.cu holds the tile kernels -- adaptive, adaptive_persistent, rowstage -- the selector prefers)
//   merge       merge-path tiles of row-ends+nonzeros: every CTA gets the same amount of work no
//               matter how skewed the row lengths are; heavy rows are split across threads, warps
//               and CTAs.  Partial sums meet through a warp-shuffle segmented scan inside the CTA and
//               a carry-out array + fix-up kernel across CTAs (deterministic, no atomics).
#include <limits.h>

#include "device_utils.cuh"
#include "internal.h"

namespace hispmv {

// ------------------------------------------------------------------------------------------------
// csr_scalar
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spmv_csr_scalar_kernel(CsrDev A, const float* __restrict__ x,
                                                              float* __restrict__ y, Epilogue ep) {
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.rows; r += stride) {
    const int b = A.row_ptr[r], e = A.row_ptr[r + 1];
    float acc = 0.0f;
    for (int j = b; j < e; ++j)
      acc = fmaf(ld_stream_f1(A.val + j, ps), ld_x(x + ld_stream_i1(A.col + j, ps), pk), acc);
    store_y(y, r, finish(acc, ep.alpha, ep.beta, ep.bias, r, ep.relu), ep.y_mc);
  }
}

// ------------------------------------------------------------------------------------------------
// csr_vector: LANES threads cooperate on one row.
// ------------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(256) spmv_csr_vector_kernel(CsrDev A, const float* __restrict__ x,
                                                              float* __restrict__ y, Epilogue ep) {
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const int lane = threadIdx.x & (LANES - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  // all lanes of a warp run the same number of outer iterations so the shuffles below stay converged
  const int64_t groups_per_warp = 32 / LANES;
  const int64_t warp_first = (group / groups_per_warp) * groups_per_warp;
  for (int64_t base = warp_first; base < A.rows; base += ngroups) {
    const int64_t r = base + (group - warp_first);
    float acc = 0.0f;
    if (r < A.rows) {
      const int b = A.row_ptr[r], e = A.row_ptr[r + 1];
      int j = b + lane;
      // two independent gathers in flight per lane
      for (; j + LANES < e; j += 2 * LANES) {
        const int c0 = ld_stream_i1(A.col + j, ps), c1 = ld_stream_i1(A.col + j + LANES, ps);
        const float v0 = ld_stream_f1(A.val + j, ps), v1 = ld_stream_f1(A.val + j + LANES, ps);
        const float x0 = ld_x(x + c0, pk), x1 = ld_x(x + c1, pk);
        acc = fmaf(v0, x0, acc);
        acc = fmaf(v1, x1, acc);
      }
      if (j < e) acc = fmaf(ld_stream_f1(A.val + j, ps), ld_x(x + ld_stream_i1(A.col + j, ps), pk), acc);
    }
    acc = subwarp_sum<LANES>(acc);
    if (lane == 0 && r < A.rows) store_y(y, r, finish(acc, ep.alpha, ep.beta, ep.bias, r, ep.relu), ep.y_mc);
  }
}

// ------------------------------------------------------------------------------------------------
// merge-path kernel
// ------------------------------------------------------------------------------------------------
// One CTA per tile of THREADS*IPT merge items (a merge item is either one nonzero or one row end).
// Tile start coordinates (row, nnz offset) come precomputed from the partitioner (MergePlan).
//   phase 1  stage the tile's row-end offsets and val*x[col] products in shared memory
//            (128-bit streaming loads of val/col from the 16-byte-aligned address below the tile start)
//   phase 2  every thread binary-searches its own diagonal in shared memory and walks IPT items
//   phase 3  warp-shuffle segmented scan hands each thread the partial sum of the row its predecessors
//            left open; the tile's own open row goes to carry[tile]
//   phase 4  coalesced epilogue y = alpha*Ax + beta*bias for the rows that end in this tile
template <int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS)
    spmv_merge_kernel(CsrDev A, MergePlan P, const float* __restrict__ x, float* __restrict__ y, Epilogue ep) {
  constexpr int TILE = THREADS * IPT;
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_prod[TILE];
  __shared__ int s_rowend[TILE + 1];
  __shared__ float s_y[TILE];
  __shared__ float s_wv[WARPS];
  __shared__ int s_wf[WARPS];

  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const int tid = threadIdx.x;
  const int64_t t = blockIdx.x;
  const int r0 = P.tile_row[t], r1 = P.tile_row[t + 1];
  const int64_t n0 = P.tile_nnz[t], n1 = P.tile_nnz[t + 1];
  const int trows = r1 - r0;
  const int tnnz = (int)(n1 - n0);
  const int titems = trows + tnnz;

  // ---- phase 1a: products -------------------------------------------------------------------
  {
    const int64_t base = n0 & ~(int64_t)3;
    for (int64_t i = base + 4 * tid; i < n1; i += 4 * THREADS) {
      const int4 c = ld_stream_i4(A.col + i, ps);
      const float4 v = ld_stream_f4(A.val + i, ps);
      const int k = (int)(i - n0);  // may be -3..-1 for the first vector of the tile
      // the col/val arrays are zero-padded past nnz, so c.* is always a valid index into x
      const bool p0 = (k >= 0) & (k < tnnz), p1 = (k + 1 >= 0) & (k + 1 < tnnz);
      const bool p2 = (k + 2 >= 0) & (k + 2 < tnnz), p3 = (k + 3 < tnnz);
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
      if (p0) x0 = ld_x(x + c.x, pk);
      if (p1) x1 = ld_x(x + c.y, pk);
      if (p2) x2 = ld_x(x + c.z, pk);
      if (p3) x3 = ld_x(x + c.w, pk);
      if (p0) s_prod[k] = v.x * x0;
      if (p1) s_prod[k + 1] = v.y * x1;
      if (p2) s_prod[k + 2] = v.z * x2;
      if (p3) s_prod[k + 3] = v.w * x3;
    }
  }
  // ---- phase 1b: row ends, relative to the tile's first nonzero ------------------------------
  for (int i = tid; i <= trows; i += THREADS) {
    const int64_t r = (int64_t)r0 + i;
    int rel = INT_MAX;
    if (r < A.rows) {
      const int64_t d = (int64_t)A.row_ptr[r + 1] - n0;
      rel = d > (int64_t)TILE + 1 ? TILE + 1 : (int)d;
    }
    s_rowend[i] = rel;
  }
  __syncthreads();

  // ---- phase 2: per-thread merge walk ----------------------------------------------------------
  const int diag = min(tid * IPT, titems);
  int i, j;
  {
    int lo = max(diag - tnnz, 0), hi = min(diag, trows);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_rowend[mid] <= diag - mid - 1) lo = mid + 1; else hi = mid;
    }
    i = lo;
    j = diag - lo;
  }
  float acc = 0.0f, first_val = 0.0f;
  int first_i = -1;
  int row_end = s_rowend[i];
#pragma unroll
  for (int k = 0; k < IPT; ++k) {
    if (diag + k < titems) {
      if (j < row_end) {
        acc += s_prod[j];
        ++j;
      } else {
        if (first_i < 0) {
          first_i = i;
          first_val = acc;
        } else {
          s_y[i] = acc;
        }
        acc = 0.0f;
        ++i;
        row_end = s_rowend[i];
      }
    }
  }

  // ---- phase 3: segmented scan of (flag = "closed a row", value = open partial) ----------------
  const int lane = tid & 31, warp = tid >> 5;
  float v = acc;
  int f = first_i >= 0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float pv = __shfl_up_sync(kFullMask, v, d);
    const int pf = __shfl_up_sync(kFullMask, f, d);
    if (lane >= d) {
      if (!f) v += pv;
      f |= pf;
    }
  }
  if (lane == 31) {
    s_wv[warp] = v;
    s_wf[warp] = f;
  }
  float ev = __shfl_up_sync(kFullMask, v, 1);
  int ef = __shfl_up_sync(kFullMask, f, 1);
  if (lane == 0) {
    ev = 0.0f;
    ef = 0;
  }
  __syncthreads();
  float pv = 0.0f;  // open partial handed over by the preceding warps
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < warp) pv = s_wf[w] ? s_wv[w] : pv + s_wv[w];
  }
  const float carry_in = ef ? ev : pv + ev;
  if (first_i >= 0) s_y[first_i] = first_val + carry_in;
  if (tid == THREADS - 1) P.carry[t] = f ? v : pv + v;  // inclusive total = the tile's open row
  __syncthreads();

  // ---- phase 4: epilogue -------------------------------------------------------------------------
  // The first row of every tile but tile 0 may still receive a carry from earlier tiles: the fix-up
  // kernel finishes it (and applies the deferred ReLU).
  for (int q = tid; q < trows; q += THREADS) {
    const int64_t r = (int64_t)r0 + q;
    const int relu = ep.relu && !(q == 0 && t > 0);
    y[r] = finish(s_y[q], ep.alpha, ep.beta, ep.bias, r, relu);
  }
}

// One thread per tile boundary; the thread that starts a run of carries for a row sums the run in tile
// order (a warp-strided tree would reorder nothing observable: order is fixed either way) and
// finishes the row.  Runs are long only for rows that span many tiles (the ~10 heavy rows of C2).
__global__ void __launch_bounds__(256) spmv_merge_fixup_kernel(MergePlan P, int32_t rows, float* __restrict__ y,
                                                               Epilogue ep) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // boundary between tile t and t+1
  if (t >= P.num_tiles - 1) return;
  const int row = P.tile_row[t + 1];
  if (row >= rows) return;
  if (t > 0 && P.tile_row[t] == row) return;  // not the first carry of this row
  float s = 0.0f;
  for (int64_t u = t; u < P.num_tiles - 1 && P.tile_row[u + 1] == row; ++u) s += P.carry[u];
  float v = fmaf(ep.alpha, s, y[row]);
  if (ep.relu) v = fmaxf(v, 0.0f);
  y[row] = v;
}

__global__ void __launch_bounds__(256) spmv_empty_kernel(int32_t rows, float* __restrict__ y, Epilogue ep) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) store_y(y, r, finish(0.0f, ep.alpha, ep.beta, ep.bias, r, ep.relu), ep.y_mc);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline int grid_for(int64_t threads_needed, int block) {
  int64_t g = (threads_needed + block - 1) / block;
  const int64_t cap = 148LL * 8 * 64;  // grid-stride beyond this
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_empty(int32_t rows, float* y, Epilogue ep, cudaStream_t s) {
  if (rows <= 0) return HISPMV_OK;
  spmv_empty_kernel<<<(rows + 255) / 256, 256, 0, s>>>(rows, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int launch_csr_scalar(const CsrDev& A, const float* x, float* y, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  spmv_csr_scalar_kernel<<<grid_for(A.rows, 256), 256, 0, s>>>(A, x, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int launch_csr_vector(const CsrDev& A, int lanes, const float* x, float* y, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  const int grid = grid_for((int64_t)A.rows * lanes, 256);
  switch (lanes) {
    case 2: spmv_csr_vector_kernel<2><<<grid, 256, 0, s>>>(A, x, y, ep); break;
    case 4: spmv_csr_vector_kernel<4><<<grid, 256, 0, s>>>(A, x, y, ep); break;
    case 8: spmv_csr_vector_kernel<8><<<grid, 256, 0, s>>>(A, x, y, ep); break;
    case 16: spmv_csr_vector_kernel<16><<<grid, 256, 0, s>>>(A, x, y, ep); break;
    case 32: spmv_csr_vector_kernel<32><<<grid, 256, 0, s>>>(A, x, y, ep); break;
    default: set_error("csr_vector: lanes must be 2,4,8,16 or 32"); return HISPMV_ERR_ARG;
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

bool merge_tile_items_supported(int tile_items) {
  return tile_items == 128 * 7 || tile_items == 256 * 7 || tile_items == 256 * 11 || tile_items == 512 * 7;
}

int launch_merge(const CsrDev& A, const MergePlan& P, const float* x, float* y, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  if (P.num_tiles > INT_MAX) {
    set_error("merge: too many tiles");
    return HISPMV_ERR_ARG;
  }
  const int grid = (int)P.num_tiles;
  switch (P.tile_items) {
    case 128 * 7: spmv_merge_kernel<128, 7><<<grid, 128, 0, s>>>(A, P, x, y, ep); break;
    case 256 * 7: spmv_merge_kernel<256, 7><<<grid, 256, 0, s>>>(A, P, x, y, ep); break;
    case 256 * 11: spmv_merge_kernel<256, 11><<<grid, 256, 0, s>>>(A, P, x, y, ep); break;
    case 512 * 7: spmv_merge_kernel<512, 7><<<grid, 512, 0, s>>>(A, P, x, y, ep); break;
    default: set_error("merge: unsupported tile_items"); return HISPMV_ERR_ARG;
  }
  HISPMV_CUDA(cudaGetLastError());
  if (P.num_tiles > 1) {
    const int64_t nb = P.num_tiles - 1;
    spmv_merge_fixup_kernel<<<(int)((nb + 255) / 256), 256, 0, s>>>(P, A.rows, y, ep);
    HISPMV_CUDA(cudaGetLastError());
  }
  return HISPMV_OK;
}

}  // namespace hispmv
