// Dense-overlay GeMV for sm_100a: y = alpha * A x + beta * bias, A row-major fp32, batch 1.
//
// Replaces the reference's dense overlay -- prepareDenseMtxForFPGA (common/src/spmv-helper.cpp:717-750)
// feeding ComputeAB in DENSE_MODE (automation_tool/assets/base_functions.cpp:188-226, out = a0*x[c] +
// a1*x[c+1]) and Compute_C (base_functions.cpp:535).  Batch-1 GeMV is 0.5 flop/byte: a pure HBM
// stream, so no tensor cores.  Design:
//   * A is stored with a leading dimension padded to 4 floats, so every row starts 16-byte aligned and
//     the stream is nothing but 128-bit evict-first loads (each element is read exactly once).
//   * x is read with 128-bit read-only loads and lives in L1 for all the rows a CTA owns (staging it in shared
//     memory by a TMA bulk copy is implemented too, but measured slower: see launch_gemv).
//   * A CTA owns a contiguous block of rows and all its threads sweep the columns of R rows at a time
//     (R*unroll independent 16-byte loads in flight per thread); a shuffle + shared-memory reduction
//     finishes each group of R rows.  Row blocks are sized so the grid is one wave of
//     sm_count * CTAS_PER_SM CTAs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "device_utils.cuh"
#include "internal.h"

namespace hispmv {

// four consecutive entries of x (zero past the end): one 128-bit read-only load when x is 16-byte aligned, which it is
// for cudaMalloc'ed vectors; x is tiny next to A and stays in L1 for all the rows a CTA owns
__device__ __forceinline__ float4 load_x4(const float* __restrict__ x, int c, int cols, bool x_vec) {
  if (x_vec && c + 3 < cols) return __ldg(reinterpret_cast<const float4*>(x + c));
  float4 v;
  v.x = c < cols ? __ldg(x + c) : 0.f;
  v.y = c + 1 < cols ? __ldg(x + c + 1) : 0.f;
  v.z = c + 2 < cols ? __ldg(x + c + 2) : 0.f;
  v.w = c + 3 < cols ? __ldg(x + c + 3) : 0.f;
  return v;
}

template <int THREADS, int R, bool STAGE_X>
__global__ void __launch_bounds__(THREADS)
    gemv_rowblock_kernel(DenseDev A, const float* __restrict__ x, float* __restrict__ y, Epilogue ep,
                         int groups_base, int groups_rem) {
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) float s_x[];  // STAGE_X: ld floats (cols rounded up to 4, tail zeroed)
  __shared__ float s_red[2][WARPS][R];
  __shared__ __align__(8) uint64_t s_bar;

  const uint64_t ps = policy_evict_first();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol4 = (int)(A.ld >> 2);
  const bool x_vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;

  if (STAGE_X) {
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int full = bulk_ok ? (A.cols & ~3) : 0;  // floats moved by the bulk engine
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0 && full > 0) {
      mbar_expect_tx(&s_bar, (uint32_t)full * 4u);
      // <= 64 KB per request keeps each copy well inside the engine's comfort zone
      for (int off = 0; off < full; off += 16384) {
        const int n = min(16384, full - off);
        bulk_g2s(s_x + off, x + off, (uint32_t)n * 4u, &s_bar);
      }
    }
    for (int c = full + tid; c < (int)A.ld; c += THREADS) s_x[c] = c < A.cols ? x[c] : 0.0f;
    if (full > 0) mbar_wait(&s_bar, 0);
    __syncthreads();
  }

  // CTA b owns groups_base (+1 for the first groups_rem CTAs) consecutive groups of R rows: the SMs' shares differ by
  // at most one group (equal row blocks of ceil(rows / grid) left some SMs a whole CTA short: 8 % on 8192 rows)
  const int64_t b = blockIdx.x;
  const int64_t gbeg = b * groups_base + min(b, (int64_t)groups_rem);
  const int64_t rbeg = gbeg * R;
  const int64_t rend = min((int64_t)A.rows, (gbeg + groups_base + (b < groups_rem ? 1 : 0)) * R);
  int buf = 0;
  for (int64_t r = rbeg; r < rend; r += R) {
    float acc[R];
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = 0.0f;
    const float* arow = A.a + r * A.ld;
    const int nr = (int)min((int64_t)R, rend - r);
    if (nr == R) {
#pragma unroll 2
      for (int c4 = tid; c4 < ncol4; c4 += THREADS) {
        float4 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = ld_stream_f4(arow + (int64_t)q * A.ld + 4 * c4, ps);
        const float4 xv = STAGE_X ? reinterpret_cast<const float4*>(s_x)[c4] : load_x4(x, 4 * c4, A.cols, x_vec);
#pragma unroll
        for (int q = 0; q < R; ++q) {
          acc[q] = fmaf(a[q].x, xv.x, acc[q]);
          acc[q] = fmaf(a[q].y, xv.y, acc[q]);
          acc[q] = fmaf(a[q].z, xv.z, acc[q]);
          acc[q] = fmaf(a[q].w, xv.w, acc[q]);
        }
      }
    } else {
      for (int c4 = tid; c4 < ncol4; c4 += THREADS) {
        const float4 xv = STAGE_X ? reinterpret_cast<const float4*>(s_x)[c4] : load_x4(x, 4 * c4, A.cols, x_vec);
#pragma unroll
        for (int q = 0; q < R; ++q) {
          if (q < nr) {
            const float4 a = ld_stream_f4(arow + (int64_t)q * A.ld + 4 * c4, ps);
            acc[q] = fmaf(a.x, xv.x, acc[q]);
            acc[q] = fmaf(a.y, xv.y, acc[q]);
            acc[q] = fmaf(a.z, xv.z, acc[q]);
            acc[q] = fmaf(a.w, xv.w, acc[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < R; ++q) s_red[buf][warp][q] = acc[q];
    }
    __syncthreads();
    if (tid < nr) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s += s_red[buf][w][tid];
      store_y(y, r + tid, finish(s, ep.alpha, ep.beta, ep.bias, r + tid, ep.relu), ep.y_mc);
    }
    buf ^= 1;  // the next group writes the other buffer, so one barrier per group is enough
  }
}

namespace {
template <int R, bool STAGE>
int launch_gemv_inst(const DenseDev& A, const float* x, float* y, Epilogue ep, int64_t max_grid, size_t x_bytes,
                     cudaStream_t s) {
  constexpr int THREADS = 256;
  const int64_t groups = ((int64_t)A.rows + R - 1) / R;
  const int64_t grid = std::min<int64_t>(max_grid, groups);
  const int groups_base = (int)(groups / grid), groups_rem = (int)(groups % grid);
  auto k = gemv_rowblock_kernel<THREADS, R, STAGE>;
  if (STAGE) {
    HISPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    k<<<(int)grid, THREADS, x_bytes, s>>>(A, x, y, ep, groups_base, groups_rem);
  } else {
    k<<<(int)grid, THREADS, 0, s>>>(A, x, y, ep, groups_base, groups_rem);
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}
}  // namespace

int launch_gemv(const DenseDev& A, const float* x, float* y, Epilogue ep, int sm_count, cudaStream_t s) {
  if (A.rows <= 0) return HISPMV_OK;
  const size_t x_bytes = (size_t)A.ld * sizeof(float);
  // x is read through L1 by default: staging it in shared memory (TMA bulk copy + mbarrier, kept as an option:
  // HISPMV_GEMV=c,r,1) costs a start-up barrier per CTA and measured slower on every shape (50000 x 10000: 0.381 ms
  // staged, 0.316 ms through L1).  Eight CTAs of 256 threads per SM, two rows per group: 1.0 of the measured HBM
  // peak on that shape; 8192^2 0.87, 8192 x 4096 0.73 (launch + ramp are a quarter of a 29 us kernel).
  bool stage = false;
  int ctas_per_sm = 8, r = 2;  // round-2 sweep (profiles/r2_gemv_sweep.txt): 8 x 2 beats 5 x 4 by 3-5 % on 8192^2 and 50000 x 10000
  if (const char* e = getenv("HISPMV_GEMV")) {  // "CTAS_PER_SM,R,STAGE" (development sweeps)
    int c = 0, rr = 0, st = 0;
    if (sscanf(e, "%d,%d,%d", &c, &rr, &st) == 3 && c >= 1 && c <= 8 && (rr == 2 || rr == 4 || rr == 8)) {
      ctas_per_sm = c;
      r = rr;
      stage = st != 0 && x_bytes <= 160 * 1024;
    }
  }
  if (stage) {
    while (ctas_per_sm > 1 && (x_bytes + 1024) * ctas_per_sm > 200 * 1024) --ctas_per_sm;
  }
  const int64_t grid = (int64_t)sm_count * ctas_per_sm;  // one wave; launch_gemv_inst spreads the row groups over it
  if (stage) {
    if (r == 2) return launch_gemv_inst<2, true>(A, x, y, ep, grid, x_bytes, s);
    if (r == 8) return launch_gemv_inst<8, true>(A, x, y, ep, grid, x_bytes, s);
    return launch_gemv_inst<4, true>(A, x, y, ep, grid, x_bytes, s);
  }
  if (r == 2) return launch_gemv_inst<2, false>(A, x, y, ep, grid, x_bytes, s);
  if (r == 8) return launch_gemv_inst<8, false>(A, x, y, ep, grid, x_bytes, s);
  return launch_gemv_inst<4, false>(A, x, y, ep, grid, x_bytes, s);
}

}  // namespace hispmv
