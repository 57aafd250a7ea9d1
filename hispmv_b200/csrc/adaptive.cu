// Tile kernels for sm_100a: y = alpha * A x + beta * bias over the row-aligned, nnz-balanced tiles of an
// AdaptivePlan (partition.cu: adaptive_tiles_device).  What they replace in the reference (semantics only):
//   ComputeAB / PreAccumulator   val * x[col], adder chain          automation_tool/assets/base_functions.cpp:228-327
//   ADD / SWB / SSW              shared-row partial sums + routing   base_functions.cpp:356-436
//   AccumBuffer / Compute_C      y_Ax[row] += ..., alpha/beta        base_functions.cpp:475-488, 535
//   balanceWorkload / prepareTile (the balanced schedule)            common/src/spmv-helper.cpp:265-347, 517-638
//
// Measured facts that shape them (tools/gather_bench.cu, tools/dsmem_bench.cu):
//   * an SM takes in only ~0.95 L1-miss sectors per clock, whatever their payload (4 or 16 bytes) and whether or not
//     they allocate in L1; a scattered x gather costs a whole 32-byte sector.  On gather-heavy matrices this, not
//     HBM, is the ceiling (C2: 107 M sectors per SpMV = 0.387 ms, the kernel takes 0.38), so lanes are mapped to
//     nonzeros in the order that makes them share sectors, and per-tile latency chains are kept short and numerous;
//   * hits served on-chip are nearly free next to that, distributed shared memory is slower than L2 for 4-byte
//     gathers, more than ~190 KB of shared memory per SM throttles the miss path, __threadfence() flushes L1.
//
// Tiles (TileDesc, 32 bytes, built once per plan so that a tile costs one metadata load instead of a chain of four):
//   STREAM  consecutive rows, each shorter than the long threshold, about stream_items row ends + nonzeros
//   LONG    one chunk (<= chunk_nnz nonzeros) of a row at or above the threshold; a row with several chunks is
//           "split": every chunk drops its partial sum in carry[tile] (a data flag, see finish_chunk), the chunk that
//           arrives last at the row's counter adds the partials in chunk order and writes y -- one launch, no atomics
//           on y, no fence, bit-reproducible.
//
// Kernels:
//   spmv_adaptive_kernel   nnz-major, one CTA per tile (the default for irregular rows)
//   spmv_rowstage_kernel   row-major behind a TMA-staged col/val stream (banded / stencil / FEM rows)
//   (three research kernels -- one warp per tile, resident CTAs with an x window in shared memory, a warp-specialised
//   TMA pipeline -- live in experimental.cu and are only built with HISPMV_EXPERIMENTAL=1)
// nnz-major: lane l of a warp takes nonzero base+l -- 128-byte coalesced col/val loads, and the 32 gathers of one
//   instruction cover consecutive nonzeros, which are column-sorted inside a row and therefore share sectors; products
//   go to shared memory and each warp then reduces a slice of the tile's rows (one lane per row of up to 16 products,
//   the whole warp for longer ones).
// row-major: LANES lanes walk each row, so the gathers of one instruction cover the same position of 32/LANES
//   consecutive rows -- on banded matrices consecutive columns, i.e. one or two lines instead of up to 32.
#include <limits.h>

#include <algorithm>

#include "device_utils.cuh"
#include "internal.h"
#include "tile_device.cuh"

namespace hispmv {

namespace {

// ================================================================================================================
// one CTA per tile
// ================================================================================================================
template <int CAP, bool SPLIT, int THREADS>  // CAP >= stream_items + long_threshold: products of a STREAM tile
__global__ void __launch_bounds__(THREADS, 2048 / THREADS)
    spmv_adaptive_kernel(CsrDev A, AdaptivePlan P, const float* __restrict__ x, float* __restrict__ y, Epilogue ep) {
  __shared__ float s_prod[CAP];
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t = P.tile_begin + blockIdx.x;
  const bool look_ahead = P.ahead > 0 && tid == THREADS - 1 && blockIdx.x + (int64_t)P.ahead < (int64_t)gridDim.x;
  TileDesc da;
  if (look_ahead) da = load_desc(P.desc + t + P.ahead);  // in flight together with this tile's own descriptor
  const TileDesc d = load_desc(P.desc + t);
  if (look_ahead && (da.chunk >= 0 || P.ahead_all)) prefetch_tile_l2(A, da);
  if (P.ahead < 0 && tid == THREADS - 1) prefetch_tile_l2(A, d);  // own tile: later loop iterations hit L2
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const GatherL1<SPLIT> gx{x, P.hot_cols, pk};
  if (d.chunk >= 0) {
    float acc = 0.0f;
    stream_products<THREADS>(A, gx, d.n0, d.n1, tid, ps, [&](int, float p) { acc += p; });
    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp != 0) return;
    float total = lane < WARPS ? s_red[lane] : 0.0f;
    total = warp_sum(total);
    finish_chunk(P.carry, P.counter, d, t, total, lane, y, ep);
    return;
  }
  const int n0 = d.n0;
  // the extents of this lane's first row are requested before the stream so that their DRAM round trip overlaps it
  const int trows = d.r1 - d.r0;
  const int rpw = (trows + WARPS - 1) / WARPS;
  const int beg = warp * rpw, end = min(trows, beg + rpw);
  int b0 = 0, e0 = 0;
  float bias0 = 0.0f;
  if (beg + lane < end) {
    b0 = A.row_ptr[d.r0 + beg + lane];
    e0 = A.row_ptr[d.r0 + beg + lane + 1];
    if (ep.beta != 0.0f) bias0 = ep.bias[d.r0 + beg + lane];
  }
  stream_products<THREADS>(A, gx, n0, d.n1, tid, ps, [&](int i, float p) { s_prod[i - n0] = p; });
  __syncthreads();
  rows_from_products(A, d.r0, n0, beg, end, b0 - n0, e0 - n0, bias0, s_prod, lane, y, ep);
}

// ================================================================================================================
// row-major behind a TMA-staged stream.  Dynamic shared memory: col[CAP + 8] | val[CAP + 8].
// ================================================================================================================
template <int CAP, int LANES, int THREADS>
__global__ void __launch_bounds__(THREADS)
    spmv_rowstage_kernel(CsrDev A, AdaptivePlan P, const float* __restrict__ x, float* __restrict__ y, Epilogue ep) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  int* s_col = reinterpret_cast<int*>(s_raw);
  float* s_val = reinterpret_cast<float*>(s_raw) + (CAP + 8);
  __shared__ float s_red[THREADS / 32];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t = P.tile_begin + blockIdx.x;
  const TileDesc d = load_desc(P.desc + t);
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const GatherL1<false> gx{x, P.hot_cols, pk};  // banded matrices reuse every gathered line a few rows later
  if (d.chunk >= 0) {
    float acc = 0.0f;
    stream_products<THREADS>(A, gx, d.n0, d.n1, tid, ps, [&](int, float p) { acc += p; });
    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp != 0) return;
    float total = lane < THREADS / 32 ? s_red[lane] : 0.0f;
    total = warp_sum(total);
    finish_chunk(P.carry, P.counter, d, t, total, lane, y, ep);
    return;
  }
  const int r0 = d.r0, trows = d.r1 - d.r0;
  const int a0 = d.n0 & ~3;                 // 16-byte aligned window [a0, a1) around the tile's nonzeros;
  const int cnt = ((d.n1 + 3) & ~3) - a0;   // col/val are zero-padded past nnz, so the tail is readable
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (tid == 0 && cnt > 0) {
    mbar_expect_tx(&s_bar, (uint32_t)cnt * 8u);
    bulk_g2s_hint(s_col, A.col + a0, (uint32_t)cnt * 4u, &s_bar, ps);
    bulk_g2s_hint(s_val, A.val + a0, (uint32_t)cnt * 4u, &s_bar, ps);
  }
  constexpr int R = THREADS / LANES;  // rows per pass
  const int sub = tid % LANES, g = tid / LANES;
  // row extents of the first pass are fetched while the bulk copies fly
  int b = 0, e = 0;
  if (g < trows) {
    b = A.row_ptr[r0 + g] - a0;
    e = A.row_ptr[r0 + g + 1] - a0;
  }
  tile_staged(&s_bar, cnt);
  for (int rb = 0; rb < trows; rb += R) {
    const int i = rb + g;
    if (rb > 0) {
      b = e = 0;
      if (i < trows) {
        b = A.row_ptr[r0 + i] - a0;
        e = A.row_ptr[r0 + i + 1] - a0;
      }
    }
    float acc0 = 0.0f, acc1 = 0.0f;
    int k = b + sub;
    for (; k + LANES < e; k += 2 * LANES) {
      const int ca = s_col[k], cb = s_col[k + LANES];
      const float xa = gx(ca), xb = gx(cb);
      acc0 = fmaf(s_val[k], xa, acc0);
      acc1 = fmaf(s_val[k + LANES], xb, acc1);
    }
    if (k < e) acc0 = fmaf(s_val[k], gx(s_col[k]), acc0);
    const float acc = subwarp_sum<LANES>(acc0 + acc1);
    if (sub == 0 && i < trows) store_y(y, r0 + i, finish(acc, ep.alpha, ep.beta, ep.bias, r0 + i, ep.relu), ep.y_mc);
  }
}


// ---- launch helpers ----------------------------------------------------------------------------------------------
template <int CAP, int LANES, int THREADS>
int launch_rowstage_inst(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep,
                         cudaStream_t s) {
  auto k = spmv_rowstage_kernel<CAP, LANES, THREADS>;
  constexpr int smem = (CAP + 8) * 8;
  HISPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = (int)(P.tile_count >= 0 ? P.tile_count : P.num_tiles);
  if (grid <= 0) return HISPMV_OK;
  k<<<grid, THREADS, smem, s>>>(A, P, x, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}
template <int CAP, int THREADS>
int launch_rowstage_cap(const CsrDev& A, const AdaptivePlan& P, int lanes, const float* x, float* y, Epilogue ep,
                        cudaStream_t s) {
  switch (lanes) {
    case 1: return launch_rowstage_inst<CAP, 1, THREADS>(A, P, x, y, ep, s);
    case 2: return launch_rowstage_inst<CAP, 2, THREADS>(A, P, x, y, ep, s);
    case 4: return launch_rowstage_inst<CAP, 4, THREADS>(A, P, x, y, ep, s);
    case 8: return launch_rowstage_inst<CAP, 8, THREADS>(A, P, x, y, ep, s);
    case 16: return launch_rowstage_inst<CAP, 16, THREADS>(A, P, x, y, ep, s);
    case 32: return launch_rowstage_inst<CAP, 32, THREADS>(A, P, x, y, ep, s);
  }
  set_error("rowstage: lanes must be 1,2,4,8,16 or 32");
  return HISPMV_ERR_ARG;
}

int check_plan(const AdaptivePlan& P, int max_cap, const char* who) {
  if (P.num_tiles > INT_MAX) {
    set_error(std::string(who) + ": too many tiles");
    return HISPMV_ERR_ARG;
  }
  if (P.stream_items + P.long_threshold > max_cap || P.chunk_nnz <= 0 || !P.desc) {
    set_error(std::string(who) + ": plan exceeds the compiled shared-memory capacity");
    return HISPMV_ERR_ARG;
  }
  return HISPMV_OK;
}


}  // namespace

int launch_adaptive(const CsrDev& A, const AdaptivePlan& P, int threads, const float* x, float* y, Epilogue ep,
                    cudaStream_t s) {
  if (A.rows <= 0 || P.num_tiles <= 0) return HISPMV_OK;
  int st = check_plan(P, 4096, "adaptive");
  if (st != HISPMV_OK) return st;
  const int need = P.stream_items + P.long_threshold;
  const int grid = (int)(P.tile_count >= 0 ? P.tile_count : P.num_tiles);
  if (grid <= 0) return HISPMV_OK;
  const bool split = P.hot_cols != 0x7fffffff;
  if (threads == 128 && need <= 2048) {
    if (need <= 1024) {
      if (split) spmv_adaptive_kernel<1024, true, 128><<<grid, 128, 0, s>>>(A, P, x, y, ep);
      else spmv_adaptive_kernel<1024, false, 128><<<grid, 128, 0, s>>>(A, P, x, y, ep);
    } else {
      if (split) spmv_adaptive_kernel<2048, true, 128><<<grid, 128, 0, s>>>(A, P, x, y, ep);
      else spmv_adaptive_kernel<2048, false, 128><<<grid, 128, 0, s>>>(A, P, x, y, ep);
    }
  } else if (need <= 2048) {
    if (split) spmv_adaptive_kernel<2048, true, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
    else spmv_adaptive_kernel<2048, false, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
  } else if (need <= 3072) {
    if (split) spmv_adaptive_kernel<3072, true, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
    else spmv_adaptive_kernel<3072, false, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
  } else {
    if (split) spmv_adaptive_kernel<4096, true, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
    else spmv_adaptive_kernel<4096, false, kGroup><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int launch_rowstage(const CsrDev& A, const AdaptivePlan& P, int lanes, int threads, const float* x, float* y,
                    Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0 || P.num_tiles <= 0) return HISPMV_OK;
  int st = check_plan(P, kRowstageMaxCap, "rowstage");
  if (st != HISPMV_OK) return st;
  const int need = P.stream_items + P.long_threshold;
  if (threads == 128) {  // small tiles, up to 16 CTAs per SM: more independent load -> gather -> reduce chains in flight
    if (need <= 2048) return launch_rowstage_cap<2048, 128>(A, P, lanes, x, y, ep, s);
    return launch_rowstage_cap<4096, 128>(A, P, lanes, x, y, ep, s);
  }
  if (need <= 2048) return launch_rowstage_cap<2048, 256>(A, P, lanes, x, y, ep, s);
  if (need <= 4096) return launch_rowstage_cap<4096, 256>(A, P, lanes, x, y, ep, s);
  if (need <= 6144) return launch_rowstage_cap<6144, 256>(A, P, lanes, x, y, ep, s);
  return launch_rowstage_cap<8192, 256>(A, P, lanes, x, y, ep, s);
}


}  // namespace hispmv
