// Device-side pieces shared by the tile kernels (adaptive.cu) and the blocked strategy (blocked.cu).
#pragma once
#include "device_utils.cuh"
#include "internal.h"

namespace hispmv {

// ---- the end of a LONG tile: warp 0 of the group holds the chunk's total in every lane ---------------------------
// Split rows meet without a fence on the arrival path: carry[] slots hold a sentinel (a NaN payload no arithmetic
// produces) until their chunk writes its partial; the chunk that arrives last at the row's counter reads the slots in
// chunk order past L1, waiting out any slot whose write is still in flight, adds them in that order and re-arms slots
// and counter for the next launch.  (The first version published carry[] with __threadfence(): on sm_100 that is
// MEMBAR.SC + CCTL.IVALL -- a ~2 k-cycle stall and an L1 flush per chunk.)
constexpr unsigned int kCarryEmpty = kCarryEmptyBits;
__device__ __forceinline__ void finish_chunk(float* carry, unsigned int* counter, const TileDesc& d, int64_t t,
                                             float total, int lane, float* __restrict__ y, const Epilogue& ep) {
  if (d.nchunks == 1) {
    if (lane == 0) store_y(y, d.r0, finish(total, ep.alpha, ep.beta, ep.bias, d.r0, ep.relu), ep.y_mc);
    return;
  }
  const int64_t first = t - d.chunk;  // tile id of this row's chunk 0
  int last = 0;
  if (lane == 0) {
    unsigned int bits = __float_as_uint(total);
    if (bits == kCarryEmpty) bits = 0x7fffffffu;  // a NaN is a NaN
    __stcg(reinterpret_cast<unsigned int*>(carry) + t, bits);
    const unsigned int prev = atomicAdd(&counter[first], 1u);
    last = (prev == (unsigned int)(d.nchunks - 1));
  }
  last = __shfl_sync(kFullMask, last, 0);
  if (!last) return;
  unsigned int* slots = reinterpret_cast<unsigned int*>(carry) + first;
  float s = 0.0f;
  for (int k = lane; k < d.nchunks; k += 32) {
    unsigned int bits = __ldcg(slots + k);
    while (bits == kCarryEmpty) {
      __nanosleep(40);
      bits = __ldcg(slots + k);
    }
    s += __uint_as_float(bits);
    __stcg(slots + k, kCarryEmpty);
  }
  s = warp_sum(s);
  if (lane == 0) {
    store_y(y, d.r0, finish(s, ep.alpha, ep.beta, ep.bias, d.r0, ep.relu), ep.y_mc);
    counter[first] = 0;  // ready for the next run / graph replay
  }
}

__device__ __forceinline__ TileDesc load_desc(const TileDesc* p) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  TileDesc d;
  d.r0 = a.x;
  d.r1 = a.y;
  d.n0 = a.z;
  d.n1 = a.w;
  d.chunk = b.x;
  d.nchunks = b.y;
  d.tile = b.z;
  d.pad = b.w;
  return d;
}


// ---- pieces shared by the tile kernels (adaptive.cu) and the research kernels (experimental.cu) -------------------
namespace {

constexpr int kGroup = 256;  // threads that cooperate on one tile
constexpr int kGroupWarps = kGroup / 32;
constexpr int kSerialRow = 16;  // rows up to this many nonzeros are summed by one lane, longer ones by the warp

// ---- x gathers ------------------------------------------------------------------------------------------------
template <bool SPLIT>
struct GatherL1 {  // SPLIT: columns below `hot` are pinned in L1 and the rest skips L1 allocation; else keep every line
  const float* x;
  int hot;
  uint64_t pk;
  __device__ __forceinline__ float operator()(int c) const {
#ifdef HISPMV_DIAG
    if (hot == -1) return 1.0f;                                    // diagnostics: no gathers at all
    if (hot <= -2) return c < -hot ? 1.0f : ld_x_keep(x + c, pk);  // diagnostics: no gathers below -hot
#endif
    if (SPLIT) return ld_x_split(x, c, hot, pk);
    return ld_x_keep(x + c, pk);
  }
};
struct GatherWindow {  // columns below `hot` live in shared memory
  const float* x;
  const float* s_x;
  int hot;
  uint64_t pk;
  __device__ __forceinline__ float operator()(int c) const { return c < hot ? s_x[c] : ld_x_bypass(x + c, pk); }
};

// ---- products of the nonzeros [n0, n1): lane-consecutive, four independent (col, val, x) triples per thread ----
// OUT(i, p) receives product p of nonzero i.  Threads whose four slots are all inside [n0, n1) take a path without
// per-slot predicates (instruction issue, not memory, was the floor of the first version of this loop).
template <int STRIDE = kGroup, class G, class OUT>
__device__ __forceinline__ void stream_products(const CsrDev& A, const G& gx, int n0, int n1, int gt, uint64_t ps,
                                                OUT out) {
  constexpr int kGroup = STRIDE;  // threads sharing the range (a CTA group, or one warp)
  for (int i0 = (n0 & ~31) + gt; i0 < n1; i0 += 4 * kGroup) {  // every warp load is one aligned 128-byte line
    const int32_t* pc = A.col + i0;
    const float* pv = A.val + i0;
    if (i0 >= n0 && i0 + 3 * kGroup < n1) {
      int c[4];
      float v[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = ld_stream_i1(pc + u * kGroup, ps);
        v[u] = ld_stream_f1(pv + u * kGroup, ps);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = gx(c[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) out(i0 + u * kGroup, v[u] * xv[u]);
    } else {  // first / last slots of the tile: same three phases, predicated per slot
      int c[4];
      float v[4], xv[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kGroup;
        ok[u] = (i >= n0) & (i < n1);
        c[u] = 0;
        v[u] = 0.0f;
        if (ok[u]) {
          c[u] = ld_stream_i1(pc + u * kGroup, ps);
          v[u] = ld_stream_f1(pv + u * kGroup, ps);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv[u] = 0.0f;
        if (ok[u]) xv[u] = gx(c[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) out(i0 + u * kGroup, v[u] * xv[u]);
    }
  }
}

// ---- rows of a STREAM tile out of the product buffer: warp gw owns rows [beg, end) of the tile -------------------
// (b0, e0) are the extents of row beg+lane relative to s_prod[0] and bias0 its bias value, loaded by the caller
// before the barrier that publishes the products (their DRAM round trips overlap the stream).  One lane sums a row of up to kSerialRow products; longer rows are taken one at a time by
// the whole warp.
__device__ __forceinline__ void rows_from_products(const CsrDev& A, int r0, int n0, int beg, int end, int b0, int e0,
                                                   float bias0, const float* s_prod, int lane, float* __restrict__ y,
                                                   const Epilogue& ep) {
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    int b = b0, e = e0;
    if (base != beg) {
      b = e = 0;
      if (i < end) {
        b = A.row_ptr[r0 + i] - n0;
        e = A.row_ptr[r0 + i + 1] - n0;
      }
    }
    const int len = e - b;
    float s = 0.0f;
    // one lane per row up to kSerialRow products; the trip count is the longest such row of this pass, so a pass over
    // 10-nnz rows costs 10 steps and a pass over 1-nnz rows one
    const int mine = len <= kSerialRow ? len : 0;
    const int steps = __reduce_max_sync(kFullMask, mine);
#pragma unroll 4
    for (int k = 0; k < steps; ++k)
      if (k < mine) s += s_prod[b + k];
    unsigned big = __ballot_sync(kFullMask, len > kSerialRow);
    while (big) {
      const int j = __ffs(big) - 1;
      big &= big - 1;
      const int bj = __shfl_sync(kFullMask, b, j), ej = __shfl_sync(kFullMask, e, j);
      float p = 0.0f;
      for (int k = bj + lane; k < ej; k += 32) p += s_prod[k];
      p = warp_sum(p);
      if (lane == j) s = p;
    }
    if (i < end) {
      float v = ep.alpha * s;
      if (ep.beta != 0.0f) v = fmaf(ep.beta, base == beg ? bias0 : ep.bias[r0 + i], v);
      if (ep.relu) v = fmaxf(v, 0.0f);
      store_y(y, r0 + i, v, ep.y_mc);
    }
  }
}

// Wait for the tile's bulk copies: one thread polls the mbarrier (try_wait suspends it in hardware), the rest of the
// CTA parks on the hardware barrier.  256 threads polling the same mbarrier flood the MIO queue (measured: the
// kernel ran 1.6x slower with mio_throttle as its top stall).
__device__ __forceinline__ void tile_staged(uint64_t* bar, int cnt) {
  if (threadIdx.x == 0 && cnt > 0) mbar_wait(bar, 0);
  __syncthreads();
}

// L2 prefetch of the col/val range of tile `ta` (two bulk-prefetch instructions from one thread).  A matrix of a few
// waves of tiles is otherwise latency-bound: every wave pays descriptor + stream round trips to DRAM back to back
// while HBM idles (C3b: 33 % DRAM utilisation, long_scoreboard the top stall).
__device__ __forceinline__ void prefetch_tile_l2(const CsrDev& A, const TileDesc& da) {
  const int a = da.n0 & ~3;
  const uint32_t bytes = (uint32_t)((((da.n1 + 3) & ~3) - a) * 4);
  if (bytes == 0) return;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.col + a), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.val + a), "r"(bytes) : "memory");
}


}  // namespace

}  // namespace hispmv
