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


}  // namespace hispmv
