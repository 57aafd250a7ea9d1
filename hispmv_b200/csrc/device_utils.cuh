// Device-side helpers: streaming 128-bit loads with cache hints, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hispmv {

constexpr unsigned kFullMask = 0xffffffffu;

// Matrix values / column indices are read exactly once per SpMV: bypass L1 allocation and mark the
// line evict-first in L2 so the stream does not push x (the only reused operand) out of the 126 MB L2.
// On sm_100a the plain .L2::evict_first qualifier exists only on 256-bit loads (LDG.E.NA.EFL2.256); the
// narrower loads carry the same priority through an L2 cache-hint policy operand.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int4 ld_stream_i4(const int32_t* p, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int ld_stream_i1(const int32_t* p, uint64_t pol) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p, uint64_t pol) {  // 8 bytes (four 16-bit local columns)
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u16(const uint16_t* p, uint64_t pol) {
  uint16_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
  return r;
}
// 256-bit streaming load (sm_100+): 8 consecutive 32-bit words from a 32-byte aligned address.
struct alignas(32) Words8 {
  uint32_t w[8];
};
__device__ __forceinline__ Words8 ld_stream_256(const void* p) {
  Words8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                 "=r"(r.w[7])
               : "l"(p));
  return r;
}
// x gathers: read-only path, allocate in L1, prefer to keep in L2.
__device__ __forceinline__ float ld_x(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}

// Split L1 policy for x: columns below `hot` (the head of a power-law column distribution) are kept in L1
// with evict-last priority, everything else skips L1 allocation so that the cold tail cannot push the head
// out.  hot = INT_MAX keeps every gathered line (banded matrices reuse them a few rows later).
__device__ __forceinline__ float ld_x_keep(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ld_x_bypass(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ld_x_split(const float* x, int c, int hot, uint64_t pol) {
  return c < hot ? ld_x_keep(x + c, pol) : ld_x_bypass(x + c, pol);
}

// ---- TMA bulk copy (cp.async.bulk) + mbarrier -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, 16-byte aligned on both sides, size a multiple of 16; L2 evict-first (read-once stream)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// mbar_wait for consumers that may wait long: back off between polls so that the pollers do not take issue slots from
// the warp they are waiting for (the sm_100 arbiter prefers higher warp ids; a tight poll loop in 30 warps starved
// a producer in warp 0 completely: 177 M try_wait instructions per SpMV in the first version of the pipeline kernel).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(100);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
  return v;
}

template <int LANES>
__device__ __forceinline__ float subwarp_sum(float v) {
#pragma unroll
  for (int d = LANES / 2; d > 0; d >>= 1) v += __shfl_down_sync(kFullMask, v, d, LANES);
  return v;
}

__device__ __forceinline__ float finish(float ax, float alpha, float beta, const float* bias, int64_t r, int relu) {
  float v = alpha * ax;
  if (beta != 0.0f) v = fmaf(beta, bias[r], v);
  if (relu) v = fmaxf(v, 0.0f);
  return v;
}

// y[i] = v, or -- when y is the NVSwitch multicast view of a vector replicated on every GPU (Epilogue::y_mc) -- one
// multimem.st that the switch delivers to all replicas: the all-gather of a chained layer's y, fused into the kernel
// that produces it.
__device__ __forceinline__ void store_y(float* y, int64_t i, float v, int y_mc) {
  if (y_mc) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(y + i), "f"(v) : "memory");
  else y[i] = v;
}

}  // namespace hispmv
