// Ceilings that bound the SpMV kernels on this GPU, measured directly (development tool):
//   1. streaming read bandwidth for 128-bit / 256-bit evict-first loads
//   2. x-gather throughput (4-byte loads through L1/L2) as a function of table size and index locality
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int MODE>  // 0: ld.global.nc.v4 ; 1: 256-bit evict_first ; 2: v4 + L1::no_allocate + evict_first policy
__global__ void stream_kernel(const float* __restrict__ a, int64_t n4, float* out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (MODE == 1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4 / 2; i += stride) {
      uint32_t w[8];
      asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(a + 8 * i));
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += __uint_as_float(w[k]);
    }
  } else {
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 v;
      if (MODE == 0) {
        v = __ldg(reinterpret_cast<const float4*>(a) + i);
      } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(a + 4 * i), "l"(pol));
      }
      acc += v.x + v.y + v.z + v.w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// index patterns
//  0 uniform random            1 power-law q^5 (Zipf s=0.8, hubs at low ids)
//  2 sequential                3 random base per 8 consecutive entries, +0..7 (short sorted runs)
//  4 power-law with hubs scattered by an affine permutation
__global__ void make_idx_kernel(int32_t* idx, int64_t n, int64_t table, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = mix64((uint64_t)i * 0x9E3779B97F4A7C15ull + 12345);
  int64_t v;
  const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  if (mode == 0) v = (int64_t)(u * table);
  else if (mode == 1 || mode == 4) {
    const double q = u * u * u * u * u;
    v = (int64_t)(q * table);
    if (mode == 4) v = (int64_t)(((unsigned __int128)v * 2654435761ull + 977) % (uint64_t)table);
  } else if (mode == 2) v = i % table;
  else {
    const uint64_t hb = mix64((uint64_t)(i / 8) + 777);
    v = (int64_t)((double)(hb >> 11) * (1.0 / 9007199254740992.0) * (table - 8)) + (i % 8);
  }
  if (v >= table) v = table - 1;
  idx[i] = (int32_t)v;
}

// each thread: int4 of indices (streamed), 4 gathers, UNROLL of those in flight
template <int UNROLL>
__global__ void gather_kernel(const int32_t* __restrict__ idx, const float* __restrict__ table, int64_t n4, float* out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * UNROLL) {
    int4 c[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = i + u * stride;
      if (j < n4) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(c[u].x), "=r"(c[u].y), "=r"(c[u].z), "=r"(c[u].w)
                     : "l"(idx + 4 * j), "l"(pol));
      } else {
        c[u] = make_int4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      acc += __ldg(table + c[u].x) + __ldg(table + c[u].y) + __ldg(table + c[u].z) + __ldg(table + c[u].w);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_ms(F f, int iters) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, L2 %d MB, clock %d MHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20,
         p.clockRate / 1000);
  float* out;
  CK(cudaMalloc(&out, 16));
  // ---- 1. streaming ------------------------------------------------------------------------------
  {
    const int64_t n = 1ll << 29;  // 2 GiB of floats
    float* a;
    CK(cudaMalloc(&a, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    for (int ctas = 2; ctas <= 16; ctas *= 2) {
      for (int threads = 256; threads <= 1024; threads *= 2) {
        if (ctas * threads > 2048) continue;
        const int grid = p.multiProcessorCount * ctas;
        float m0 = time_ms([&] { stream_kernel<0><<<grid, threads>>>(a, n / 4, out); }, 5);
        float m1 = time_ms([&] { stream_kernel<1><<<grid, threads>>>(a, n / 4, out); }, 5);
        float m2 = time_ms([&] { stream_kernel<2><<<grid, threads>>>(a, n / 4, out); }, 5);
        printf("stream ctas/SM=%2d threads=%4d : ldg128 %7.1f GB/s | ld256.EF %7.1f GB/s | ld128.NA.EF %7.1f GB/s\n", ctas,
               threads, n * 4 / m0 / 1e6, n * 4 / m1 / 1e6, n * 4 / m2 / 1e6);
      }
    }
    CK(cudaFree(a));
  }
  // ---- 2. gathers --------------------------------------------------------------------------------
  {
    const int64_t n = 1ll << 28;  // 268M gathers, 1 GiB of indices
    int32_t* idx;
    CK(cudaMalloc(&idx, n * 4));
    const int64_t tables[] = {1ll << 13, 1ll << 14, 1ll << 15, 1ll << 16, 1ll << 18, 10000000ll, 100000000ll};  // elements
    const char* names[] = {"uniform", "zipf0.8", "sequential", "runs-of-8", "zipf0.8-scattered"};
    for (int64_t table : tables) {
      float* t;
      CK(cudaMalloc(&t, table * 4));
      CK(cudaMemset(t, 0, table * 4));
      for (int mode = 0; mode < 5; ++mode) {
        make_idx_kernel<<<(int)((n + 255) / 256), 256>>>(idx, n, table, mode);
        CK(cudaDeviceSynchronize());
        const int grid = p.multiProcessorCount * 8;
        float m2 = time_ms([&] { gather_kernel<2><<<grid, 256>>>(idx, t, n / 4, out); }, 3);
        float m4 = time_ms([&] { gather_kernel<4><<<grid, 256>>>(idx, t, n / 4, out); }, 3);
        const float m = m2 < m4 ? m2 : m4;
        // an SpMV nonzero = 8 streamed bytes + one gather; here 4 streamed bytes + one gather
        printf("gather table=%6.1f MB %-18s: %7.2f Ggather/s (u2 %.3f ms, u4 %.3f ms)  => SpMV-equivalent ceiling %7.1f GB/s\n",
               table * 4 / 1e6, names[mode], n / m / 1e6, m2, m4, 8.0 * n / m / 1e6);
      }
      CK(cudaFree(t));
    }
    CK(cudaFree(idx));
  }
  return 0;
}
