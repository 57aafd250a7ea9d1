// How fast can an SM pull many SHORT contiguous runs (a few hundred bytes each, at scattered 16-byte-aligned places of a
// table far larger than L2) into shared memory?  (development microbenchmark for pass 2 of the blocked strategy: a
// panel's partial sums are ~200 such runs, one per column slab.)
//   A. one cp.async.bulk per run, issued by as many threads as there are runs, completion on one mbarrier
//   B. the same runs fetched by warps with plain loads (one run per warp step, four steps in flight) and stored to
//      shared memory -- what pb_reduce_kernel does today
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/bulk_small_bench tools/bulk_small_bench.cu
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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// run r of panel p starts at float offset 4 * (hash % (table/4 - len)): scattered like the (panel, slab) runs of `part`
__device__ __forceinline__ int64_t run_start(int64_t p, int r, int64_t table4, int len4) {
  return 4 * (int64_t)(mix64((uint64_t)p * 1315423911ull + (uint64_t)r * 2654435761ull) % (uint64_t)(table4 - len4));
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
    bulk_runs(const float* __restrict__ tab, int64_t table4, int panels, int runs, int len4, float* out) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* stage = reinterpret_cast<float*>(s_raw);
  const int tid = threadIdx.x;
  const uint32_t b = smem_u32(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  float acc = 0.f;
  const uint32_t bytes = (uint32_t)len4 * 16u;
  for (int64_t p = blockIdx.x; p < panels; p += gridDim.x) {
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * (uint32_t)runs) : "memory");
    for (int r = tid; r < runs; r += THREADS) {
      const float* src = tab + run_start(p, r, table4, len4);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(stage + (size_t)r * len4 * 4)),
                   "l"(src), "r"(bytes), "r"(b)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(b), "r"(phase)
          : "memory");
    }
    phase ^= 1;
    const int n = runs * len4 * 4;
    for (int i = tid * 4; i < n; i += THREADS * 4) {
      const float4 v = *reinterpret_cast<const float4*>(stage + i);
      acc += v.x + v.y + v.z + v.w;
    }
    __syncthreads();  // everyone has read the stage before the next panel's copies land in it
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
    warp_runs(const float* __restrict__ tab, int64_t table4, int panels, int runs, int len4, float* out) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  float* stage = reinterpret_cast<float*>(s_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = THREADS / 32;
  float acc = 0.f;
  const int len = len4 * 4;
  for (int64_t p = blockIdx.x; p < panels; p += gridDim.x) {
    for (int r0 = warp; r0 < runs; r0 += WARPS * 4) {
      float v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * WARPS;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[u][k] = 0.f;
          if (r < runs && lane + 32 * k < len)
            asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[u][k]) : "l"(tab + run_start(p, r, table4, len4) + lane + 32 * k));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * WARPS;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (r < runs && lane + 32 * k < len) stage[(size_t)r * len + lane + 32 * k] = v[u][k];
      }
    }
    __syncthreads();
    const int n = runs * len;
    for (int i = tid * 4; i < n; i += THREADS * 4) {
      const float4 q = *reinterpret_cast<const float4*>(stage + i);
      acc += q.x + q.y + q.z + q.w;
    }
    __syncthreads();
  }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int64_t table = 64ll << 20;  // floats: 256 MB, far beyond L2
  float *tab, *out;
  CK(cudaMalloc(&tab, table * 4));
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(tab, 0, table * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int panels = 4096;
  printf("# panels=%d, 512 threads, 2 CTAs per SM; run = contiguous floats at a scattered 16-byte-aligned place\n", panels);
  printf("# %-8s %6s %6s %10s %12s %10s\n", "how", "runs", "bytes", "ms", "Mruns/s", "GB/s");
  const int lens4[] = {4, 8, 12, 16, 32, 64};
  const int runs_list[] = {64, 200, 512};
  for (int how = 0; how < 2; ++how)
    for (int runs : runs_list)
      for (int len4 : lens4) {
        const size_t smem = (size_t)runs * len4 * 16;
        if (smem > 100 * 1024) continue;
        if (how == 1 && len4 > 32) continue;
        auto kern = how == 0 ? bulk_runs<512> : warp_runs<512>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaEventRecord(e0));
          kern<<<2 * sms, 512, smem>>>(tab, table / 4, panels, runs, len4, out);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        const double nruns = (double)panels * runs;
        printf("  %-8s %6d %6d %10.4f %12.1f %10.1f\n", how == 0 ? "bulk" : "warp", runs, len4 * 16, best, nruns / best * 1e-3,
               nruns * len4 * 16 / best * 1e-6);
      }
  return 0;
}
