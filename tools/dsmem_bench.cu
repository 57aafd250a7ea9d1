// Can thread-block-cluster distributed shared memory serve x gathers faster than L2?  (development microbenchmark)
// Each CTA of a cluster holds WORDS floats of a "hot window" in shared memory; every thread then gathers random
// 4-byte words from the whole cluster-wide window (ld.shared::cluster through mapa), indices streamed from global.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dsmem_bench tools/dsmem_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void make_idx(int32_t* idx, int64_t n, int window) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int32_t)(mix64(i * 77 + 5) % (uint64_t)window);
}

// mode 0: gather from the cluster-wide window (DSMEM); mode 1: same indices folded into the local CTA's smem
template <int MODE>
__global__ void gather_dsmem(const int32_t* __restrict__ idx, int64_t n4, int words_per_cta, float* out) {
  extern __shared__ float s_x[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks();
  for (int i = threadIdx.x; i < words_per_cta; i += blockDim.x) s_x[i] = 1.0f;
  cluster.sync();
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(idx) + i);
    const int cs[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (MODE == 0) {
        const unsigned rank = (unsigned)cs[k] / (unsigned)words_per_cta;
        const unsigned off = (unsigned)cs[k] % (unsigned)words_per_cta;
        const float* remote = cluster.map_shared_rank(s_x, rank % csize);
        acc += remote[off];
      } else {
        acc += s_x[(unsigned)cs[k] % (unsigned)words_per_cta];
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
  cluster.sync();
}

template <int MODE>
float run(int cluster_size, int threads, int words_per_cta, const int32_t* idx, int64_t n, float* out, int sms) {
  auto kern = gather_dsmem<MODE>;
  const size_t smem = (size_t)words_per_cta * 4;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cluster_size > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  const int grid = (sms / cluster_size) * cluster_size;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(e0));
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, idx, n / 4, words_per_cta, out);
    if (e != cudaSuccess) {
      printf("launch failed (cluster %d): %s\n", cluster_size, cudaGetErrorString(e));
      cudaGetLastError();
      return -1.f;
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int64_t n = 1ll << 27;
  int32_t* idx;
  float* out;
  CK(cudaMalloc(&idx, n * 4));
  CK(cudaMalloc(&out, 16));
  const int words = 40960;  // 160 KB per CTA
  for (int cs : {1, 2, 4, 8, 16}) {
    make_idx<<<(int)((n + 255) / 256), 256>>>(idx, n, words * cs);
    CK(cudaDeviceSynchronize());
    for (int threads : {512, 1024}) {
      const float m0 = run<0>(cs, threads, words, idx, n, out, p.multiProcessorCount);
      const float m1 = run<1>(cs, threads, words, idx, n, out, p.multiProcessorCount);
      if (m0 > 0)
        printf("cluster=%2d threads=%4d window=%6.2f MB : DSMEM gather %7.2f G/s (%.3f ms) | local-smem gather %7.2f G/s (%.3f ms)\n",
               cs, threads, words * cs * 4 / 1e6, n / m0 / 1e6, m0, n / m1 / 1e6, m1);
    }
  }
  return 0;
}
