// x replication over NVSwitch multicast (NVLS): the rank that owns x stores it once to a multicast address and the
// switch delivers the stores to every GPU of the group -- no SM on the receiving side runs anything, and the sender's
// NVLink egress carries x once instead of once per peer.  The multicast mapping itself comes from
// torch.distributed._symmetric_memory (device memory + rendezvous are plumbing); the store kernel is ours.
#include <cstdlib>

#include "device_utils.cuh"
#include "internal.h"

namespace hispmv {
namespace {

__device__ __forceinline__ void mc_store4(float* dst, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(512) multicast_copy_kernel(float* __restrict__ mc_dst, const float* __restrict__ src,
                                                             int64_t n4, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
    const float4 a = __ldg(s4 + i), b = __ldg(s4 + i + stride), c = __ldg(s4 + i + 2 * stride), d = __ldg(s4 + i + 3 * stride);
    mc_store4(mc_dst + 4 * i, a);
    mc_store4(mc_dst + 4 * (i + stride), b);
    mc_store4(mc_dst + 4 * (i + 2 * stride), c);
    mc_store4(mc_dst + 4 * (i + 3 * stride), d);
  }
  for (; i < n4; i += stride) mc_store4(mc_dst + 4 * i, __ldg(s4 + i));
  // tail (n not a multiple of 4)
  const int64_t t = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(mc_dst + t), "f"(src[t]) : "memory");
  __threadfence_system();
}

// The same block of x stored into the replica of every peer with plain 128-bit stores over NVLink (peer pointers from
// the symmetric-memory rendezvous): CTA c serves peer c % n_peers.  Unicast stores are not held to the switch's
// multicast rate (~400-450 GB/s into a GPU): with every rank sending 1/N of x to N-1 peers the links carry
// (N-1)/N of x in and out of each GPU concurrently.
struct PeerPtrs {
  float* p[16];
};
__global__ void __launch_bounds__(512) peer_copy_kernel(PeerPtrs dst, int n_peers, const float* __restrict__ src, int64_t n4,
                                                        int64_t n) {
  const int peer = blockIdx.x % n_peers;
  const int64_t lane_cta = blockIdx.x / n_peers, ctas = (gridDim.x - peer + n_peers - 1) / n_peers;
  float* __restrict__ d = dst.p[peer];
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(d);
  const int64_t stride = ctas * blockDim.x;
  int64_t i = lane_cta * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a = __ldg(s4 + i), b = __ldg(s4 + i + stride), c = __ldg(s4 + i + 2 * stride), e = __ldg(s4 + i + 3 * stride);
    d4[i] = a;
    d4[i + stride] = b;
    d4[i + 2 * stride] = c;
    d4[i + 3 * stride] = e;
  }
  for (; i < n4; i += stride) d4[i] = __ldg(s4 + i);
  const int64_t t = 4 * n4 + lane_cta * blockDim.x + threadIdx.x;
  if (t < n) d[t] = src[t];
  __threadfence_system();
}

}  // namespace
}  // namespace hispmv

using namespace hispmv;

extern "C" int hispmv_peer_copy(void* const* peer_dst, int n_peers, const float* d_src, int64_t n, int ctas_per_peer,
                                void* stream) {
  if (!peer_dst || n_peers < 1 || n_peers > 16 || !d_src || n < 0 || (reinterpret_cast<uintptr_t>(d_src) & 15)) {
    set_error("peer_copy: 1..16 peers, 16-byte aligned pointers");
    return HISPMV_ERR_ARG;
  }
  PeerPtrs pp;
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_dst[i] || (reinterpret_cast<uintptr_t>(peer_dst[i]) & 15)) {
      set_error("peer_copy: 1..16 peers, 16-byte aligned pointers");
      return HISPMV_ERR_ARG;
    }
    pp.p[i] = static_cast<float*>(peer_dst[i]);
  }
  if (n == 0) return HISPMV_OK;
  const int grid = n_peers * (ctas_per_peer > 0 ? ctas_per_peer : 2);
  peer_copy_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(pp, n_peers, d_src, n / 4, n);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

extern "C" int hispmv_multicast_copy(void* mc_dst, const float* d_src, int64_t n, int sm_budget, void* stream) {
  if (!mc_dst || !d_src || n < 0 || (reinterpret_cast<uintptr_t>(mc_dst) & 15) || (reinterpret_cast<uintptr_t>(d_src) & 15)) {
    set_error("multicast_copy: pointers must be 16-byte aligned");
    return HISPMV_ERR_ARG;
  }
  if (n == 0) return HISPMV_OK;
  if (sm_budget < 0) {  // copy engine writing to the multicast address: no SM at all
    HISPMV_CUDA(cudaMemcpyAsync(mc_dst, d_src, (size_t)n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HISPMV_OK;
  }
  const int grid = sm_budget > 0 ? sm_budget : 32;
  // threads per CTA: 512 by default; small CTAs (HISPMV_MC_THREADS=128 with one CTA per SM) fit beside kernels that
  // leave few registers free, so every SM carries the same small share of the copy
  static const int forced = getenv("HISPMV_MC_THREADS") ? atoi(getenv("HISPMV_MC_THREADS")) : 0;
  const int threads = (forced == 64 || forced == 128 || forced == 256) ? forced : 512;
  multicast_copy_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(static_cast<float*>(mc_dst), d_src, n / 4, n);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}
