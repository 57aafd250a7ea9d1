// Research kernels, kept for the record: measured, tested, never selected by the planner.  Built
// only with `make EXPERIMENTAL=1` (-DHISPMV_EXPERIMENTAL); the default library carries stubs that refuse.
//   spmv_warptile_kernel             one warp per (small) tile, no CTA barrier                HISPMV_WARPTILE=B,T,CH
//   spmv_adaptive_persistent_kernel  resident CTAs, x[0, hot) in shared memory               HISPMV_PERSIST=cols
//   spmv_pipeline_kernel             warp-specialised TMA producer -> mbarrier ring -> gather/reduce teams  HISPMV_PIPELINE=1
#include <limits.h>

#include <algorithm>

#include "device_utils.cuh"
#include "internal.h"
#include "tile_device.cuh"

namespace hispmv {

#ifndef HISPMV_EXPERIMENTAL

static int refuse(const char* what) {
  set_error(std::string(what) + ": this library was built without the research kernels (make EXPERIMENTAL=1)");
  return HISPMV_ERR_STATE;
}
int launch_warptile(const CsrDev&, const AdaptivePlan&, const float*, float*, Epilogue, cudaStream_t) {
  return refuse("warp-per-tile kernel");
}
int launch_adaptive_persistent(const CsrDev&, const AdaptivePlan&, const float*, float*, Epilogue, int, cudaStream_t) {
  return refuse("persistent x-window kernel");
}
int launch_pipeline(const CsrDev&, const AdaptivePlan&, const float*, float*, Epilogue, int, cudaStream_t) {
  return refuse("pipeline kernel");
}
bool experimental_kernels_built() { return false; }

#else

bool experimental_kernels_built() { return true; }

namespace {

// ================================================================================================================
// one WARP per tile
// ================================================================================================================
// The same tiles at warp granularity (a few hundred items): a warp loads its tile's col/val, gathers, keeps the
// products in its own slice of shared memory and reduces its rows -- no CTA-wide barrier anywhere, so no warp ever
// waits for another warp's gathers, and an SM has 64 independent latency chains in flight instead of 8.
template <int WCAP, bool SPLIT>  // WCAP >= stream_items + long_threshold and >= chunk_nnz
__global__ void __launch_bounds__(kGroup, 8)
    spmv_warptile_kernel(CsrDev A, AdaptivePlan P, const float* __restrict__ x, float* __restrict__ y, Epilogue ep) {
  __shared__ float s_prod_all[kGroupWarps][WCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = P.tile_begin + (int64_t)blockIdx.x * kGroupWarps + warp;
  const int64_t t_end = P.tile_begin + (P.tile_count >= 0 ? P.tile_count : P.num_tiles);
  if (t >= t_end) return;
  float* s_prod = s_prod_all[warp];
  const TileDesc d = load_desc(P.desc + t);
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const GatherL1<SPLIT> gx{x, P.hot_cols, pk};
  if (d.chunk >= 0) {
    float acc = 0.0f;
    stream_products<32>(A, gx, d.n0, d.n1, lane, ps, [&](int, float p) { acc += p; });
    acc = warp_sum(acc);
    finish_chunk(P.carry, P.counter, d, t, acc, lane, y, ep);
    return;
  }
  const int n0 = d.n0;
  const int trows = d.r1 - d.r0;
  int b0 = 0, e0 = 0;
  float bias0 = 0.0f;
  if (lane < trows) {
    b0 = A.row_ptr[d.r0 + lane];
    e0 = A.row_ptr[d.r0 + lane + 1];
    if (ep.beta != 0.0f) bias0 = ep.bias[d.r0 + lane];
  }
  stream_products<32>(A, gx, n0, d.n1, lane, ps, [&](int i, float p) { s_prod[i - n0] = p; });
  __syncwarp();
  rows_from_products(A, d.r0, n0, 0, trows, b0 - n0, e0 - n0, bias0, s_prod, lane, y, ep);
}

// ================================================================================================================
// persistent, x window in shared memory
// ================================================================================================================
constexpr int kPsGroups = 4;

template <int CAP>
struct PsGroupSmem {
  float prod[CAP];
  float red[kGroupWarps];
  __align__(16) TileDesc ring[3];  // descriptors of the current tile and the two after it
};

// named barrier g+1 over the kGroup threads of group g; immediate ids so that ptxas reserves only 5 of the SM's 16
// hardware barriers per CTA (a register id makes it reserve all 16, which caps residency at one CTA per SM)
__device__ __forceinline__ void group_sync(int g) {
  switch (g) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(kGroup) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(kGroup) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(kGroup) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(kGroup) : "memory"); break;
  }
}
static_assert(kPsGroups == 4, "group_sync names four barriers");

__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// descriptor of tile t -> shared memory, asynchronously (two 16-byte cp.async); past the end: an end marker
__device__ __forceinline__ void desc_async(const AdaptivePlan& P, unsigned int t, TileDesc* dst) {
  if ((int64_t)t < P.num_tiles) {
    const uint32_t d = smem_u32(dst);
    const TileDesc* src = P.desc + t;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(reinterpret_cast<const char*>(src) + 16)
                 : "memory");
  } else {
    dst->tile = -1;
    dst->chunk = -1;
    dst->n0 = dst->n1 = 0;
  }
}
__device__ __forceinline__ void desc_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int CAP, int MINBLOCKS>
__global__ void __launch_bounds__(kGroup* kPsGroups, MINBLOCKS)
    spmv_adaptive_persistent_kernel(CsrDev A, AdaptivePlan P, const float* __restrict__ x, float* __restrict__ y,
                                    Epilogue ep, int total_groups) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_x = reinterpret_cast<float*>(s_raw);
  const int hot = P.hot_cols;
  const int tid = threadIdx.x, g = tid / kGroup, gt = tid % kGroup, lane = tid & 31, gw = gt >> 5;
  PsGroupSmem<CAP>* gs = reinterpret_cast<PsGroupSmem<CAP>*>(s_raw + (((size_t)hot * 4 + 15) & ~(size_t)15)) + g;
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const GatherWindow gx{x, s_x, hot, pk};

  // Thread 0 of each group runs the tile pipeline, one step per processed tile k:
  //   tile index k+3   atomicAdd on the global counter (result used one step later)
  //   descriptor k+2   cp.async into the group's ring (lands during tile k)
  //   stream of k+1    cp.async.bulk.prefetch.L2 of its col/val range, so the loads of the next tile hit L2
  unsigned int t_ahead = 0;
  if (gt == 0) {
    const unsigned int t0 = atomicAdd(P.sched, 1u);
    const unsigned int t1 = atomicAdd(P.sched, 1u);
    t_ahead = atomicAdd(P.sched, 1u);
    desc_async(P, t0, &gs->ring[0]);
    desc_async(P, t1, &gs->ring[1]);
  }
  for (int i = tid; i < hot; i += kGroup * kPsGroups) s_x[i] = ld_x_bypass(x + i, pk);
  __syncthreads();

  for (int k = 0;; ++k) {
    if (gt == 0) desc_wait();
    group_sync(g);
    const TileDesc d = gs->ring[k % 3];
    if (d.tile < 0) break;
    if (gt == 0) {
      const TileDesc* dn = &gs->ring[(k + 1) % 3];
      if (dn->tile >= 0 && dn->n1 > dn->n0) {
        const int a0 = dn->n0 & ~3;
        const uint32_t bytes = (uint32_t)(((dn->n1 + 3) & ~3) - a0) * 4u;
        prefetch_l2(A.col + a0, bytes);
        prefetch_l2(A.val + a0, bytes);
      }
      desc_async(P, t_ahead, &gs->ring[(k + 2) % 3]);
      t_ahead = atomicAdd(P.sched, 1u);
    }
    const int64_t t = d.tile;
    if (d.chunk >= 0) {
      float acc = 0.0f;
      stream_products(A, gx, d.n0, d.n1, gt, ps, [&](int, float p) { acc += p; });
      acc = warp_sum(acc);
      if (lane == 0) gs->red[gw] = acc;
      group_sync(g);
      if (gw == 0) {
        float total = lane < kGroupWarps ? gs->red[lane] : 0.0f;
        total = warp_sum(total);
        finish_chunk(P.carry, P.counter, d, t, total, lane, y, ep);
      }
      continue;
    }
    const int n0 = d.n0;
    float* s_prod = gs->prod;
    stream_products(A, gx, n0, d.n1, gt, ps, [&](int i, float p) { s_prod[i - n0] = p; });
    const int trows = d.r1 - d.r0;
    const int rpw = (trows + kGroupWarps - 1) / kGroupWarps;
    const int beg = gw * rpw, end = min(trows, beg + rpw);
    int b0 = 0, e0 = 0;
    float bias0 = 0.0f;
    if (beg + lane < end) {
      b0 = A.row_ptr[d.r0 + beg + lane] - n0;
      e0 = A.row_ptr[d.r0 + beg + lane + 1] - n0;
      if (ep.beta != 0.0f) bias0 = ep.bias[d.r0 + beg + lane];
    }
    group_sync(g);
    rows_from_products(A, d.r0, n0, beg, end, b0, e0, bias0, s_prod, lane, y, ep);
  }
  // the last group to run dry re-arms the tile counter for the next launch / graph replay
  if (gt == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(P.sched + 1, 1u + (t_ahead == 0xffffffffu));  // t_ahead has landed
    if (prev == (unsigned int)(total_groups - 1)) {
      P.sched[0] = 0;
      P.sched[1] = 0;
    }
  }
}

// ================================================================================================================
// warp-specialised persistent pipeline (one CTA of 992 threads per SM)
// ================================================================================================================
// The one-CTA-per-tile kernel serialises, inside every CTA, a DRAM round trip (col/val), an L2 round trip (gathers),
// a barrier and the row reduction (with two more DRAM round trips for row extents and bias); its ncu profile shows
// the SM's L1-miss request path -- the real ceiling on gather-heavy matrices -- busy only ~65-70 % of the time.
// Here a producer warp keeps a ring of tile-sized stages full by TMA, so that everything a tile needs is already in
// shared memory when a team of warps picks it up:
//   warp 30         producer.  CTA b owns tiles b, b+G, b+2G, ... (strided: statistically balanced, deterministic,
//                   no atomics).  Its 32 lanes fetch 32 descriptors at a time; lane 0 then waits for a free stage and
//                   fills it with up to four bulk copies: col, val, the tile's row_ptr slice and its bias slice.
//   teams           kPipeTeams x kPipeTeamWarps warps; a team takes every kPipeTeams-th stage: columns from shared
//                   memory, U gathers in flight per thread, product written over val; a named barrier over the team;
//                   then each warp reduces a slice of the tile's rows out of shared memory (LONG chunks: per-warp
//                   partials, warp 0 finishes the chunk) and the stage goes back to the producer.  While one team
//                   reduces, the others gather, so the miss path always has a burst in flight.
// (A first version with separate reduce warps was reduce-bound: 7 warps need ~3 k cycles for a tile of short rows.)
// full[s]  (1 arrival + TMA bytes)  producer -> team          empty[s] (team warps)  team -> producer
constexpr int kPipeTeams = 3;
constexpr int kPipeTeamWarps = 10;
constexpr int kPipeTeamThreads = kPipeTeamWarps * 32;
constexpr int kPipeProducerWarp = kPipeTeams * kPipeTeamWarps;  // warp 30: the arbiter favours high warp ids
constexpr int kPipeThreads = (kPipeProducerWarp + 1) * 32;

#ifdef HISPMV_DIAG
__device__ long long g_pipe_dbg[8 * 512];  // [event][tile k] clock64 stamps of CTA 0
#define PIPE_STAMP(ev, k) do { if (blockIdx.x == 0 && (k) < 512) g_pipe_dbg[(ev) * 512 + (k)] = clock64(); } while (0)
#else
#define PIPE_STAMP(ev, k) do { } while (0)
#endif

template <int CAP, int RCAP>
struct PipeStage {
  __align__(128) int col[CAP + 8];
  __align__(16) float val[CAP + 8];
  __align__(16) int rp[RCAP + 8];      // row_ptr[r0 & ~3 ...]
  __align__(16) float bias[RCAP + 8];  // bias[r0 & ~3 ...]
  __align__(16) TileDesc desc;
  float partial[kPipeTeamWarps];
};
template <int STAGES>
struct PipeBars {
  uint64_t full[STAGES], empty[STAGES];
};

// one lane polls, the warp follows: keeps hundreds of threads from hammering the same mbarrier
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait_backoff(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void team_sync(int team) {
  switch (team) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(kPipeTeamThreads) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(kPipeTeamThreads) : "memory"); break;
    default: asm volatile("bar.sync 3, %0;" ::"n"(kPipeTeamThreads) : "memory"); break;
  }
}
static_assert(kPipeTeams == 3, "team_sync names three barriers");

__device__ __forceinline__ TileDesc desc_or_end(const AdaptivePlan& P, int64_t t) {
  TileDesc d;
  if (t < P.num_tiles) {
    d = load_desc(P.desc + t);
  } else {
    d.r0 = d.r1 = d.n0 = d.n1 = d.nchunks = d.pad = 0;
    d.chunk = -1;
    d.tile = -1;  // end marker
  }
  return d;
}
__device__ __forceinline__ TileDesc shfl_desc(const TileDesc& d, int src) {
  TileDesc o;
  o.r0 = __shfl_sync(kFullMask, d.r0, src);
  o.r1 = __shfl_sync(kFullMask, d.r1, src);
  o.n0 = __shfl_sync(kFullMask, d.n0, src);
  o.n1 = __shfl_sync(kFullMask, d.n1, src);
  o.chunk = __shfl_sync(kFullMask, d.chunk, src);
  o.nchunks = __shfl_sync(kFullMask, d.nchunks, src);
  o.tile = __shfl_sync(kFullMask, d.tile, src);
  o.pad = 0;
  return o;
}

template <int CAP, int RCAP, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(kPipeThreads, 1)
    spmv_pipeline_kernel(CsrDev A, AdaptivePlan P, const float* __restrict__ x, float* __restrict__ y, Epilogue ep) {
  using Stage = PipeStage<CAP, RCAP>;
  extern __shared__ __align__(128) unsigned char s_raw[];
  Stage* stages = reinterpret_cast<Stage*>(s_raw);
  PipeBars<STAGES>* bars = reinterpret_cast<PipeBars<STAGES>*>(s_raw + sizeof(Stage) * STAGES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // bias travels by TMA when it is 16-byte aligned (always, for cudaMalloc'ed vectors); only whole 4-float groups
  // inside the vector are copied, a ragged last group is read directly
  const bool use_bias = ep.beta != 0.0f;
  const bool bias_tma = use_bias && (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0;
  const int bias_full = A.rows & ~3;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kPipeTeamWarps);
    }
  }
  __syncthreads();

  if (warp == kPipeProducerWarp) {
    // ------------------------------------------------------------------------------------------- producer
    const uint64_t ps = policy_evict_first(), pn = policy_evict_normal();
    int k = 0, ends = 0;
    for (int64_t kb = 0; ends < kPipeTeams; kb += 32) {
      const TileDesc mine = desc_or_end(P, (int64_t)blockIdx.x + (kb + lane) * (int64_t)gridDim.x);
      for (int j = 0; j < 32 && ends < kPipeTeams; ++j, ++k) {
        const TileDesc d = shfl_desc(mine, j);
        const int s = k % STAGES, u = k / STAGES;
        Stage& st = stages[s];
        // lane 0 owns the barriers; the four copies of a tile are issued by four different lanes (issuing one
        // cp.async.bulk costs the issuing thread several hundred cycles: one lane doing all four was the bottleneck)
        if (lane == 0) {
          PIPE_STAMP(0, k);
          if (u > 0) mbar_wait(&bars->empty[s], (u - 1) & 1);
          PIPE_STAMP(1, k);
          st.desc = d;
        }
        if (d.tile < 0) {  // one end marker per team
          if (lane == 0) mbar_arrive(&bars->full[s]);
          ++ends;
          continue;
        }
        const int a0 = d.n0 & ~3;
        const int cnt = ((d.n1 + 3) & ~3) - a0;
        uint32_t bytes = (uint32_t)cnt * 8u;
        int ra = 0, rcnt = 0, bcnt = 0;
        if (d.chunk < 0) {
          ra = d.r0 & ~3;
          rcnt = ((d.r1 + 1 + 3) & ~3) - ra;  // row_ptr[ra, ra + rcnt): the allocation is padded by 4 entries
          bytes += (uint32_t)rcnt * 4u;
          if (bias_tma) {
            bcnt = min((d.r1 + 3) & ~3, bias_full) - ra;
            if (bcnt < 0) bcnt = 0;
            bytes += (uint32_t)bcnt * 4u;
          }
        }
        if (lane == 0) {
          if (bytes > 0) mbar_expect_tx(&bars->full[s], bytes);
          else mbar_arrive(&bars->full[s]);
        }
        __syncwarp();
        if (lane == 0 && cnt > 0) bulk_g2s_hint(st.col, A.col + a0, (uint32_t)cnt * 4u, &bars->full[s], ps);
        if (lane == 1 && cnt > 0) bulk_g2s_hint(st.val, A.val + a0, (uint32_t)cnt * 4u, &bars->full[s], ps);
        if (lane == 2 && rcnt > 0) bulk_g2s_hint(st.rp, A.row_ptr + ra, (uint32_t)rcnt * 4u, &bars->full[s], pn);
        if (lane == 3 && bcnt > 0) bulk_g2s_hint(st.bias, ep.bias + ra, (uint32_t)bcnt * 4u, &bars->full[s], pn);
      }
    }
    return;
  }

  // ----------------------------------------------------------------------------------------------- teams
  const uint64_t pk = policy_evict_last();
  const GatherL1<SPLIT> gx{x, P.hot_cols, pk};
  const int team = warp / kPipeTeamWarps, tw = warp % kPipeTeamWarps;
  const int tt = tw * 32 + lane;
  constexpr int U = CAP / kPipeTeamThreads;  // gathers in flight per thread
  static_assert(CAP % kPipeTeamThreads == 0, "tile capacity must be a multiple of the team size");
  for (int k = team;; k += kPipeTeams) {
    const int s = k % STAGES, u = k / STAGES;
    Stage& st = stages[s];
    warp_wait(&bars->full[s], u & 1, lane);
    if (tw == 0 && lane == 0) PIPE_STAMP(2, k);
    const TileDesc d = st.desc;
    if (d.tile < 0) break;
    const int a0 = d.n0 & ~3;
    const int k0 = d.n0 - a0, k1 = k0 + (d.n1 - d.n0);
    int c[U];
    float xv[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int i = k0 + tt + q * kPipeTeamThreads;
      c[q] = i < k1 ? st.col[i] : -1;
    }
#pragma unroll
    for (int q = 0; q < U; ++q) xv[q] = c[q] >= 0 ? gx(c[q]) : 0.0f;
    if (d.chunk >= 0) {
      float acc = 0.0f;
#pragma unroll
      for (int q = 0; q < U; ++q)
        if (c[q] >= 0) acc = fmaf(st.val[k0 + tt + q * kPipeTeamThreads], xv[q], acc);
      acc = warp_sum(acc);
      if (lane == 0) st.partial[tw] = acc;
      team_sync(team);
      if (tw == 0) {
        float total = lane < kPipeTeamWarps ? st.partial[lane] : 0.0f;
        total = warp_sum(total);
        finish_chunk(P.carry, P.counter, d, d.tile, total, lane, y, ep);
      }
    } else {
#pragma unroll
      for (int q = 0; q < U; ++q)
        if (c[q] >= 0) st.val[k0 + tt + q * kPipeTeamThreads] *= xv[q];
      team_sync(team);
      if (tw == 0 && lane == 0) PIPE_STAMP(3, k);
      const int ra = d.r0 & ~3;
      const int trows = d.r1 - d.r0;
      const int rpw = (trows + kPipeTeamWarps - 1) / kPipeTeamWarps;
      const int beg = tw * rpw, end = min(trows, beg + rpw);
      const int* rp = st.rp + (d.r0 - ra);      // rp[i] = row_ptr[r0 + i]
      const float* bs = st.bias + (d.r0 - ra);  // bs[i] = bias[r0 + i] (where copied)
      const float* prod = st.val;               // product of nonzero n at prod[n - a0]
      for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        int b = 0, e = 0;
        if (i < end) {
          b = rp[i] - a0;
          e = rp[i + 1] - a0;
        }
        const int len = e - b;
        float sum = 0.0f;
        const int mine = len <= kSerialRow ? len : 0;
        const int steps = __reduce_max_sync(kFullMask, mine);
#pragma unroll 4
        for (int q = 0; q < steps; ++q)
          if (q < mine) sum += prod[b + q];
        unsigned big = __ballot_sync(kFullMask, len > kSerialRow);
        while (big) {
          const int j = __ffs(big) - 1;
          big &= big - 1;
          const int bj = __shfl_sync(kFullMask, b, j), ej = __shfl_sync(kFullMask, e, j);
          float p = 0.0f;
          for (int q = bj + lane; q < ej; q += 32) p += prod[q];
          p = warp_sum(p);
          if (lane == j) sum = p;
        }
        if (i < end) {
          float v = ep.alpha * sum;
          if (use_bias) {
            const int r = d.r0 + i;
            v = fmaf(ep.beta, (bias_tma && r < bias_full) ? bs[i] : ep.bias[r], v);
          }
          if (ep.relu) v = fmaxf(v, 0.0f);
          y[d.r0 + i] = v;
        }
      }
    }
    __syncwarp();
    if (tw == 0 && lane == 0) PIPE_STAMP(5, k);
    if (lane == 0) mbar_arrive(&bars->empty[s]);
  }
}

#ifdef HISPMV_DIAG
}  // namespace
}  // namespace hispmv
extern "C" int hispmv_debug_pipe(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, hispmv::g_pipe_dbg, sizeof(long long) * 8 * 512);
}
namespace hispmv {
namespace {
#endif


// ---- launch helpers ----------------------------------------------------------------------------------------------
template <int CAP, int MINBLOCKS>
int launch_persistent_inst(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, int sm_count,
                           size_t smem, cudaStream_t s) {
  auto k = spmv_adaptive_persistent_kernel<CAP, MINBLOCKS>;
  HISPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = (P.num_tiles + kPsGroups - 1) / kPsGroups;
  if (grid > (int64_t)sm_count * MINBLOCKS) grid = (int64_t)sm_count * MINBLOCKS;
  k<<<(int)grid, kGroup * kPsGroups, smem, s>>>(A, P, x, y, ep, (int)grid * kPsGroups);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}
template <int CAP>
int launch_persistent_cap(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, int sm_count,
                          cudaStream_t s) {
  const size_t smem = (((size_t)P.hot_cols * 4 + 15) & ~(size_t)15) + sizeof(PsGroupSmem<CAP>) * kPsGroups;
  // small windows leave room for two resident CTAs per SM (eight tile groups instead of four)
  if (smem * 2 <= (size_t)225 * 1024) return launch_persistent_inst<CAP, 2>(A, P, x, y, ep, sm_count, smem, s);
  if (smem <= (size_t)226 * 1024) return launch_persistent_inst<CAP, 1>(A, P, x, y, ep, sm_count, smem, s);
  set_error("adaptive_persistent: x window does not fit in shared memory");
  return HISPMV_ERR_ARG;
}


}  // namespace

int launch_warptile(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, cudaStream_t s) {
  if (A.rows <= 0 || P.num_tiles <= 0) return HISPMV_OK;
  int st = check_plan(P, kWarpTileCap, "warptile");
  if (st != HISPMV_OK) return st;
  if (P.chunk_nnz > 65536) {
    set_error("warptile: chunk_nnz too large");
    return HISPMV_ERR_ARG;
  }
  const int64_t tiles = P.tile_count >= 0 ? P.tile_count : P.num_tiles;
  const int grid = (int)((tiles + kGroupWarps - 1) / kGroupWarps);
  if (grid <= 0) return HISPMV_OK;
  if (P.hot_cols != 0x7fffffff) spmv_warptile_kernel<kWarpTileCap, true><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
  else spmv_warptile_kernel<kWarpTileCap, false><<<grid, kGroup, 0, s>>>(A, P, x, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int launch_adaptive_persistent(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep,
                               int sm_count, cudaStream_t s) {
  if (A.rows <= 0 || P.num_tiles <= 0) return HISPMV_OK;
  int st = check_plan(P, 4096, "adaptive_persistent");
  if (st != HISPMV_OK) return st;
  if (P.hot_cols < 0 || P.hot_cols > A.cols || !P.sched) {
    set_error("adaptive_persistent: bad x window");
    return HISPMV_ERR_ARG;
  }
  const int need = P.stream_items + P.long_threshold;
  if (need <= 2048) return launch_persistent_cap<2048>(A, P, x, y, ep, sm_count, s);
  if (need <= 3072) return launch_persistent_cap<3072>(A, P, x, y, ep, sm_count, s);
  return launch_persistent_cap<4096>(A, P, x, y, ep, sm_count, s);
}

namespace {
template <int CAP, int RCAP, int STAGES, bool SPLIT>
int launch_pipeline_inst(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, int sm_count,
                         cudaStream_t s) {
  auto k = spmv_pipeline_kernel<CAP, RCAP, STAGES, SPLIT>;
  const size_t smem = sizeof(PipeStage<CAP, RCAP>) * STAGES + sizeof(PipeBars<STAGES>);
  HISPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(P.num_tiles, sm_count);
  k<<<grid, kPipeThreads, smem, s>>>(A, P, x, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}
}  // namespace

int launch_pipeline(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, int sm_count,
                    cudaStream_t s) {
  if (A.rows <= 0 || P.num_tiles <= 0) return HISPMV_OK;
  int st = check_plan(P, kPipelineCap, "pipeline");
  if (st != HISPMV_OK) return st;
  if (P.chunk_nnz > kPipelineCap || P.stream_items > kPipelineRows) {
    set_error("pipeline: tiles must fit a stage (chunk_nnz <= 2048, stream_items <= 1536)");
    return HISPMV_ERR_ARG;
  }
  // CAP nonzeros + RCAP row extents per stage: 26.6 KB, five stages = 133 KB (shared memory beyond ~190 KB per SM
  // throttles the L1-miss path)
  if (P.hot_cols != 0x7fffffff)
    return launch_pipeline_inst<kPipelineCap, kPipelineRows + 8, 5, true>(A, P, x, y, ep, sm_count, s);
  return launch_pipeline_inst<kPipelineCap, kPipelineRows + 8, 5, false>(A, P, x, y, ep, sm_count, s);
}


#endif  // HISPMV_EXPERIMENTAL

}  // namespace hispmv
