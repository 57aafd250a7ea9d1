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
//   research switches, tested but never auto-selected: spmv_warptile_kernel (one warp per tile),
//   spmv_adaptive_persistent_kernel (resident CTAs, x window in shared memory), spmv_pipeline_kernel
//   (warp-specialised TMA producer -> mbarrier ring -> gather/reduce teams)
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
