// The blocked (two-pass, column-slab x row-panel) SpMV strategy for sm_100a: x gathers are served from shared memory
// instead of L2.  See PbPlan in internal.h for the layout; DESIGN.md section 5 for the measurements behind it.
//
// What it replaces in the reference (semantics only -- the reference's stream is an FPGA format):
//   tileAndPad's column tiles (a tile covers (num_fp32s_b/2)*1024 columns so that its piece of x fits the
//   on-chip B buffers)                                                     common/src/spmv-helper.cpp:139-227
//   ComputeAB reading x from the on-chip buffer, one product per cycle     automation_tool/assets/base_functions.cpp:158-254
//   AccumBuffer / Compute_C per row tile                                   base_functions.cpp:439-540
// The reference streams every column tile's nonzeros past an x slice held in BRAM/URAM and accumulates y_Ax per row
// tile on chip; here the slab's slice of x lives in the SM's shared memory (pass 1, pb_expand_kernel) and the per-row
// accumulation happens one row panel at a time in shared memory (pass 2, pb_reduce_kernel), with the products crossing
// HBM once in between.
//
// Integer artefacts (slab starts, the blocked order, 16-bit local columns, 16-bit CSR positions, the (panel, slab)
// segment table) are bit-exact against oracle/oracle.c: oracle_pb_plan.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "tile_device.cuh"

namespace hispmv {

namespace {

struct DevBuf {  // RAII for scratch allocations inside one call
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (p) cudaFree(p);
    p = nullptr;
    if (bytes == 0) bytes = 16;
    return check_cuda(cudaMalloc(&p, bytes), "cudaMalloc(scratch)", __FILE__, __LINE__);
  }
  template <typename T>
  T* as() {
    return static_cast<T*>(p);
  }
};

inline int blocks_for(int64_t n, int block) { return (int)std::max<int64_t>(1, (n + block - 1) / block); }

// ================================================================================================================
// plan
// ================================================================================================================
__global__ void pb_key_kernel(const int32_t* __restrict__ col, int64_t nnz, int32_t W, uint16_t* __restrict__ key,
                              uint32_t* __restrict__ idx) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) {
    key[j] = (uint16_t)(col[j] / W);
    idx[j] = (uint32_t)j;
  }
}

__global__ void pb_iota_kernel(uint32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

// ustart[s] = number of sorted entries whose slab is < s, for s = 0 .. S
__global__ void pb_slab_bounds_kernel(const uint16_t* __restrict__ key, int64_t nnz, int32_t S,
                                      int32_t* __restrict__ ustart) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > S) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int32_t)key[mid] < s) lo = mid + 1; else hi = mid;
  }
  ustart[s] = (int32_t)lo;
}

// Storage inside a group of kPbGroup = 512 entries: entry e = 16 * lane + w sits where a warp's coalesced vector loads
// hand lane `lane` its 16 consecutive entries (val: four 128-bit loads, lcol: two).
__host__ __device__ __forceinline__ int64_t pb_phys_val(int64_t k) {
  const int64_t g = k / kPbGroup, e = k % kPbGroup;
  const int64_t lane = e >> 4, w = e & 15;
  return g * kPbGroup + (((w >> 2) << 5) + lane) * 4 + (w & 3);
}
__host__ __device__ __forceinline__ int64_t pb_phys_lcol(int64_t k) {
  const int64_t g = k / kPbGroup, e = k % kPbGroup;
  const int64_t lane = e >> 4, w = e & 15;
  return g * kPbGroup + (((w >> 3) << 5) + lane) * 8 + (w & 7);
}

// blocked copy of entry k of the slab-sorted order; brow[dst] = its row (padding positions keep -1)
__global__ void pb_scatter_kernel(const uint16_t* __restrict__ key, const uint32_t* __restrict__ idx, int64_t nnz,
                                  const int32_t* __restrict__ row_ptr, int32_t rows, const int32_t* __restrict__ col,
                                  const float* __restrict__ val, int32_t W, const int32_t* __restrict__ ustart,
                                  const int32_t* __restrict__ pstart, float* __restrict__ o_val,
                                  uint16_t* __restrict__ o_lcol, int32_t* __restrict__ brow) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int32_t s = key[k];
  const int32_t j = (int32_t)idx[k];
  const int64_t dst = (int64_t)pstart[s] + (k - ustart[s]);
  int32_t lo = 0, hi = rows;  // the row of CSR position j: last r with row_ptr[r] <= j
  while (hi - lo > 1) {
    const int32_t mid = (int32_t)(((int64_t)lo + hi) >> 1);
    if (__ldg(row_ptr + mid) <= j) lo = mid; else hi = mid;
  }
  o_val[pb_phys_val(dst)] = val[j];
  o_lcol[pb_phys_lcol(dst)] = (uint16_t)(col[j] - s * W);
  brow[dst] = lo;
}

// an entry ends a piece when the next blocked position holds another row, is padding, or opens the next group
__global__ void pb_end_kernel(const int32_t* __restrict__ brow, int64_t padded, int32_t* __restrict__ end) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= padded) return;
  const int32_t r = brow[k];
  end[k] = (r >= 0 && (((k + 1) % kPbGroup) == 0 || brow[k + 1] != r)) ? 1 : 0;
}

__global__ void pb_pack_kernel(const int32_t* __restrict__ end, const int32_t* __restrict__ pid, int64_t padded,
                               uint16_t* __restrict__ flags, int32_t* __restrict__ group_base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one word = sixteen entries = one lane's share
  if (i * 16 >= padded) return;
  uint32_t f = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int4 e = *reinterpret_cast<const int4*>(end + i * 16 + q * 4);
    f |= (uint32_t)(e.x | (e.y << 1) | (e.z << 2) | (e.w << 3)) << (4 * q);
  }
  flags[i] = (uint16_t)f;
  if ((i * 16) % kPbGroup == 0) group_base[(i * 16) / kPbGroup] = pid[i * 16];
}

__global__ void pb_piece_emit_kernel(const int32_t* __restrict__ brow, const int32_t* __restrict__ end,
                                     const int32_t* __restrict__ pid, int64_t padded,
                                     const int32_t* __restrict__ pstart, int32_t S, int32_t* __restrict__ piece_row,
                                     int32_t* __restrict__ piece_slab) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= padded || !end[k]) return;
  const int32_t q = pid[k];
  int32_t lo = 0, hi = S;  // the slab of blocked position k: last s with pstart[s] <= k
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (__ldg(pstart + mid) <= k) lo = mid; else hi = mid;
  }
  piece_row[q] = brow[k];
  piece_slab[q] = lo;
}

// pieces sorted by row (stable, so a row's pieces stay in slab order): position i of that order belongs to piece order[i]
__global__ void pb_pcsr_kernel(const uint32_t* __restrict__ order, int64_t np, int32_t* __restrict__ piece_pcsr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) piece_pcsr[order[i]] = (int32_t)i;
}

// prow_ptr[r] = number of pieces whose row is < r
__global__ void pb_prow_ptr_kernel(const uint32_t* __restrict__ row_sorted, int64_t np, int32_t rows,
                                   int32_t* __restrict__ prow_ptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > rows) return;
  int64_t lo = 0, hi = np;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_sorted[mid] < r) lo = mid + 1; else hi = mid;
  }
  prow_ptr[r] = (int32_t)lo;
}

// the panel that holds position i of the per-row piece order: the first t with desc[t].n1 > i
__device__ __forceinline__ int32_t pb_panel_of(const TileDesc* __restrict__ desc, int64_t np, int32_t i) {
  int64_t lo = 0, hi = np - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(&desc[mid].n1) <= i) lo = mid + 1; else hi = mid;
  }
  return (int32_t)lo;
}

__global__ void pb_perm_kernel(const int32_t* __restrict__ piece_pcsr, int64_t np, const TileDesc* __restrict__ desc,
                               int64_t npan, uint16_t* __restrict__ perm, int32_t* __restrict__ pan) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np) return;
  const int32_t i = piece_pcsr[q];
  const int32_t p = pb_panel_of(desc, npan, i);
  perm[q] = (uint16_t)(i - __ldg(&desc[p].n0));
  pan[q] = p;
}

__global__ void pb_head_kernel(const int32_t* __restrict__ slab, const int32_t* __restrict__ pan, int64_t np,
                               int32_t* __restrict__ head) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np) return;
  head[q] = (q == 0 || slab[q] != slab[q - 1] || pan[q] != pan[q - 1]) ? 1 : 0;
}

__global__ void pb_seg_emit_kernel(const int32_t* __restrict__ slab, const int32_t* __restrict__ pan,
                                   const int32_t* __restrict__ head, const int32_t* __restrict__ segidx, int64_t np,
                                   int32_t S, uint64_t* __restrict__ seg_key, int32_t* __restrict__ seg_q) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np || !head[q]) return;
  const int32_t i = segidx[q];
  seg_key[i] = (uint64_t)pan[q] * (uint64_t)S + (uint64_t)slab[q];
  seg_q[i] = (int32_t)q;
}

// length of segment order[i] (segments are contiguous in piece ids: the next head ends it)
__global__ void pb_seg_len_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ seg_q, int64_t nseg,
                                  int64_t np, int32_t* __restrict__ len_sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const int64_t o = order[i];
  const int64_t end = o + 1 < nseg ? seg_q[o + 1] : np;
  len_sorted[i] = (int32_t)(end - seg_q[o]);
}

__global__ void pb_seg_final_kernel(const uint64_t* __restrict__ seg_key_sorted, const uint32_t* __restrict__ order,
                                    const int32_t* __restrict__ seg_q, const int32_t* __restrict__ G,
                                    const TileDesc* __restrict__ desc, int32_t S, int64_t nseg,
                                    PbSeg* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const int64_t p = (int64_t)(seg_key_sorted[i] / (uint64_t)S);
  PbSeg sg;
  sg.start = seg_q[order[i]];
  sg.off = G[i] - __ldg(&desc[p].n0);  // pieces of earlier panels add up to the panel's first position
  out[i] = sg;
}

// panel_seg[p] = number of segments whose panel is < p, p = 0 .. np
__global__ void pb_panel_seg_kernel(const uint64_t* __restrict__ seg_key_sorted, int64_t nseg, int64_t np, int32_t S,
                                    int32_t* __restrict__ panel_seg) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > np) return;
  const uint64_t target = (uint64_t)p * (uint64_t)S;
  int64_t lo = 0, hi = nseg;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (seg_key_sorted[mid] < target) lo = mid + 1; else hi = mid;
  }
  panel_seg[p] = (int32_t)lo;
}

// chunks: segment i (in panel-major order, `len` pieces from piece id seg[i].start) becomes ceil(len / kPbChunk) runs
// (the last segment of every panel is followed by one empty run, so a warp may read one descriptor past its panel's
// last run without a bounds check)
__global__ void pb_chunk_count_kernel(const int32_t* __restrict__ len_sorted, const uint64_t* __restrict__ seg_key_sorted,
                                      int32_t S, int64_t nseg, int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const bool last_of_panel = i + 1 == nseg || seg_key_sorted[i + 1] / (uint64_t)S != seg_key_sorted[i] / (uint64_t)S;
  (void)last_of_panel;
  cnt[i] = (len_sorted[i] + kPbChunk - 1) / kPbChunk;
}
__global__ void pb_chunk_emit_kernel(const PbSeg* __restrict__ seg, const int32_t* __restrict__ len_sorted,
                                     const int32_t* __restrict__ first, int64_t nseg, int32_t total,
                                     int2* __restrict__ chunk) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const int32_t len = len_sorted[i], start = seg[i].start;
  const int32_t end = i + 1 < nseg ? first[i + 1] : total;
  int32_t o = first[i];
  for (int32_t k = 0; k < len; k += kPbChunk) chunk[o++] = make_int2(start + k, min(kPbChunk, len - k));
  if (o < end) chunk[o] = make_int2(0, 0);  // the empty run that closes a panel
}
__global__ void pb_panel_chunk_kernel(const int32_t* __restrict__ panel_seg, const int32_t* __restrict__ first,
                                      int64_t np, int64_t nseg, int32_t total, int32_t* __restrict__ panel_chunk) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > np) return;
  const int32_t sgi = panel_seg[p];
  panel_chunk[p] = sgi < nseg ? first[sgi] : total;
}

__global__ void pb_max_segs_kernel(const int32_t* __restrict__ panel_seg, int64_t np, int* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = 0;
  if (p < np) v = panel_seg[p + 1] - panel_seg[p];
  v = __reduce_max_sync(kFullMask, v);
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// ---- the staged gather of pass 2 (round-2 v10): every (panel, slab) segment of a STREAM panel is fetched by ONE bulk copy
// of the 16-byte-aligned range of part[] that covers it, into the panel's staging area in shared memory; perm2 holds the
// panel-relative slot of every staged position (0xFFFF for the alignment padding), end_bits one bit per slot that ends a
// row.  Segments of LONG panels are not staged (length 0).
__global__ void pb_stage_len_kernel(const PbSeg* __restrict__ seg, const int32_t* __restrict__ len_sorted,
                                    const uint64_t* __restrict__ key_sorted, const TileDesc* __restrict__ desc, int32_t S,
                                    int64_t nseg, int32_t* __restrict__ alen) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const int64_t t = (int64_t)(key_sorted[i] / (uint64_t)S);
  const int32_t b = seg[i].start, e = b + len_sorted[i];
  alen[i] = __ldg(&desc[t].chunk) >= 0 ? 0 : ((e + 3) & ~3) - (b & ~3);
}

__global__ void pb_bit_words_kernel(const TileDesc* __restrict__ desc, int64_t npan, int32_t* __restrict__ words) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npan) return;
  const TileDesc d = desc[t];
  words[t] = d.chunk >= 0 ? 0 : (d.n1 - d.n0 + 31) >> 5;
}

__global__ void pb_panel_aux_kernel(const int32_t* __restrict__ panel_seg, const int32_t* __restrict__ gpos,
                                    const int32_t* __restrict__ bbase, int64_t npan, int64_t nseg, int32_t pos_total,
                                    int32_t bit_total, int2* __restrict__ aux) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > npan) return;
  const int32_t sgi = t < npan ? panel_seg[t] : (int32_t)nseg;
  aux[t] = make_int2(sgi < nseg ? gpos[sgi] : pos_total, t < npan ? bbase[t] : bit_total);
}

// one warp per segment: its copy descriptor and the slots of its staged positions
__global__ void pb_seg_copy_kernel(const PbSeg* __restrict__ seg, const int32_t* __restrict__ len_sorted,
                                   const uint64_t* __restrict__ key_sorted, const int32_t* __restrict__ alen,
                                   const int32_t* __restrict__ gpos, const int2* __restrict__ aux,
                                   const uint16_t* __restrict__ perm, int32_t S, int64_t nseg, int2* __restrict__ seg_copy,
                                   uint16_t* __restrict__ perm2, int32_t* __restrict__ chunk_src) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= nseg) return;
  const int64_t t = (int64_t)(key_sorted[i] / (uint64_t)S);
  const int32_t b = seg[i].start, e = b + len_sorted[i], a0 = b & ~3, al = alen[i], g = gpos[i];
  if (lane == 0) seg_copy[i] = make_int2(a0, ((g - aux[t].x) >> 2) | ((al >> 2) << 16));
  for (int32_t j = lane; j < al; j += 32) {
    const int32_t q = a0 + j;
    perm2[(int64_t)g + j] = (q >= b && q < e) ? perm[q] : (uint16_t)0xFFFF;
    if ((j & 3) == 0) chunk_src[((int64_t)g + j) >> 2] = q;  // where each staged quad of partial sums comes from
  }
}

__global__ void pb_end_bits_kernel(const int32_t* __restrict__ prow_ptr, int32_t rows, const TileDesc* __restrict__ desc,
                                   int64_t npan, const int2* __restrict__ aux, uint32_t* __restrict__ bits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int32_t b = prow_ptr[r], e = prow_ptr[r + 1];
  if (e <= b) return;
  const int32_t t = pb_panel_of(desc, npan, e - 1);  // the panel that holds the row's last piece
  const TileDesc d = desc[t];
  if (d.chunk >= 0) return;
  const int32_t j = e - 1 - d.n0;
  atomicOr(&bits[(int64_t)aux[t].y + (j >> 5)], 1u << (j & 31));
}

// shared-memory words pass 2 needs for its largest STREAM panel: the skewed slots
__global__ void pb_reduce_words_kernel(const TileDesc* __restrict__ desc, const int2* __restrict__ aux, int64_t npan,
                                       int* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = 0;
  if (t < npan && desc[t].chunk < 0) {
    const int n = desc[t].n1 - desc[t].n0;
    v = ((n + 31) >> 5) * 33 + 4;
  }
  v = __reduce_max_sync(kFullMask, v);
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

int grow(DevBuf& tmp, size_t& have, size_t want) {
  if (want <= have) return HISPMV_OK;
  have = want;
  return tmp.alloc(want);
}

}  // namespace

void pb_free(PbArrays* a) {
  cudaFree(a->d_slab_ptr);
  cudaFree(a->d_val);
  cudaFree(a->d_lcol);
  cudaFree(a->d_flags);
  cudaFree(a->d_group_base);
  cudaFree(a->d_prow_ptr);
  cudaFree(a->d_perm);
  cudaFree(a->d_panel_seg);
  cudaFree(a->d_seg);
  cudaFree(a->d_panel_chunk);
  cudaFree(a->d_chunk);
  cudaFree(a->d_seg_copy);
  cudaFree(a->d_perm2);
  cudaFree(a->d_chunk_src);
  cudaFree(a->d_panel_aux);
  cudaFree(a->d_end_bits);
  cudaFree(a->d_work);
  cudaFree(a->d_part[0]);
  cudaFree(a->d_part[1]);
  cudaFree(a->d_piece_pcsr);
  cudaFree(a->d_piece_slab);
  delete[] a->h_slab_ptr;
  *a = PbArrays();
}

int pb_order_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t rows, int32_t cols,
                    int64_t nnz, int32_t slab_cols, PbArrays* out, cudaStream_t stream) {
  pb_free(out);
  if (slab_cols < 4 || slab_cols > kPbMaxSlabCols || (slab_cols & 3) || cols <= 0 || nnz <= 0 || rows <= 0) {
    set_error("blocked plan: bad slab width or empty matrix");
    return HISPMV_ERR_ARG;
  }
  const int32_t W = slab_cols;
  const int64_t S64 = ((int64_t)cols + W - 1) / W;
  if (S64 > 65535) {
    set_error("blocked plan: more than 65535 column slabs");
    return HISPMV_ERR_ARG;
  }
  const int32_t S = (int32_t)S64;
  const int B = 256;
  int st;
  int sbits = 1;
  while ((1 << sbits) < S) ++sbits;

  // ---- stable sort of the CSR positions by slab: slab-major, CSR order inside a slab -----------------------------
  DevBuf key_a, key_b, idx_a, idx_b, tmp, ustart, pstart;
  size_t tb = 0;
  if ((st = key_a.alloc((size_t)nnz * 2)) || (st = key_b.alloc((size_t)nnz * 2)) || (st = idx_a.alloc((size_t)nnz * 4)) ||
      (st = idx_b.alloc((size_t)nnz * 4)) || (st = ustart.alloc(((size_t)S + 1) * 4)) ||
      (st = pstart.alloc(((size_t)S + 1) * 4)))
    return st;
  cub::DoubleBuffer<uint16_t> keys(key_a.as<uint16_t>(), key_b.as<uint16_t>());
  cub::DoubleBuffer<uint32_t> idx(idx_a.as<uint32_t>(), idx_b.as<uint32_t>());
  pb_key_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(d_col, nnz, W, keys.Current(), idx.Current());
  HISPMV_CUDA(cudaGetLastError());
  size_t need = 0;
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, idx, nnz, 0, sbits, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, keys, idx, nnz, 0, sbits, stream));
  pb_slab_bounds_kernel<<<blocks_for((int64_t)S + 1, 128), 128, 0, stream>>>(keys.Current(), nnz, S,
                                                                            ustart.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  std::vector<int32_t> h_ustart((size_t)S + 1);
  HISPMV_CUDA(cudaMemcpyAsync(h_ustart.data(), ustart.p, ((size_t)S + 1) * 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  // every slab starts at a multiple of kPbGroup in the blocked arrays (a warp's group never straddles two slabs)
  out->h_slab_ptr = new int32_t[(size_t)S + 1];
  int64_t pos = 0;
  for (int32_t s = 0; s < S; ++s) {
    out->h_slab_ptr[s] = (int32_t)pos;
    pos += (int64_t)h_ustart[(size_t)s + 1] - h_ustart[(size_t)s];
    pos = (pos + kPbGroup - 1) / kPbGroup * kPbGroup;
    if (pos >= (int64_t)INT32_MAX - kPbGroup) {
      set_error("blocked plan: more than 2^31 entries");
      return HISPMV_ERR_ARG;
    }
  }
  out->h_slab_ptr[S] = (int32_t)pos;
  const int64_t padded = pos;
  HISPMV_CUDA(cudaMemcpyAsync(pstart.p, out->h_slab_ptr, ((size_t)S + 1) * 4, cudaMemcpyHostToDevice, stream));

  // ---- the blocked copy ------------------------------------------------------------------------------------------
  const size_t slack = 64;  // vector loads of the last group may look past the end
  const int64_t ngroups = padded / kPbGroup;
  HISPMV_CUDA(cudaMalloc((void**)&out->d_val, ((size_t)padded + slack) * 4));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_lcol, ((size_t)padded + slack) * 2));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_flags, (size_t)padded / 8 + slack));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_group_base, ((size_t)ngroups + 1) * 4));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_slab_ptr, ((size_t)S + 1) * 4));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_prow_ptr, ((size_t)rows + 1 + 4) * 4));
  HISPMV_CUDA(cudaMemsetAsync(out->d_val, 0, ((size_t)padded + slack) * 4, stream));
  HISPMV_CUDA(cudaMemsetAsync(out->d_lcol, 0, ((size_t)padded + slack) * 2, stream));
  HISPMV_CUDA(cudaMemsetAsync(out->d_flags, 0, (size_t)padded / 8 + slack, stream));
  HISPMV_CUDA(cudaMemcpyAsync(out->d_slab_ptr, pstart.p, ((size_t)S + 1) * 4, cudaMemcpyDeviceToDevice, stream));
  DevBuf brow, end, pid;
  if ((st = brow.alloc(((size_t)padded + 4) * 4))) return st;
  HISPMV_CUDA(cudaMemsetAsync(brow.p, 0xff, ((size_t)padded + 4) * 4, stream));  // -1: padding
  pb_scatter_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(keys.Current(), idx.Current(), nnz, d_row_ptr, rows, d_col,
                                                          d_val, W, ustart.as<int32_t>(), pstart.as<int32_t>(),
                                                          out->d_val, out->d_lcol, brow.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  key_a.alloc(0);  // the sort buffers are no longer needed
  key_b.alloc(0);
  idx_a.alloc(0);
  idx_b.alloc(0);

  // ---- pieces ----------------------------------------------------------------------------------------------------
  if ((st = end.alloc(((size_t)padded + 4) * 4)) || (st = pid.alloc(((size_t)padded + 4) * 4))) return st;
  pb_end_kernel<<<blocks_for(padded, B), B, 0, stream>>>(brow.as<int32_t>(), padded, end.as<int32_t>());
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, end.as<int32_t>(), pid.as<int32_t>(), padded, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, end.as<int32_t>(), pid.as<int32_t>(), padded, stream));
  int32_t h_last[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(&h_last[0], end.as<int32_t>() + (padded - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h_last[1], pid.as<int32_t>() + (padded - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t np = (int64_t)h_last[0] + h_last[1];
  pb_pack_kernel<<<blocks_for(padded / 16, B), B, 0, stream>>>(end.as<int32_t>(), pid.as<int32_t>(), padded, out->d_flags,
                                                              out->d_group_base);
  const int32_t np32 = (int32_t)np;
  HISPMV_CUDA(cudaMemcpyAsync(out->d_group_base + ngroups, &np32, 4, cudaMemcpyHostToDevice, stream));
  DevBuf prow_a, prow_b, ord_a, ord_b;
  HISPMV_CUDA(cudaMalloc((void**)&out->d_piece_pcsr, (size_t)std::max<int64_t>(np, 1) * 4));
  HISPMV_CUDA(cudaMalloc((void**)&out->d_piece_slab, (size_t)std::max<int64_t>(np, 1) * 4));
  if ((st = prow_a.alloc((size_t)np * 4)) || (st = prow_b.alloc((size_t)np * 4)) || (st = ord_a.alloc((size_t)np * 4)) ||
      (st = ord_b.alloc((size_t)np * 4)))
    return st;
  pb_piece_emit_kernel<<<blocks_for(padded, B), B, 0, stream>>>(brow.as<int32_t>(), end.as<int32_t>(), pid.as<int32_t>(),
                                                                padded, pstart.as<int32_t>(), S, prow_a.as<int32_t>(),
                                                                out->d_piece_slab);
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  brow.alloc(0);
  end.alloc(0);
  pid.alloc(0);
  // a row's pieces in slab order: stable sort of the piece ids by row
  cub::DoubleBuffer<uint32_t> prow(prow_a.as<uint32_t>(), prow_b.as<uint32_t>());
  cub::DoubleBuffer<uint32_t> ord(ord_a.as<uint32_t>(), ord_b.as<uint32_t>());
  pb_iota_kernel<<<blocks_for(np, B), B, 0, stream>>>(ord.Current(), np);
  int rbits = 1;
  while (rbits < 32 && ((int64_t)1 << rbits) < (int64_t)rows) ++rbits;
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, prow, ord, np, 0, rbits, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, prow, ord, np, 0, rbits, stream));
  pb_pcsr_kernel<<<blocks_for(np, B), B, 0, stream>>>(ord.Current(), np, out->d_piece_pcsr);
  pb_prow_ptr_kernel<<<blocks_for((int64_t)rows + 1, B), B, 0, stream>>>(prow.Current(), np, rows, out->d_prow_ptr);
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  out->slab_cols = W;
  out->num_slabs = S;
  out->padded_nnz = padded;
  out->num_pieces = np;
  return HISPMV_OK;
}

int pb_segments_device(PbArrays* a, const TileDesc* d_desc, int64_t num_panels, int32_t rows, cudaStream_t stream) {
  const int64_t np = a->num_pieces;
  const int32_t S = a->num_slabs;
  const int B = 256;
  int st;
  if (np <= 0 || num_panels <= 0 || !a->d_piece_pcsr) {
    set_error("blocked plan: no pieces");
    return HISPMV_ERR_STATE;
  }
  DevBuf pan, head, segidx, tmp;
  size_t tb = 0, need = 0;
  HISPMV_CUDA(cudaMalloc((void**)&a->d_perm, ((size_t)np + 64) * 2));
  if ((st = pan.alloc((size_t)np * 4)) || (st = head.alloc((size_t)np * 4)) || (st = segidx.alloc((size_t)np * 4)))
    return st;
  pb_perm_kernel<<<blocks_for(np, B), B, 0, stream>>>(a->d_piece_pcsr, np, d_desc, num_panels, a->d_perm,
                                                      pan.as<int32_t>());
  pb_head_kernel<<<blocks_for(np, B), B, 0, stream>>>(a->d_piece_slab, pan.as<int32_t>(), np, head.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, head.as<int32_t>(), segidx.as<int32_t>(), np, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, head.as<int32_t>(), segidx.as<int32_t>(), np, stream));
  int32_t h_last[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(&h_last[0], head.as<int32_t>() + (np - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h_last[1], segidx.as<int32_t>() + (np - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t nseg = (int64_t)h_last[0] + h_last[1];
  DevBuf seg_q, skey_a, skey_b, ord_a, ord_b, len_sorted, G;
  if ((st = seg_q.alloc((size_t)nseg * 4)) || (st = skey_a.alloc((size_t)nseg * 8)) ||
      (st = skey_b.alloc((size_t)nseg * 8)) || (st = ord_a.alloc((size_t)nseg * 4)) ||
      (st = ord_b.alloc((size_t)nseg * 4)) || (st = len_sorted.alloc((size_t)nseg * 4)) || (st = G.alloc((size_t)nseg * 4)))
    return st;
  pb_seg_emit_kernel<<<blocks_for(np, B), B, 0, stream>>>(a->d_piece_slab, pan.as<int32_t>(), head.as<int32_t>(),
                                                          segidx.as<int32_t>(), np, S, skey_a.as<uint64_t>(),
                                                          seg_q.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> skeys(skey_a.as<uint64_t>(), skey_b.as<uint64_t>());
  cub::DoubleBuffer<uint32_t> ords(ord_a.as<uint32_t>(), ord_b.as<uint32_t>());
  pb_iota_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(ords.Current(), nseg);
  HISPMV_CUDA(cudaGetLastError());
  int kbits = 1;
  {
    const uint64_t top = (uint64_t)num_panels * (uint64_t)S;
    while (kbits < 64 && ((uint64_t)1 << kbits) < top) ++kbits;
  }
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, skeys, ords, nseg, 0, kbits, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, skeys, ords, nseg, 0, kbits, stream));
  pb_seg_len_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(ords.Current(), seg_q.as<int32_t>(), nseg, np,
                                                           len_sorted.as<int32_t>());
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, len_sorted.as<int32_t>(), G.as<int32_t>(), nseg, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, len_sorted.as<int32_t>(), G.as<int32_t>(), nseg, stream));
  HISPMV_CUDA(cudaMalloc((void**)&a->d_seg, ((size_t)nseg + 1) * sizeof(PbSeg)));
  HISPMV_CUDA(cudaMalloc((void**)&a->d_panel_seg, ((size_t)num_panels + 1) * 4));
  pb_seg_final_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(skeys.Current(), ords.Current(), seg_q.as<int32_t>(),
                                                             G.as<int32_t>(), d_desc, S, nseg, a->d_seg);
  pb_panel_seg_kernel<<<blocks_for(num_panels + 1, B), B, 0, stream>>>(skeys.Current(), nseg, num_panels, S,
                                                                      a->d_panel_seg);
  // the chunk table pass 2 walks: every segment cut into runs of at most kPbChunk pieces
  DevBuf ccnt, cfirst;
  if ((st = ccnt.alloc((size_t)nseg * 4)) || (st = cfirst.alloc((size_t)nseg * 4))) return st;
  pb_chunk_count_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(len_sorted.as<int32_t>(), skeys.Current(), S, nseg,
                                                               ccnt.as<int32_t>());
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, ccnt.as<int32_t>(), cfirst.as<int32_t>(), nseg, stream));
  if ((st = grow(tmp, tb, need))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, ccnt.as<int32_t>(), cfirst.as<int32_t>(), nseg, stream));
  int32_t h_c[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(&h_c[0], ccnt.as<int32_t>() + (nseg - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h_c[1], cfirst.as<int32_t>() + (nseg - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t nchunk = (int64_t)h_c[0] + h_c[1];
  HISPMV_CUDA(cudaMalloc((void**)&a->d_chunk, ((size_t)nchunk + 8) * sizeof(int2)));
  HISPMV_CUDA(cudaMemsetAsync(a->d_chunk + nchunk, 0, 8 * sizeof(int2), stream));  // reads past a panel's last pair see count 0
  HISPMV_CUDA(cudaMalloc((void**)&a->d_panel_chunk, ((size_t)num_panels + 1) * 4));
  pb_chunk_emit_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(a->d_seg, len_sorted.as<int32_t>(), cfirst.as<int32_t>(),
                                                              nseg, (int32_t)nchunk, a->d_chunk);
  pb_panel_chunk_kernel<<<blocks_for(num_panels + 1, B), B, 0, stream>>>(a->d_panel_seg, cfirst.as<int32_t>(), num_panels,
                                                                        nseg, (int32_t)nchunk, a->d_panel_chunk);
  HISPMV_CUDA(cudaGetLastError());
  a->num_chunks = nchunk;
  DevBuf mx;
  if ((st = mx.alloc(sizeof(int)))) return st;
  HISPMV_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), stream));
  pb_max_segs_kernel<<<blocks_for(num_panels, B), B, 0, stream>>>(a->d_panel_chunk, num_panels, mx.as<int>());  // most runs of a panel
  HISPMV_CUDA(cudaGetLastError());
  int h_mx = 0;
  HISPMV_CUDA(cudaMemcpyAsync(&h_mx, mx.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  a->num_seg = nseg;
  a->max_panel_segs = h_mx;
  // the staged gather: aligned copy ranges, their places in each panel's staging area, perm2, end bits
  {
    DevBuf alen, gpos, bw, bbase, mw;
    if ((st = alen.alloc((size_t)nseg * 4)) || (st = gpos.alloc((size_t)nseg * 4)) ||
        (st = bw.alloc((size_t)num_panels * 4)) || (st = bbase.alloc((size_t)num_panels * 4)) || (st = mw.alloc(sizeof(int))))
      return st;
    pb_stage_len_kernel<<<blocks_for(nseg, B), B, 0, stream>>>(a->d_seg, len_sorted.as<int32_t>(), skeys.Current(), d_desc, S,
                                                               nseg, alen.as<int32_t>());
    pb_bit_words_kernel<<<blocks_for(num_panels, B), B, 0, stream>>>(d_desc, num_panels, bw.as<int32_t>());
    HISPMV_CUDA(cudaGetLastError());
    HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, alen.as<int32_t>(), gpos.as<int32_t>(), nseg, stream));
    if ((st = grow(tmp, tb, need))) return st;
    HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, alen.as<int32_t>(), gpos.as<int32_t>(), nseg, stream));
    HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, bw.as<int32_t>(), bbase.as<int32_t>(), num_panels, stream));
    if ((st = grow(tmp, tb, need))) return st;
    HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, need, bw.as<int32_t>(), bbase.as<int32_t>(), num_panels, stream));
    int32_t h_t[4] = {0, 0, 0, 0};
    HISPMV_CUDA(cudaMemcpyAsync(&h_t[0], alen.as<int32_t>() + (nseg - 1), 4, cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaMemcpyAsync(&h_t[1], gpos.as<int32_t>() + (nseg - 1), 4, cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaMemcpyAsync(&h_t[2], bw.as<int32_t>() + (num_panels - 1), 4, cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaMemcpyAsync(&h_t[3], bbase.as<int32_t>() + (num_panels - 1), 4, cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaStreamSynchronize(stream));
    a->stage_total = (int64_t)h_t[0] + h_t[1];
    a->bit_words = (int64_t)h_t[2] + h_t[3];
    HISPMV_CUDA(cudaMalloc((void**)&a->d_seg_copy, ((size_t)nseg + 1) * sizeof(int2)));
    HISPMV_CUDA(cudaMalloc((void**)&a->d_perm2, ((size_t)a->stage_total + 64) * 2));
    HISPMV_CUDA(cudaMalloc((void**)&a->d_chunk_src, ((size_t)a->stage_total / 4 + 16) * 4));
    HISPMV_CUDA(cudaMemsetAsync(a->d_chunk_src + a->stage_total / 4, 0, 16 * 4, stream));
    HISPMV_CUDA(cudaMalloc((void**)&a->d_panel_aux, ((size_t)num_panels + 1) * sizeof(int2)));
    HISPMV_CUDA(cudaMalloc((void**)&a->d_end_bits, ((size_t)a->bit_words + 4) * 4));
    HISPMV_CUDA(cudaMemsetAsync(a->d_end_bits, 0, ((size_t)a->bit_words + 4) * 4, stream));
    HISPMV_CUDA(cudaMemsetAsync(a->d_perm2 + a->stage_total, 0xFF, 64 * 2, stream));
    HISPMV_CUDA(cudaMemsetAsync(mw.p, 0, sizeof(int), stream));
    pb_panel_aux_kernel<<<blocks_for(num_panels + 1, B), B, 0, stream>>>(a->d_panel_seg, gpos.as<int32_t>(),
                                                                        bbase.as<int32_t>(), num_panels, nseg,
                                                                        (int32_t)a->stage_total, (int32_t)a->bit_words,
                                                                        a->d_panel_aux);
    pb_seg_copy_kernel<<<blocks_for(nseg * 32, B), B, 0, stream>>>(a->d_seg, len_sorted.as<int32_t>(), skeys.Current(),
                                                                   alen.as<int32_t>(), gpos.as<int32_t>(), a->d_panel_aux,
                                                                   a->d_perm, S, nseg, a->d_seg_copy, a->d_perm2,
                                                                   a->d_chunk_src);
    pb_end_bits_kernel<<<blocks_for(rows, B), B, 0, stream>>>(a->d_prow_ptr, rows, d_desc, num_panels, a->d_panel_aux,
                                                              a->d_end_bits);
    pb_reduce_words_kernel<<<blocks_for(num_panels, B), B, 0, stream>>>(d_desc, a->d_panel_aux, num_panels, mw.as<int>());
    HISPMV_CUDA(cudaGetLastError());
    int h_mw = 0;
    HISPMV_CUDA(cudaMemcpyAsync(&h_mw, mw.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaStreamSynchronize(stream));
    a->reduce_words = h_mw;
  }
  cudaFree(a->d_piece_pcsr);
  cudaFree(a->d_piece_slab);
  a->d_piece_pcsr = nullptr;
  a->d_piece_slab = nullptr;
  return HISPMV_OK;
}

// Pass-1 work ranges: n_cta contiguous pieces of the blocked order, balanced by entries + slab_cost per slab a CTA
// has to stage (a CTA that walks many thin slabs of the column tail spends its time loading x, not streaming).
namespace {
// n_cta contiguous runs of the groups [g_begin, g_end), balanced by entries + pieces + slab loads (see pb_make_work)
void pb_partition_groups(const PbArrays* a, const std::vector<int32_t>& gb, int64_t g_begin, int64_t g_end, int n_cta,
                         int64_t slab_cost, int64_t piece_cost16, int2* work) {
  const int32_t S = a->num_slabs;
  const int32_t* sp = a->h_slab_ptr;
  // a group costs its 512 entries plus piece_cost16/16 entries for every piece it ends (a piece is a scan step, a
  // staged value and 4 bytes stored: groups of the thin column tail, where nearly every entry is a piece, take longer)
  auto group_cost = [&](int64_t g) { return (int64_t)kPbGroup + piece_cost16 * (int64_t)(gb[(size_t)g + 1] - gb[(size_t)g]) / 16; };
  int64_t slabs_used = 0;
  for (int32_t s = 0; s < S; ++s)
    slabs_used += sp[s + 1] > sp[s] && (int64_t)sp[s + 1] > g_begin * kPbGroup && (int64_t)sp[s] < g_end * kPbGroup;
  int64_t remaining = slab_cost * slabs_used;
  for (int64_t g = g_begin; g < g_end; ++g) remaining += group_cost(g);
  int64_t g = g_begin;
  int32_t s = 0;
  for (int b = 0; b < n_cta; ++b) {
    const int64_t g0 = g;
    int64_t budget = (remaining + (n_cta - b) - 1) / (n_cta - b);
    int64_t spent = 0;
    bool fresh = true;  // the CTA has to stage the slab it starts in, even when the previous CTA already paid for it
    while (g < g_end && (budget > 0 || b == n_cta - 1)) {
      const int64_t k = g * kPbGroup;
      while (s < S && sp[s + 1] <= k) {
        ++s;
        fresh = true;
      }
      if (s >= S) break;
      if (fresh) {
        if (k == sp[s]) {  // first CTA on this slab: the cost is part of `remaining`
          budget -= slab_cost;
          spent += slab_cost;
        }
        fresh = false;
        if (budget <= 0 && g > g0 && b != n_cta - 1) break;
      }
      const int64_t c = group_cost(g);
      ++g;
      budget -= c;
      spent += c;
    }
    if (b == n_cta - 1) g = g_end;
    work[b] = make_int2((int)(g0 * kPbGroup), (int)(g * kPbGroup));
    remaining -= spent;
    if (remaining < 0) remaining = 0;
  }
}
}  // namespace

int pb_make_work(PbArrays* a, int n_cta, int64_t slab_cost, int64_t piece_cost16, cudaStream_t stream) {
  cudaFree(a->d_work);
  a->d_work = nullptr;
  a->num_work = 0;
  a->head_cols = 0;
  if (n_cta < 1) n_cta = 1;
  const int32_t S = a->num_slabs;
  const int32_t* sp = a->h_slab_ptr;
  const int64_t ngroups = a->padded_nnz / kPbGroup;
  std::vector<int32_t> gb((size_t)ngroups + 1);
  HISPMV_CUDA(cudaMemcpyAsync(gb.data(), a->d_group_base, ((size_t)ngroups + 1) * 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  // [0, n_cta): the whole blocked order.  [n_cta, 3 n_cta): the same order in two parts for the host-buffer call -- the
  // HEAD slabs (the fewest that hold two thirds of the entries, at most a quarter of the columns) and the rest, each cut
  // into n_cta ranges of its own, so that pass 1 over the head runs while the rest of x is still crossing PCIe and the
  // tail part, once x is complete, is spread over all SMs
  std::vector<int2> work((size_t)3 * n_cta);
  pb_partition_groups(a, gb, 0, ngroups, n_cta, slab_cost, piece_cost16, work.data());
  int32_t sh = 0;
  while (sh < S && (int64_t)sp[sh] * 3 < (int64_t)a->padded_nnz * 2 && (int64_t)(sh + 1) * 4 <= (int64_t)S) ++sh;
  if (sh > 0 && sh < S && sp[sh] > 0 && sp[sh] < a->padded_nnz) {
    pb_partition_groups(a, gb, 0, sp[sh] / kPbGroup, n_cta, slab_cost, piece_cost16, work.data() + n_cta);
    pb_partition_groups(a, gb, sp[sh] / kPbGroup, ngroups, n_cta, slab_cost, piece_cost16, work.data() + 2 * n_cta);
    a->head_cols = (int64_t)sh * a->slab_cols;
  }
  HISPMV_CUDA(cudaMalloc((void**)&a->d_work, work.size() * sizeof(int2)));
  HISPMV_CUDA(cudaMemcpyAsync(a->d_work, work.data(), work.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  a->num_work = n_cta;
  return HISPMV_OK;
}

// ================================================================================================================
// pass 1: one partial sum per piece, the slab's piece of x staged in shared memory by TMA bulk copies
// ================================================================================================================
namespace {

constexpr int kExpandThreads = 512;
constexpr int kBulkPiece = 4096;  // floats per cp.async.bulk (one copy costs its issuing thread ~650 cycles: 12 lanes
                                  // issue the 12 pieces of a 48 K-column slab side by side)

// One warp, one group of 512 consecutive entries: lane l holds entries 16 l .. 16 l + 15 (values v[], local columns
// c[], end flags f).  Every piece's total goes to stage[rank of the piece inside the group]; returns the group's piece
// count.  History (C2, 100 M entries): 4 entries per lane with branches cost 240 warp instructions per 128 entries and
// was issue-bound at 35 % of HBM bandwidth; straight-line code 162 (46 %); 16 entries per lane amortise the warp-wide
// scan -- the only cross-lane step -- over four times as many entries (71 M instructions instead of 187 M), but with
// every lane storing its own pieces the 16 scalar stores of a lane hit 16 different sectors (7.5x write amplification,
// 286 us): the pieces are staged in shared memory and leave as coalesced 128-byte stores.  The shuffle tree and the
// in-lane order are fixed, so the sums are bit-reproducible.
struct GroupRegs {
  float4 v[4];
  uint4 c[2];
  uint32_t f;
  int base;
};

__device__ __forceinline__ int group_pieces(const GroupRegs& G, const int lane, const float* __restrict__ s_x,
                                            float* __restrict__ stage) {
  float p[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t w0 = k < 2 ? (k == 0 ? G.c[0].x : G.c[0].z) : (k == 2 ? G.c[1].x : G.c[1].z);
    const uint32_t w1 = k < 2 ? (k == 0 ? G.c[0].y : G.c[0].w) : (k == 2 ? G.c[1].y : G.c[1].w);
    p[4 * k + 0] = G.v[k].x * s_x[w0 & 0xffffu];
    p[4 * k + 1] = G.v[k].y * s_x[w0 >> 16];
    p[4 * k + 2] = G.v[k].z * s_x[w1 & 0xffffu];
    p[4 * k + 3] = G.v[k].w * s_x[w1 >> 16];
  }
  const uint32_t f = G.f;
  if (__all_sync(kFullMask, f == 0xffffu)) {  // every entry is its own piece (hypersparse rows): products are the partials
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(stage + 16 * lane + 4 * k) = make_float4(p[4 * k], p[4 * k + 1], p[4 * k + 2], p[4 * k + 3]);
    return kPbGroup;
  }
  // what the lane leaves open: the sum after its last end (everything when it has no end)
  float tail = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) tail = (f & (1u << j)) ? 0.0f : tail + p[j];
  // inclusive segmented scan over the lanes of (tail, lane has an end); the piece counts ride in the low bits of the
  // flag word (bit 31 = some lane up to here has an end)
  const int cnt = __popc(f);
  float sv = tail;
  uint32_t sr = (uint32_t)cnt | (f ? 0x80000000u : 0u);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float uv = __shfl_up_sync(kFullMask, sv, d);
    const uint32_t ur = __shfl_up_sync(kFullMask, sr, d);
    if (lane >= d) {
      if (!(sr & 0x80000000u)) sv += uv;
      sr = (sr + (ur & 0x7fffffffu)) | (ur & 0x80000000u);
    }
  }
  float run = __shfl_up_sync(kFullMask, sv, 1);  // the open piece's sum over the lanes before this one
  if (lane == 0) run = 0.0f;                     // pieces never cross a group
  const int incl = (int)(sr & 0x7fffffffu);
  float* o = stage + (incl - cnt);               // pieces that end in earlier lanes come first
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    run += p[j];
    if (f & (1u << j)) {
      *o++ = run;
      run = 0.0f;
    }
  }
  return __shfl_sync(kFullMask, incl, 31);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    pb_expand_kernel(PbPlan P, const float* __restrict__ x, int32_t cols) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  float* s_x = reinterpret_cast<float*>(s_raw);  // [slab_cols]
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = THREADS / 32;
  float* s_stage = s_x + P.slab_cols + warp * kPbGroup;  // this warp's pieces of one group
  const int2 w = P.work[P.work_begin + blockIdx.x];
  if (w.x >= w.y) return;
  long long dbg_t0 = 0, dbg_loads = 0;
  if (P.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  const int32_t* __restrict__ slab_ptr = P.slab_ptr;
  const float* __restrict__ g_val = P.val;
  const uint16_t* __restrict__ g_lcol = P.lcol;
  const uint16_t* __restrict__ g_flags = P.flags;
  const int32_t* __restrict__ g_base = P.group_base;
  float* __restrict__ g_part = P.part;
  const int slab_cols = P.slab_cols;
  int s;
  {  // the slab that holds entry w.x: the first s with slab_ptr[s + 1] > w.x
    int lo = 0, hi = P.num_slabs - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(slab_ptr + mid + 1) <= w.x) lo = mid + 1; else hi = mid;
    }
    s = lo;
  }
  const uint64_t ps = policy_evict_first(), pk = policy_evict_last();
  const bool x_aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  uint32_t parity = 0;
  auto load_group = [&](int g, GroupRegs& G) {
    const float* pv = g_val + (size_t)g * kPbGroup + lane * 4;
    const uint16_t* pc = g_lcol + (size_t)g * kPbGroup + lane * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) G.v[q] = ld_stream_f4(pv + q * 128, ps);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int4 t = ld_stream_i4(reinterpret_cast<const int32_t*>(pc + q * 256), ps);
      G.c[q] = make_uint4((uint32_t)t.x, (uint32_t)t.y, (uint32_t)t.z, (uint32_t)t.w);
    }
    G.f = __ldg(g_flags + (size_t)g * 32 + lane);
    G.base = __ldg(g_base + g);
  };
  int k = w.x;
  while (k < w.y) {
    const int kend = min(w.y, __ldg(slab_ptr + s + 1));
    if (kend > k) {
      const int c0 = s * slab_cols;
      const int n = min(slab_cols, cols - c0);
      const int nb = x_aligned ? (n & ~3) : 0;  // floats that travel by bulk copy (c0 is a multiple of 4)
      if (nb > 0 && tid < 32) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous slab was read through the generic proxy
        if (tid == 0) mbar_expect_tx(&bar, (uint32_t)nb * 4u);
        __syncwarp();
        for (int p = tid * kBulkPiece; p < nb; p += 32 * kBulkPiece)
          bulk_g2s_hint(s_x + p, x + c0 + p, (uint32_t)min(kBulkPiece, nb - p) * 4u, &bar, pk);
      }
      // groups of this slab's range: warp w takes group g0 + w, g0 + w + WARPS, ...; the next group's loads are issued
      // before the current one is worked on (the first one's travel with the slab)
      const int g_end = kend / kPbGroup;
      int g = k / kPbGroup + warp;
      GroupRegs cur, nxt;
      if (g < g_end) load_group(g, cur);
      for (int i = nb + tid; i < n; i += THREADS) s_x[i] = ld_x_keep(x + c0 + i, pk);
      if (nb > 0) {
        if (tid == 0) mbar_wait(&bar, parity);
        parity ^= 1u;
      }
      __syncthreads();
      while (g < g_end) {
        const int gn = g + WARPS;
        if (gn < g_end) load_group(gn, nxt);
        const int count = group_pieces(cur, lane, s_x, s_stage);
        __syncwarp();
        float* out = g_part + cur.base;
        for (int i = lane; i < count; i += 32) out[i] = s_stage[i];
        __syncwarp();
        cur = nxt;
        g = gn;
      }
      __syncthreads();  // every gather from this slab has been issued before the next one overwrites it
      ++dbg_loads;
    }
    k = kend;
    ++s;
  }
  if (P.dbg && tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    P.dbg[3 * blockIdx.x] = t1 - dbg_t0;
    P.dbg[3 * blockIdx.x + 1] = dbg_loads;
    P.dbg[3 * blockIdx.x + 2] = (long long)((w.y - w.x) / kPbGroup) |
                                ((long long)(__ldg(g_base + w.y / kPbGroup) - __ldg(g_base + w.x / kPbGroup)) << 32);
  }
}

// ================================================================================================================
// pass 2: one CTA per panel
// ================================================================================================================
// History (C2, 30 M pieces): the first version spent 130 thread instructions per piece (a per-lane segment walk with a
// shuffle per step, a lane-per-row reduction whose trip count followed the longest row of each pass) and was
// issue-bound at 20 % of HBM bandwidth; hints + a thread-sequential sweep that tracked row indices still paid for the
// 31 % empty rows of C2 one by one.  Now:
//   gather   the plan cuts every (panel, slab) segment into runs of at most 16 pieces (PbPlan::chunk); a half-warp
//            takes one run per step -- descriptor broadcast, 16 partials and their 16-bit places (perm) coalesced --
//            and drops each partial at its place in the panel's per-row order; four runs in flight, the next four
//            descriptors already requested
//   reduce   s_bits marks the slot that ends each row; every thread adds up C consecutive slots (C odd: no bank
//            conflicts) and closes rows at the marks; a row that spans threads is closed by a segmented scan over the
//            threads' open sums (shuffles inside a warp, shared memory across the eight warps, fixed order); the
//            row's total waits in its last slot and a final row-per-thread pass applies alpha / beta / ReLU with
//            coalesced bias loads and y stores
constexpr int kReduceThreads = 512;
constexpr int kRowsAhead = 6;  // rows per thread (of 512) whose bias and extent are requested before the gather

// shared-memory accesses by 32-bit shared address (the compiler otherwise re-derives the dynamic window's base from the
// generic pointer around every predicated store: four extra instructions per access in the first versions)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_f32_if(uint32_t addr, float v, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %2, 0;\n"
      "@p st.shared.f32 [%0], %1;\n"
      "}\n" ::"r"(addr),
      "f"(v), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_stream_u16(const uint16_t* p) {  // zero-extended straight into a 32-bit register:
  uint32_t v;                                                            // a 16-bit destination put a conversion right
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=r"(v) : "l"(p));  // behind every load and serialised them
  return v;
}

// slot j of the panel's per-row order lives at word j + j / 32: thread t of the sweep owns slots [32 t, 32 t + 32), and
// the one-word skew per 32 slots keeps the 32 lanes of a warp on 32 different banks
__device__ __forceinline__ uint32_t skew(uint32_t j) { return j + (j >> 5); }

// 64-bit streaming load of four 16-bit places
__device__ __forceinline__ uint2 ldg_stream_u2x32(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 6 : 2)
    pb_reduce_kernel(PbPlan P, float* __restrict__ y, Epilogue ep) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  // STREAM panel: [33 * nwords] the panel's partials in per-row order (skewed)
  float* s_prod = reinterpret_cast<float*>(s_raw);
  const uint32_t sp = smem_u32(s_raw);
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[WARPS];
  __shared__ float s_wv[WARPS];
  __shared__ int s_wf[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t t = P.panel_begin + blockIdx.x;
  const TileDesc d = load_desc(P.desc + t);
  const int n = d.n1 - d.n0;
  const bool is_long = d.chunk >= 0;
  const int trows = d.r1 - d.r0;

  if (is_long) {
    // ---- a chunk of a LONG row: every piece goes into one sum.  Warp w takes run w, w + WARPS, ... of the chunk table,
    // U runs in flight, the next U descriptors already requested.  Idle lanes of a short run issue nothing (sending
    // them to one dummy address made that L2 line a hot spot).
    const int ch0 = __ldg(P.panel_chunk + t);
    const int nch = __ldg(P.panel_chunk + t + 1) - ch0;
    const int2* __restrict__ g_chunk = P.chunk + ch0;
    constexpr int U = 4;
    int2 nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ci = warp + u * WARPS;
      nxt[u] = ci < nch ? __ldg(g_chunk + ci) : make_int2(0, 0);
    }
    const uint64_t part_base = reinterpret_cast<uint64_t>(P.part + lane);
    float acc = 0.0f;
    for (int cb = warp; cb < nch; cb += WARPS * U) {
      int2 cur[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cur[u] = nxt[u];
        const int ci = cb + (u + U) * WARPS;
        nxt[u] = ci < nch ? __ldg(g_chunk + ci) : make_int2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lane < cur[u].y)
          acc += ldg_stream_f32(reinterpret_cast<const float*>(part_base + ((uint64_t)(uint32_t)cur[u].x << 2)));
    }
    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp != 0) return;
    float total = lane < WARPS ? s_red[lane] : 0.0f;
    total = warp_sum(total);
    finish_chunk(P.carry, P.counter, d, t, total, lane, y, ep);
    return;
  }

  // ---- STREAM panel.  Gather: the panel's pieces are one contiguous run of part[] per column slab.  The plan lists the
  // 16-byte-aligned quads of partial sums that cover those runs in panel-major order (chunk_src) next to the slots of
  // their four values (perm2, 0xFFFF = alignment padding), so the gather is flat: thread q takes quad q, q + THREADS,
  // ...: one index, one 128-bit load, four predicated shared-memory stores -- no run descriptors, no per-run
  // instructions (warps walking a run table: 33 M warp instructions on C2; one bulk copy per run: 27 M, the copies are
  // issued lane by lane through the uniform datapath; this: 5 M).
  const int nwords = (n + 31) >> 5;
  const int2 aux0 = __ldg(P.panel_aux + t), aux1 = __ldg(P.panel_aux + t + 1);
  const int nq = (aux1.x - aux0.x) >> 2;               // staged quads
  constexpr int A = THREADS == 256 ? 6 : 4;  // quads per thread whose index and slots are requested up front
  const uint16_t* perm2 = P.perm2 + aux0.x;
  const int32_t* __restrict__ csrc = P.chunk_src + (aux0.x >> 2);
  int src[A];
  uint2 places[A];
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const int q4 = tid + a * THREADS;
    src[a] = q4 < nq ? __ldg(csrc + q4) : 0;
    places[a] = q4 < nq ? ldg_stream_u2x32(perm2 + 4 * q4) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
  }
  const uint32_t* __restrict__ g_bits = P.end_bits + aux0.y;
  // thread w owns word w of the slots (the launcher picks THREADS >= nwords): its end marks and the last bit of the word before
  const uint32_t my_bits = tid < nwords ? __ldg(g_bits + tid) : 0u;
  const uint32_t prev_bits = (tid < nwords && tid > 0) ? __ldg(g_bits + tid - 1) : 0x80000000u;
  constexpr int RA = THREADS == 256 ? 3 : kRowsAhead;
  float bias_pre[RA];
  int last_slot[RA];  // the slot that holds the row's total after the sweep, -1 for a row without pieces
#pragma unroll
  for (int a = 0; a < RA; ++a) {  // their DRAM round trips overlap everything up to the epilogue
    const int i = tid + a * THREADS;
    bias_pre[a] = (ep.beta != 0.0f && i < trows) ? ep.bias[d.r0 + i] : 0.0f;
    const int b = i < trows ? __ldg(P.prow_ptr + d.r0 + i) : 0;
    const int e = i < trows ? __ldg(P.prow_ptr + d.r0 + i + 1) : 0;
    last_slot[a] = e > b ? e - 1 - d.n0 : -1;
  }
  for (int i = tid + RA * THREADS; i < trows; i += THREADS) {  // rows beyond those: at least have the lines on their way
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.prow_ptr + d.r0 + i));
    if (ep.beta != 0.0f) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.bias + d.r0 + i));
  }
  if (tid < 32 && n + tid < 32 * nwords) s_prod[skew((uint32_t)(n + tid))] = 0.0f;  // the last word's unused slots
  const float* __restrict__ part = P.part;
  const uint64_t ps = policy_evict_first();
  auto drop4 = [&](const float4& v, uint2 pl) {  // four partial sums -> their slots in the panel's per-row order
    const uint32_t p0 = pl.x & 0xFFFFu, p1 = pl.x >> 16, p2 = pl.y & 0xFFFFu, p3 = pl.y >> 16;
    sts_f32_if(sp + 4u * skew(p0), v.x, p0 != 0xFFFFu);
    sts_f32_if(sp + 4u * skew(p1), v.y, p1 != 0xFFFFu);
    sts_f32_if(sp + 4u * skew(p2), v.z, p2 != 0xFFFFu);
    sts_f32_if(sp + 4u * skew(p3), v.w, p3 != 0xFFFFu);
  };
  constexpr int V = 3;  // 128-bit loads in flight per thread
#pragma unroll
  for (int a0 = 0; a0 < A; a0 += V) {
    float4 v[V];
#pragma unroll
    for (int a = a0; a < a0 + V && a < A; ++a)
      v[a - a0] = tid + a * THREADS < nq ? ld_stream_f4(part + src[a], ps) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = a0; a < a0 + V && a < A; ++a) drop4(v[a - a0], places[a]);
  }
  for (int q4 = tid + A * THREADS; q4 < nq; q4 += THREADS)
    drop4(ld_stream_f4(part + __ldg(csrc + q4), ps), ldg_stream_u2x32(perm2 + 4 * q4));
  __syncthreads();

  // ---- reduce: thread w sums the 32 slots of word w of the per-row order, closing rows at the marks ----------------
  float sv = 0.0f;         // what the thread leaves open: the tail after its last end, or its whole word
  unsigned sf = 0;         // thread closed a row
  float lead = 0.0f;       // my share of a row that began in an earlier word and ends in mine ...
  int lead_slot = -1;      // ... at this (skewed) slot
  if (tid < nwords) {
    const int w = tid;
    const uint32_t bits = my_bits;
    bool pending = w > 0 && !(prev_bits >> 31);   // my first slot's row began before this word
    const uint32_t base = sp + 4u * (33u * (uint32_t)w);
    float run = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {  // straight-line: selects and one predicated store per slot, no branches
      run += lds_f32(base + 4u * k);
      const bool end = (bits >> k) & 1u;
      const bool mine = end && pending;
      lead = mine ? run : lead;
      lead_slot = mine ? 33 * w + k : lead_slot;
      sts_f32_if(base + 4u * k, run, end && !pending);  // the row's total waits in its last slot for the epilogue
      pending = pending && !end;
      run = end ? 0.0f : run;
    }
    sv = run;
    sf = bits != 0u;
  }
  // segmented scan over the threads of (open sum, closed a row): carry = the open row's sum over the threads before
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const float uv = __shfl_up_sync(kFullMask, sv, dd);
    const unsigned uf = __shfl_up_sync(kFullMask, sf, dd);
    if (lane >= dd) {
      if (!sf) sv += uv;
      sf |= uf;
    }
  }
  if (lane == 31) {
    s_wv[warp] = sv;
    s_wf[warp] = (int)sf;
  }
  float carry = __shfl_up_sync(kFullMask, sv, 1);
  const unsigned before = __shfl_up_sync(kFullMask, sf, 1);  // some earlier lane of this warp closed a row
  __syncthreads();
  if (lead_slot >= 0) {
    float c_in = 0.0f;
    bool need_warps = true;
    if (lane > 0) {
      c_in = carry;
      need_warps = !before;
    }
    if (need_warps) {  // add the open sums of the warps before this one, back to the last warp that closed a row
      float wsum = 0.0f;
      int w0 = warp;
      while (w0 > 0) {
        --w0;
        if (s_wf[w0]) break;
      }
      // w0 is the last closing warp before this one (or 0): its tail, then every pass-through warp after it, in order
      for (int w = w0; w < warp; ++w) wsum += s_wv[w];
      c_in = wsum + c_in;
    }
    s_prod[lead_slot] = c_in + lead;
  }
  __syncthreads();

  // ---- epilogue: one thread per row, coalesced ------------------------------------------------------------------
  auto finish_row = [&](int i, float bias, int slot) {
    const float sum = slot >= 0 ? s_prod[skew((uint32_t)slot)] : 0.0f;
    float v = ep.alpha * sum;
    if (ep.beta != 0.0f) v = fmaf(ep.beta, bias, v);
    if (ep.relu) v = fmaxf(v, 0.0f);
    store_y(y, d.r0 + i, v, ep.y_mc);
  };
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    const int i = tid + a * THREADS;
    if (i < trows) finish_row(i, bias_pre[a], last_slot[a]);
  }
  for (int i = tid + RA * THREADS; i < trows; i += THREADS) {
    const int b = __ldg(P.prow_ptr + d.r0 + i), e = __ldg(P.prow_ptr + d.r0 + i + 1);
    finish_row(i, ep.beta != 0.0f ? ep.bias[d.r0 + i] : 0.0f, e > b ? e - 1 - d.n0 : -1);
  }
}

}  // namespace

template <int THREADS>
int launch_pb_expand_t(const PbPlan& P, int32_t cols, const float* x, size_t smem, cudaStream_t s) {
  static size_t configured = 0;
  if (smem > configured) {
    HISPMV_CUDA(cudaFuncSetAttribute(pb_expand_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  static const bool debug = getenv("HISPMV_PB_DEBUG") != nullptr;
  if (debug) {  // development: how long every CTA was busy (static ranges: the slowest one is the kernel's duration)
    static int calls = 0;
    PbPlan Q = P;
    HISPMV_CUDA(cudaMallocManaged((void**)&Q.dbg, (size_t)P.num_work * 3 * sizeof(long long)));
    pb_expand_kernel<THREADS><<<P.num_work, THREADS, smem, s>>>(Q, x, cols);
    HISPMV_CUDA(cudaStreamSynchronize(s));
    if (++calls == 5) {
      long long mx = 0, sum = 0;
      for (int i = 0; i < P.num_work; ++i) {
        mx = std::max(mx, Q.dbg[3 * i]);
        sum += Q.dbg[3 * i];
      }
      fprintf(stderr, "pb_expand: %d CTAs, busy ns avg %.0f max %lld\n", P.num_work, (double)sum / P.num_work, mx);
      for (int i = 0; i < P.num_work; ++i) {
        const long long groups = Q.dbg[3 * i + 2] & 0xffffffffll, pieces = Q.dbg[3 * i + 2] >> 32;
        fprintf(stderr, "  cta %3d  ns %7lld  slab loads %3lld  groups %6lld  pieces %8lld  ns/group %.1f\n", i, Q.dbg[3 * i],
                Q.dbg[3 * i + 1], groups, pieces, (double)Q.dbg[3 * i] / (double)std::max<long long>(1, groups));
      }
    }
    cudaFree(Q.dbg);
    return HISPMV_OK;
  }
  pb_expand_kernel<THREADS><<<P.num_work, THREADS, smem, s>>>(P, x, cols);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

// The kernel is bound by the latency of its shared-memory gathers, so it runs as many warps as fit next to the x slice:
// every warp needs a 2 KB stage for its pieces (HISPMV_PB_EXPAND_THREADS overrides).
int launch_pb_expand(const PbPlan& P, int32_t cols, const float* x, cudaStream_t s) {
  if (P.num_work <= 0) return HISPMV_OK;
  static const int forced = getenv("HISPMV_PB_EXPAND_THREADS") ? atoi(getenv("HISPMV_PB_EXPAND_THREADS")) : 0;
  const size_t room = (size_t)227 * 1024 - (size_t)P.slab_cols * 4;
  int threads = 512;
  if (room >= 24 * kPbGroup * 4) threads = 768;
  else if (room >= 20 * kPbGroup * 4) threads = 640;
  if (forced == 512 || forced == 640 || forced == 768) threads = std::min(threads, forced);
  const size_t smem = (size_t)P.slab_cols * 4 + (size_t)(threads / 32) * kPbGroup * 4;
  if (threads == 768) return launch_pb_expand_t<768>(P, cols, x, smem, s);
  if (threads == 640) return launch_pb_expand_t<640>(P, cols, x, smem, s);
  return launch_pb_expand_t<512>(P, cols, x, smem, s);
}

int launch_pb_reduce(const CsrDev& A, const PbPlan& P, float* y, Epilogue ep, cudaStream_t s) {
  (void)A;
  const int64_t count = P.panel_count < 0 ? P.num_panels - P.panel_begin : P.panel_count;
  if (count <= 0) return HISPMV_OK;
  // the largest STREAM panel: slots skewed by one word per 32 (pb_reduce_words_kernel)
  const size_t smem = ((size_t)P.reduce_words + 8) * 4;
  if (smem > 227 * 1024) {
    set_error("blocked plan: a panel does not fit shared memory");
    return HISPMV_ERR_STATE;
  }
  // One thread sweeps one 32-slot word, so a panel may hold at most 32 * THREADS slots (cap_words bounds every panel).
  const int max_slots = P.cap_words;
  if (max_slots > 32 * kReduceThreads) {
    set_error("blocked plan: panel_items + long_threshold must not exceed 16384");
    return HISPMV_ERR_STATE;
  }
  // panels of up to 8192 slots run with 256 threads, six CTAs per SM: every panel is a chain of three dependent memory
  // round trips (descriptor -> quad indices -> partial sums), so what counts is how many panels an SM has in flight
  // (HISPMV_PB_THREADS=512 overrides)
  static const int forced = getenv("HISPMV_PB_THREADS") ? atoi(getenv("HISPMV_PB_THREADS")) : 0;
  const bool fits256 = max_slots <= 32 * 256;
  const bool small = fits256 && forced != 512;  // <= 34 KB of slots: six CTAs of 256 threads per SM
  static size_t configured[2] = {0, 0};
  if (smem > configured[small]) {
    if (small)
      HISPMV_CUDA(cudaFuncSetAttribute(pb_reduce_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      HISPMV_CUDA(cudaFuncSetAttribute(pb_reduce_kernel<kReduceThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    configured[small] = smem;
  }
  if (small)
    pb_reduce_kernel<256><<<(unsigned)count, 256, smem, s>>>(P, y, ep);
  else
    pb_reduce_kernel<kReduceThreads><<<(unsigned)count, kReduceThreads, smem, s>>>(P, y, ep);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

// Selector input: how many (row, slab) runs the matrix has = the number of pieces before group boundaries split any.
// runs = non-empty rows + slab changes between neighbouring entries - the changes that fall on a row start: two flat,
// coalesced passes (one warp walking each row took 16 ms on C2, whose longest rows hold a million entries).
namespace {
__global__ void pb_runs_rows_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ col, int32_t rows, int32_t W,
                                    unsigned long long* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = 0;  // +1 for a non-empty row, -1 when the slab change at its first entry was counted by the entry pass
  if (r < rows) {
    const int b = rp[r], e = rp[r + 1];
    if (e > b) v = 1 - ((b > 0 && col[b] / W != col[b - 1] / W) ? 1 : 0);
  }
  v = __reduce_add_sync(kFullMask, v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)(long long)v);
}
__global__ void pb_runs_entries_kernel(const int32_t* __restrict__ col, int64_t nnz, int32_t W,
                                       unsigned long long* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned int mine = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; j < nnz; j += stride)
    mine += col[j] / W != col[j - 1] / W ? 1u : 0u;
  mine = __reduce_add_sync(kFullMask, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, (unsigned long long)mine);
}
}  // namespace

int pb_count_runs_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, int32_t slab_cols, int64_t* runs,
                         cudaStream_t stream) {
  *runs = 0;
  if (rows <= 0) return HISPMV_OK;
  DevBuf b;
  int st;
  if ((st = b.alloc(sizeof(unsigned long long)))) return st;
  HISPMV_CUDA(cudaMemsetAsync(b.p, 0, sizeof(unsigned long long), stream));
  int32_t nnz = 0;
  HISPMV_CUDA(cudaMemcpyAsync(&nnz, d_row_ptr + rows, 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  pb_runs_rows_kernel<<<blocks_for(rows, 256), 256, 0, stream>>>(d_row_ptr, d_col, rows, slab_cols,
                                                                 b.as<unsigned long long>());
  if (nnz > 1) {
    const int grid = (int)std::min<int64_t>(blocks_for(nnz, 256), 148 * 32);
    pb_runs_entries_kernel<<<grid, 256, 0, stream>>>(d_col, nnz, slab_cols, b.as<unsigned long long>());
  }
  HISPMV_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  HISPMV_CUDA(cudaMemcpyAsync(&h, b.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  *runs = (int64_t)h;
  return HISPMV_OK;
}

// The blocked strategy streams 6.25 bytes per nonzero in pass 1 and about 14 bytes per PIECE over both passes, instead
// of 8 bytes per nonzero plus a 32-byte L2 sector per scattered gather.  It wins when the gathers are scattered (not
// banded), x is far larger than an SM's L1, the matrix is large enough to fill two launches, and the rows are
// concentrated enough that pieces are few.  Measured on B200 (profiles/r2_blocked_crossover.txt, 100 M nonzeros): about
// 1.5 ns per nonzero + 3.2 ns per piece against 3.7-4.3 ns per nonzero for the one-pass kernel: 1.5x faster from 0.30
// to 0.40 runs per nonzero (power-law rows, any column skew), 7 % slower at 1.0 (uniform rows and columns: every
// nonzero its own piece); the break-even is near 0.8.  Rule: runs <= 0.6 nnz.  The column threshold is where x leaves an
// SM's L1: from 100 000 to 10 M columns the ratio stays 1.5-1.6 (profiles/r2_blocked_cols_sweep.txt).  Integer rule,
// restated in oracle/oracle.c (oracle_select_blocked).
int select_blocked(int32_t rows, int32_t cols, int64_t nnz, int64_t slab_runs, const ColProbe& probe,
                   int allow_split_rows) {
  const bool banded = probe.cmp >= 64 && probe.near * 4 >= probe.cmp * 3;
  if (!allow_split_rows || banded || rows <= 0) return 0;
  if ((int64_t)cols < 100000 || nnz < 16000000) return 0;
  if (((int64_t)cols + kPbSlabCols - 1) / kPbSlabCols > 4096) return 0;
  return slab_runs * 5 <= nnz * 3 ? 1 : 0;
}

}  // namespace hispmv
