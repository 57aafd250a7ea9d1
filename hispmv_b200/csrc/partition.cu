// GPU-side partitioner: COO -> CSR, row-length histogram, runtime kernel selector, merge-path tile
// descriptors with heavy rows split across CTAs, nnz-balanced row blocks for multi-GPU.
//
// Replaces the reference's host preprocessing (semantics, not structure):
//   tileAndPad(COO)     count / prefix / bucket / sort rows by column    common/src/spmv-helper.cpp:139-227
//   cooToCsr            the same as a flat CSR                           gpu/src/spmvHelper.cpp:117-156
//   balanceWorkload     which rows must be shared between PEs            common/src/spmv-helper.cpp:265-347
//   computeTileSize / prepareTile  the balanced schedule                 common/src/spmv-helper.cpp:429-638
//   DSE.getBestConfig   per-matrix configuration choice                  automation_tool/src/dse.py:23-95
//
// Integer outputs (row_ptr, col order, tile coordinates, split-row list, shard bounds, kernel choice)
// are bit-exact against the single-threaded restatement in oracle/oracle.c.
// CUB (shipped with the CUDA toolkit) provides the radix sort and the prefix sum; everything else is
// written here.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "internal.h"

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

// order-preserving map float -> uint32 (so that an unsigned radix sort orders like operator< on float)
__device__ __forceinline__ uint32_t float_order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void iota_and_valkey_kernel(const float* __restrict__ vals, int64_t nnz, uint32_t* __restrict__ key,
                                       uint32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    if (key) key[i] = float_order_key(vals[i]);
    idx[i] = (uint32_t)i;
  }
}

// key64 = row:col of entry idx[i]; also validates the index ranges (the reference does not; we refuse)
__global__ void make_rowcol_key_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                                       const uint32_t* __restrict__ idx, int64_t nnz, int32_t n_rows, int32_t n_cols,
                                       uint64_t* __restrict__ key, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    const uint32_t e = idx ? idx[i] : (uint32_t)i;
    const int32_t r = rows[e], c = cols[e];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) *bad = 1;
    key[i] = ((uint64_t)(uint32_t)r << 32) | (uint32_t)c;
  }
}

__global__ void has_adjacent_dup_kernel(const uint64_t* __restrict__ key, int64_t nnz, int* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < nnz && key[i] == key[i + 1]) *flag = 1;
}

__global__ void scatter_sorted_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ idx,
                                      const float* __restrict__ vals, int64_t nnz, int64_t padded,
                                      int32_t* __restrict__ col, float* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    col[i] = (int32_t)(uint32_t)(key[i] & 0xffffffffu);
    val[i] = vals[idx[i]];
  } else if (i < padded) {
    col[i] = 0;
    val[i] = 0.0f;
  }
}

// row_ptr[r] = number of sorted entries whose row is < r  (lower bound on the high half of the key)
__global__ void row_ptr_from_sorted_kernel(const uint64_t* __restrict__ key, int64_t nnz, int32_t rows,
                                           int32_t* __restrict__ row_ptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > rows) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)(key[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  row_ptr[r] = (int32_t)lo;
}

__global__ void row_stats_kernel(const int32_t* __restrict__ row_ptr, int32_t rows,
                                 unsigned long long* __restrict__ hist, int* __restrict__ max_len) {
  __shared__ unsigned int s_hist[HISPMV_HIST_BINS];
  __shared__ int s_max;
  if (threadIdx.x < HISPMV_HIST_BINS) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int local_max = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const int len = row_ptr[r + 1] - row_ptr[r];
    const int bin = len <= 0 ? 0 : 32 - __clz(len);
    atomicAdd(&s_hist[bin], 1u);
    local_max = max(local_max, len);
  }
  atomicMax(&s_max, local_max);
  __syncthreads();
  if (threadIdx.x < HISPMV_HIST_BINS && s_hist[threadIdx.x])
    atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
  if (threadIdx.x == 0) atomicMax(max_len, s_max);
}

// Merge-path diagonal search over (row ends) x (nonzero indices); see oracle_merge_tiles.
__global__ void merge_tiles_kernel(const int32_t* __restrict__ row_ptr, int32_t rows, int64_t nnz, int32_t tile_items,
                                   int64_t num_tiles, int32_t* __restrict__ tile_row, int64_t* __restrict__ tile_nnz) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > num_tiles) return;
  const int64_t total = (int64_t)rows + nnz;
  int64_t diag = t * (int64_t)tile_items;
  if (diag > total) diag = total;
  int64_t lo = diag > nnz ? diag - nnz : 0;
  int64_t hi = diag < rows ? diag : rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_ptr[mid + 1] <= diag - mid - 1) lo = mid + 1; else hi = mid;
  }
  tile_row[t] = (int32_t)lo;
  tile_nnz[t] = diag - lo;
}

// flag[t-1] = 1 when boundary t (1 <= t < num_tiles) is the first one that cuts row tile_row[t]
// after at least one of its nonzeros.
__global__ void split_flag_kernel(const int32_t* __restrict__ row_ptr, int32_t rows,
                                  const int32_t* __restrict__ tile_row, const int64_t* __restrict__ tile_nnz,
                                  int64_t num_tiles, int32_t* __restrict__ flag) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (t >= num_tiles) return;
  const int32_t r = tile_row[t];
  int f = 0;
  if (r < rows && tile_nnz[t] > (int64_t)row_ptr[r]) {
    f = 1;
    if (t > 1 && tile_row[t - 1] == r && tile_nnz[t - 1] > (int64_t)row_ptr[r]) f = 0;
  }
  flag[t - 1] = f;
}

__global__ void split_scatter_kernel(const int32_t* __restrict__ tile_row, const int32_t* __restrict__ flag,
                                     const int32_t* __restrict__ pos, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = tile_row[i + 1];
}

__global__ void rebase_row_ptr_kernel(const int32_t* __restrict__ src, int32_t row_begin, int32_t n_rows,
                                      int32_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) dst[i] = src[row_begin + i] - src[row_begin];
}

__global__ void zero_pad_kernel(int32_t* col, float* val, int64_t nnz, int64_t padded) {
  const int64_t i = nnz + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < padded) {
    col[i] = 0;
    val[i] = 0.0f;
  }
}

__global__ void shard_bounds_kernel(const int32_t* __restrict__ row_ptr, int32_t rows, int64_t nnz, int n_parts,
                                    int32_t* __restrict__ bounds) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_parts) return;
  if (k == 0) { bounds[0] = 0; return; }
  if (k == n_parts) { bounds[k] = rows; return; }
  const int64_t target = (nnz * (int64_t)k) / n_parts;
  int64_t lo = 0, hi = rows;  // first row index r with row_ptr[r] >= target
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  bounds[k] = (int32_t)lo;
}

__global__ void col_probe_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ col, int32_t rows,
                                 int32_t samples, unsigned long long* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int near = 0, cmp = 0;
  if (j < samples) {
    const int64_t r = 1 + ((int64_t)j * (rows - 1)) / samples;
    const int32_t a0 = rp[r - 1], a1 = rp[r], a2 = rp[r + 1];
    int n = min(a1 - a0, a2 - a1);
    n = min(n, 32);
    for (int k = 0; k < n; ++k) {
      const int32_t d = col[a1 + k] - col[a0 + k];
      near += (d <= 32 && d >= -32);
    }
    cmp = n > 0 ? n : 0;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    near += __shfl_xor_sync(0xffffffffu, near, d);
    cmp += __shfl_xor_sync(0xffffffffu, cmp, d);
  }
  if ((threadIdx.x & 31) == 0 && cmp) {
    atomicAdd(out, (unsigned long long)near);
    atomicAdd(out + 1, (unsigned long long)cmp);
  }
}

inline int64_t padded_nnz(int64_t nnz) { return ((nnz + 3) & ~(int64_t)3) + 4; }

}  // namespace

int alloc_padded_nnz_arrays(const int32_t* src_col, const float* src_val, int64_t nnz, cudaMemcpyKind kind,
                            int32_t** d_col, float** d_val, cudaStream_t stream) {
  const int64_t pad = padded_nnz(nnz);
  *d_col = nullptr;
  *d_val = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)d_col, pad * sizeof(int32_t)));
  int st = check_cuda(cudaMalloc((void**)d_val, pad * sizeof(float)), "cudaMalloc(val)", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(*d_col);
    *d_col = nullptr;
    return st;
  }
  if (nnz > 0) {
    HISPMV_CUDA(cudaMemcpyAsync(*d_col, src_col, nnz * sizeof(int32_t), kind, stream));
    HISPMV_CUDA(cudaMemcpyAsync(*d_val, src_val, nnz * sizeof(float), kind, stream));
  }
  zero_pad_kernel<<<1, 32, 0, stream>>>(*d_col, *d_val, nnz, pad);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int coo_to_csr_device(const int32_t* d_rows, const int32_t* d_cols, const float* d_vals, int64_t nnz, int32_t rows,
                      int32_t cols, int32_t** d_row_ptr, int32_t** d_col, float** d_val, cudaStream_t stream) {
  *d_row_ptr = nullptr;
  *d_col = nullptr;
  *d_val = nullptr;
  if (nnz >= (int64_t)INT32_MAX) {
    set_error("coo_to_csr: nnz must be below 2^31 per GPU (int32 row_ptr, as in the reference)");
    return HISPMV_ERR_ARG;
  }
  const int64_t pad = padded_nnz(nnz);
  DevBuf key_a, key_b, idx_a, idx_b, vkey_a, vkey_b, tmp, flags;
  int st;
  if ((st = key_a.alloc(nnz * 8)) || (st = key_b.alloc(nnz * 8)) || (st = idx_a.alloc(nnz * 4)) ||
      (st = idx_b.alloc(nnz * 4)) || (st = flags.alloc(2 * sizeof(int))))
    return st;
  HISPMV_CUDA(cudaMemsetAsync(flags.p, 0, 2 * sizeof(int), stream));
  int* d_bad = flags.as<int>();
  int* d_dup = flags.as<int>() + 1;
  const int B = 256;

  cub::DoubleBuffer<uint64_t> keys(key_a.as<uint64_t>(), key_b.as<uint64_t>());
  cub::DoubleBuffer<uint32_t> idx(idx_a.as<uint32_t>(), idx_b.as<uint32_t>());
  int end_bit = 64;  // only sort the bits that can be set
  {
    int rb = 1, cb = 1;
    while (rb < 32 && (1LL << rb) < (int64_t)rows) ++rb;
    while (cb < 32 && (1LL << cb) < (int64_t)cols) ++cb;
    end_bit = 32 + rb;
    (void)cb;
  }

  if (nnz > 0) {
    // pass 1: sort by (row, col) only
    iota_and_valkey_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(d_vals, nnz, nullptr, idx.Current());
    make_rowcol_key_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(d_rows, d_cols, nullptr, nnz, rows, cols,
                                                                 keys.Current(), d_bad);
    HISPMV_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, idx, nnz, 0, end_bit, stream));
    if ((st = tmp.alloc(tmp_bytes))) return st;
    HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys, idx, nnz, 0, end_bit, stream));
    has_adjacent_dup_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(keys.Current(), nnz, d_dup);
    int h_flags[2];
    HISPMV_CUDA(cudaMemcpyAsync(h_flags, flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
    HISPMV_CUDA(cudaStreamSynchronize(stream));
    if (h_flags[0]) {
      set_error("coo_to_csr: a row or column index is outside [0,rows) x [0,cols)");
      return HISPMV_ERR_ARG;
    }
    if (h_flags[1]) {
      // Duplicated (row, col) pairs: the reference orders them by value (std::sort over (col, val) pairs,
      // spmv-helper.cpp:216).  Redo as an LSD sort: value key first, then a stable sort on (row, col).
      if ((st = vkey_a.alloc(nnz * 4)) || (st = vkey_b.alloc(nnz * 4))) return st;
      cub::DoubleBuffer<uint32_t> vkeys(vkey_a.as<uint32_t>(), vkey_b.as<uint32_t>());
      iota_and_valkey_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(d_vals, nnz, vkeys.Current(), idx.Current());
      size_t tb2 = 0;
      HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, vkeys, idx, nnz, 0, 32, stream));
      if (tb2 > tmp_bytes) {
        if ((st = tmp.alloc(tb2))) return st;
        tmp_bytes = tb2;
      }
      HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb2, vkeys, idx, nnz, 0, 32, stream));
      make_rowcol_key_kernel<<<blocks_for(nnz, B), B, 0, stream>>>(d_rows, d_cols, idx.Current(), nnz, rows, cols,
                                                                   keys.Current(), d_bad);
      size_t tb3 = 0;
      HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb3, keys, idx, nnz, 0, end_bit, stream));
      if (tb3 > tmp_bytes) {
        if ((st = tmp.alloc(tb3))) return st;
        tmp_bytes = tb3;
      }
      HISPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb3, keys, idx, nnz, 0, end_bit, stream));
    }
  }

  HISPMV_CUDA(cudaMalloc((void**)d_row_ptr, ((size_t)rows + 1 + 4) * sizeof(int32_t)));  // +4: TMA windows
  st = check_cuda(cudaMalloc((void**)d_col, pad * sizeof(int32_t)), "cudaMalloc(col)", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)d_val, pad * sizeof(float)), "cudaMalloc(val)", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(*d_row_ptr);
    cudaFree(*d_col);
    *d_row_ptr = nullptr;
    *d_col = nullptr;
    return st;
  }
  scatter_sorted_kernel<<<blocks_for(pad, B), B, 0, stream>>>(keys.Current(), idx.Current(), d_vals, nnz, pad, *d_col,
                                                              *d_val);
  row_ptr_from_sorted_kernel<<<blocks_for((int64_t)rows + 1, B), B, 0, stream>>>(keys.Current(), nnz, rows,
                                                                                 *d_row_ptr);
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaStreamSynchronize(stream));  // scratch buffers die at scope exit
  return HISPMV_OK;
}

int col_probe_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, ColProbe* out, cudaStream_t stream) {
  out->near = out->cmp = 0;
  if (rows < 2) return HISPMV_OK;
  const int32_t samples = std::min<int32_t>(rows - 1, kProbeSamples);
  DevBuf buf;
  int st;
  if ((st = buf.alloc(2 * sizeof(unsigned long long)))) return st;
  HISPMV_CUDA(cudaMemsetAsync(buf.p, 0, 2 * sizeof(unsigned long long), stream));
  col_probe_kernel<<<blocks_for(samples, 128), 128, 0, stream>>>(d_row_ptr, d_col, rows, samples,
                                                                 buf.as<unsigned long long>());
  HISPMV_CUDA(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(h, buf.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  out->near = (int64_t)h[0];
  out->cmp = (int64_t)h[1];
  return HISPMV_OK;
}

int row_stats_device(const int32_t* d_row_ptr, int32_t rows, RowStats* out, cudaStream_t stream) {
  DevBuf buf;
  int st;
  const size_t bytes = HISPMV_HIST_BINS * sizeof(unsigned long long) + sizeof(int) * 2;
  if ((st = buf.alloc(bytes))) return st;
  HISPMV_CUDA(cudaMemsetAsync(buf.p, 0, bytes, stream));
  unsigned long long* d_hist = buf.as<unsigned long long>();
  int* d_max = reinterpret_cast<int*>(d_hist + HISPMV_HIST_BINS);
  if (rows > 0) {
    const int grid = (int)std::min<int64_t>(blocks_for(rows, 256), 148 * 16);
    row_stats_kernel<<<grid, 256, 0, stream>>>(d_row_ptr, rows, d_hist, d_max);
    HISPMV_CUDA(cudaGetLastError());
  }
  unsigned long long h_hist[HISPMV_HIST_BINS];
  int h_max = 0;
  int32_t h_last = 0;
  HISPMV_CUDA(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h_last, d_row_ptr + rows, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  for (int i = 0; i < HISPMV_HIST_BINS; ++i) out->hist[i] = (int64_t)h_hist[i];
  out->rows = rows;
  out->nnz = h_last;
  out->max_row_nnz = h_max;
  out->empty_rows = (int32_t)h_hist[0];
  return HISPMV_OK;
}

int merge_tiles_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int32_t tile_items, int64_t* num_tiles,
                       int32_t** d_tile_row, int64_t** d_tile_nnz, cudaStream_t stream) {
  const int64_t total = (int64_t)rows + nnz;
  const int64_t nt = std::max<int64_t>(1, (total + tile_items - 1) / tile_items);
  *num_tiles = nt;
  *d_tile_row = nullptr;
  *d_tile_nnz = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)d_tile_row, (nt + 1) * sizeof(int32_t)));
  int st = check_cuda(cudaMalloc((void**)d_tile_nnz, (nt + 1) * sizeof(int64_t)), "cudaMalloc(tile_nnz)", __FILE__,
                      __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(*d_tile_row);
    *d_tile_row = nullptr;
    return st;
  }
  merge_tiles_kernel<<<blocks_for(nt + 1, 256), 256, 0, stream>>>(d_row_ptr, rows, nnz, tile_items, nt, *d_tile_row,
                                                                  *d_tile_nnz);
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int split_rows_device(const int32_t* d_row_ptr, int32_t rows, const int32_t* d_tile_row, const int64_t* d_tile_nnz,
                      int64_t num_tiles, int32_t** d_out, int64_t* count, cudaStream_t stream) {
  *d_out = nullptr;
  *count = 0;
  const int64_t n = num_tiles - 1;  // interior boundaries
  if (n <= 0) return HISPMV_OK;
  DevBuf flag, pos, tmp;
  int st;
  if ((st = flag.alloc(n * sizeof(int32_t))) || (st = pos.alloc(n * sizeof(int32_t)))) return st;
  split_flag_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(d_row_ptr, rows, d_tile_row, d_tile_nnz, num_tiles,
                                                            flag.as<int32_t>());
  size_t tb = 0;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.as<int32_t>(), pos.as<int32_t>(), n, stream));
  if ((st = tmp.alloc(tb))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.as<int32_t>(), pos.as<int32_t>(), n, stream));
  int32_t last_pos = 0, last_flag = 0;
  HISPMV_CUDA(cudaMemcpyAsync(&last_pos, pos.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&last_flag, flag.as<int32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t cnt = (int64_t)last_pos + last_flag;
  *count = cnt;
  if (cnt > 0) {
    HISPMV_CUDA(cudaMalloc((void**)d_out, cnt * sizeof(int32_t)));
    split_scatter_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(d_tile_row, flag.as<int32_t>(), pos.as<int32_t>(), n,
                                                                 *d_out);
    HISPMV_CUDA(cudaGetLastError());
    HISPMV_CUDA(cudaStreamSynchronize(stream));
  }
  return HISPMV_OK;
}

int csr_slice_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t row_begin,
                     int32_t row_end, int32_t** o_row_ptr, int32_t** o_col, float** o_val, int64_t* o_nnz,
                     cudaStream_t stream) {
  int32_t h[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(&h[0], d_row_ptr + row_begin, 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h[1], d_row_ptr + row_end, 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t nnz = (int64_t)h[1] - h[0];
  const int32_t n_rows = row_end - row_begin;
  *o_nnz = nnz;
  *o_row_ptr = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)o_row_ptr, ((size_t)n_rows + 1 + 4) * sizeof(int32_t)));  // +4: TMA windows
  rebase_row_ptr_kernel<<<blocks_for((int64_t)n_rows + 1, 256), 256, 0, stream>>>(d_row_ptr, row_begin, n_rows,
                                                                                  *o_row_ptr);
  int st = alloc_padded_nnz_arrays(d_col + h[0], d_val + h[0], nnz, cudaMemcpyDeviceToDevice, o_col, o_val, stream);
  if (st != HISPMV_OK) {
    cudaFree(*o_row_ptr);
    *o_row_ptr = nullptr;
    return st;
  }
  HISPMV_CUDA(cudaGetLastError());
  return HISPMV_OK;
}

int shard_bounds_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int n_parts, int32_t* h_bounds,
                        cudaStream_t stream) {
  DevBuf b;
  int st;
  if ((st = b.alloc((n_parts + 1) * sizeof(int32_t)))) return st;
  shard_bounds_kernel<<<blocks_for(n_parts + 1, 64), 64, 0, stream>>>(d_row_ptr, rows, nnz, n_parts, b.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaMemcpyAsync(h_bounds, b.p, (n_parts + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  return HISPMV_OK;
}

// ------------------------------------------------------------------------------------------------
// CSR handed in by the caller (hispmv_add_sparse_csr[_dev]): the kernels assume row_ptr non-decreasing from 0 to nnz,
// 0 <= col < cols, and columns non-decreasing inside a row (the column-slab and blocked plans cut rows by column).
// One warp per row checks all three; flags: [0] row_ptr broken, [1] column out of range, [2] a row's columns out of
// order (the caller then re-sorts through the COO path, which orders entries like the reference's std::sort).
// ------------------------------------------------------------------------------------------------
namespace {
// Flat and coalesced (the first version walked every row with one warp: 20 ms on C2, whose longest rows hold a million
// entries).  Columns are sorted inside every row iff every descent col[j] < col[j-1] sits at a row start, i.e. iff the
// descents counted over all entries equal the descents counted at the row starts.
__global__ void csr_validate_rows_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ col, int32_t rows,
                                         int64_t nnz, int* __restrict__ flags, unsigned long long* __restrict__ counts) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int at_start = 0;
  if (r < rows) {
    const int64_t b = rp[r], e = rp[r + 1];
    if (b > e || b < 0 || e > nnz) {
      flags[0] = 1;
    } else if (e > b && b > 0 && col[b] < col[b - 1]) {
      at_start = 1;
    }
  }
  at_start = __reduce_add_sync(0xffffffffu, at_start);
  if ((threadIdx.x & 31) == 0 && at_start) atomicAdd(counts, (unsigned long long)at_start);
}
__global__ void csr_validate_entries_kernel(const int32_t* __restrict__ col, int32_t cols, int64_t nnz,
                                            int* __restrict__ flags, unsigned long long* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned int descents = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) {
    const int32_t c = col[j];
    if (c < 0 || c >= cols) flags[1] = 1;
    if (j > 0 && c < col[j - 1]) ++descents;
  }
  descents = __reduce_add_sync(0xffffffffu, descents);
  if ((threadIdx.x & 31) == 0 && descents) atomicAdd(counts + 1, (unsigned long long)descents);
}
__global__ void csr_expand_rows_kernel(const int32_t* __restrict__ rp, int32_t rows, int64_t nnz,
                                       int32_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  int64_t lo = 0, hi = rows;  // the row of entry j: last r with rp[r] <= j
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rp[mid] <= j) lo = mid; else hi = mid;
  }
  out[j] = (int32_t)lo;
}
}  // namespace

int csr_validate_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, int32_t cols, int64_t nnz,
                        int* h_flags3, cudaStream_t stream) {
  h_flags3[0] = h_flags3[1] = h_flags3[2] = 0;
  if (rows <= 0) return HISPMV_OK;
  DevBuf f, cnt;
  int st;
  if ((st = f.alloc(3 * sizeof(int))) || (st = cnt.alloc(2 * sizeof(unsigned long long)))) return st;
  HISPMV_CUDA(cudaMemsetAsync(f.p, 0, 3 * sizeof(int), stream));
  HISPMV_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(unsigned long long), stream));
  // row_ptr first: the entry pass below may only look at col[] where row_ptr says entries are
  csr_validate_rows_kernel<<<blocks_for(rows, 256), 256, 0, stream>>>(d_row_ptr, d_col, rows, nnz, f.as<int>(),
                                                                      cnt.as<unsigned long long>());
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaMemcpyAsync(h_flags3, f.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  if (h_flags3[0]) return HISPMV_OK;
  if (nnz > 0) {
    const int grid = (int)std::min<int64_t>(blocks_for(nnz, 256), 148 * 32);
    csr_validate_entries_kernel<<<grid, 256, 0, stream>>>(d_col, cols, nnz, f.as<int>(), cnt.as<unsigned long long>());
    HISPMV_CUDA(cudaGetLastError());
  }
  unsigned long long h_cnt[2] = {0, 0};
  HISPMV_CUDA(cudaMemcpyAsync(h_flags3, f.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(h_cnt, cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  h_flags3[2] = h_cnt[1] != h_cnt[0] ? 1 : 0;  // a descent somewhere inside a row
  return HISPMV_OK;
}

int csr_expand_rows_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int32_t* d_rows_out,
                           cudaStream_t stream) {
  if (nnz > 0) {
    csr_expand_rows_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(d_row_ptr, rows, nnz, d_rows_out);
    HISPMV_CUDA(cudaGetLastError());
  }
  return HISPMV_OK;
}

// ------------------------------------------------------------------------------------------------
// Adaptive, row-aligned tiling (see AdaptivePlan in internal.h; restated in oracle_adaptive_tiles).
//   w_r = 0 for long rows (len >= T), 1 + len otherwise;  S = exclusive prefix sum of w
//   a short row r opens a STREAM tile when r == 0, when row r-1 is long, or when S_r / B != S_{r-1} / B
//   a long row r contributes ceil(len / CH) LONG tiles (chunk 0, 1, ...)
// Tiles are numbered in row order, so a STREAM tile's rows end where the next tile's first row begins.
// ------------------------------------------------------------------------------------------------
namespace {

struct ShortWork {
  const int32_t* rp;
  int32_t T;
  __host__ __device__ int64_t operator()(int64_t r) const {
    const int32_t len = rp[r + 1] - rp[r];
    return len >= T ? 0 : 1 + (int64_t)len;
  }
};

__global__ void adaptive_count_kernel(const int32_t* __restrict__ rp, const int64_t* __restrict__ S, int32_t rows,
                                      int32_t B, int32_t T, int32_t CH, int32_t* __restrict__ cnt,
                                      int32_t* __restrict__ split_flag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int32_t len = rp[r + 1] - rp[r];
  int32_t c;
  if (len >= T) {
    c = (len + CH - 1) / CH;
  } else if (r == 0) {
    c = 1;
  } else {
    const bool prev_long = (rp[r] - rp[r - 1]) >= T;
    c = (prev_long || (S[r] / B != S[r - 1] / B)) ? 1 : 0;
  }
  cnt[r] = c;
  split_flag[r] = (len >= T && c >= 2) ? 1 : 0;
}

__global__ void adaptive_scatter_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ cnt,
                                        const int32_t* __restrict__ first, const int32_t* __restrict__ split_flag,
                                        const int32_t* __restrict__ split_pos, int32_t rows, int32_t T,
                                        int32_t* __restrict__ tile_row, int32_t* __restrict__ tile_chunk,
                                        int32_t* __restrict__ split_rows) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int32_t c = cnt[r];
  if (c > 0) {
    const bool is_long = (rp[r + 1] - rp[r]) >= T;
    const int32_t t0 = first[r];
    for (int32_t k = 0; k < c; ++k) {
      tile_row[t0 + k] = (int32_t)r;
      tile_chunk[t0 + k] = is_long ? k : -1;
    }
  }
  if (split_flag[r]) split_rows[split_pos[r]] = (int32_t)r;
}

}  // namespace

int adaptive_tiles_device(const int32_t* d_row_ptr, int32_t rows, int32_t stream_items, int32_t long_threshold,
                          int32_t chunk_nnz, int64_t* num_tiles, int32_t** d_tile_row, int32_t** d_tile_chunk,
                          int32_t** d_split_rows, int64_t* num_split, cudaStream_t stream) {
  *num_tiles = 0;
  *num_split = 0;
  *d_tile_row = nullptr;
  *d_tile_chunk = nullptr;
  *d_split_rows = nullptr;
  if (rows <= 0) {
    HISPMV_CUDA(cudaMalloc((void**)d_tile_row, sizeof(int32_t)));
    HISPMV_CUDA(cudaMemsetAsync(*d_tile_row, 0, sizeof(int32_t), stream));
    return HISPMV_OK;
  }
  DevBuf S, cnt, first, sflag, spos, tmp;
  int st;
  if ((st = S.alloc((size_t)rows * 8)) || (st = cnt.alloc((size_t)rows * 4)) || (st = first.alloc((size_t)rows * 4)) ||
      (st = sflag.alloc((size_t)rows * 4)) || (st = spos.alloc((size_t)rows * 4)))
    return st;
  cub::CountingInputIterator<int64_t> counting(0);
  cub::TransformInputIterator<int64_t, ShortWork, cub::CountingInputIterator<int64_t>> work(
      counting, ShortWork{d_row_ptr, long_threshold});
  size_t tb = 0, tb2 = 0;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, work, S.as<int64_t>(), rows, stream));
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt.as<int32_t>(), first.as<int32_t>(), rows, stream));
  if ((st = tmp.alloc(std::max(tb, tb2)))) return st;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, work, S.as<int64_t>(), rows, stream));
  adaptive_count_kernel<<<blocks_for(rows, 256), 256, 0, stream>>>(d_row_ptr, S.as<int64_t>(), rows, stream_items,
                                                                   long_threshold, chunk_nnz, cnt.as<int32_t>(),
                                                                   sflag.as<int32_t>());
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, cnt.as<int32_t>(), first.as<int32_t>(), rows, stream));
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, sflag.as<int32_t>(), spos.as<int32_t>(), rows, stream));
  int32_t h[4];
  HISPMV_CUDA(cudaMemcpyAsync(&h[0], cnt.as<int32_t>() + (rows - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h[1], first.as<int32_t>() + (rows - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h[2], sflag.as<int32_t>() + (rows - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaMemcpyAsync(&h[3], spos.as<int32_t>() + (rows - 1), 4, cudaMemcpyDeviceToHost, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  const int64_t nt = (int64_t)h[0] + h[1];
  const int64_t ns = (int64_t)h[2] + h[3];
  HISPMV_CUDA(cudaMalloc((void**)d_tile_row, (size_t)(nt + 1) * 4));
  st = check_cuda(cudaMalloc((void**)d_tile_chunk, (size_t)std::max<int64_t>(nt, 1) * 4), "cudaMalloc", __FILE__, __LINE__);
  if (st == HISPMV_OK && ns > 0)
    st = check_cuda(cudaMalloc((void**)d_split_rows, (size_t)ns * 4), "cudaMalloc", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(*d_tile_row);
    cudaFree(*d_tile_chunk);
    *d_tile_row = nullptr;
    *d_tile_chunk = nullptr;
    return st;
  }
  adaptive_scatter_kernel<<<blocks_for(rows, 256), 256, 0, stream>>>(d_row_ptr, cnt.as<int32_t>(), first.as<int32_t>(),
                                                                     sflag.as<int32_t>(), spos.as<int32_t>(), rows,
                                                                     long_threshold, *d_tile_row, *d_tile_chunk,
                                                                     *d_split_rows);
  HISPMV_CUDA(cudaGetLastError());
  HISPMV_CUDA(cudaMemcpyAsync(*d_tile_row + nt, &rows, 4, cudaMemcpyHostToDevice, stream));
  HISPMV_CUDA(cudaStreamSynchronize(stream));
  *num_tiles = nt;
  *num_split = ns;
  return HISPMV_OK;
}

namespace {
__global__ void tile_desc_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ tile_row,
                                 const int32_t* __restrict__ tile_chunk, int64_t num_tiles, int32_t CH,
                                 TileDesc* __restrict__ desc) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  TileDesc d;
  d.r0 = tile_row[t];
  d.chunk = tile_chunk[t];
  d.tile = (int32_t)t;
  d.pad = 0;
  if (d.chunk >= 0) {
    const int32_t rb = rp[d.r0], re = rp[d.r0 + 1];
    d.r1 = d.r0 + 1;
    d.nchunks = (re - rb + CH - 1) / CH;
    d.n0 = rb + d.chunk * CH;
    d.n1 = min(re, d.n0 + CH);
  } else {
    d.r1 = tile_row[t + 1];
    d.nchunks = 0;
    d.n0 = rp[d.r0];
    d.n1 = rp[d.r1];
  }
  desc[t] = d;
}
}  // namespace

namespace {
__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace
cudaError_t fill_u32_device(uint32_t* p, uint32_t value, size_t n, cudaStream_t stream) {
  if (n) fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, value, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Column slabs (x larger than L2): slab s keeps the nonzeros with bounds[s] <= col < bounds[s+1] of every row, as a
// CSR of its own over the same rows.  Columns are sorted inside a row, so a row's share of a slab is one contiguous
// segment found by two binary searches.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void slab_count_kernel(const int32_t* __restrict__ rp, const int32_t* __restrict__ col, int32_t rows,
                                  int32_t lo_col, int32_t hi_col, int32_t* __restrict__ first,
                                  int32_t* __restrict__ len) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > rows) return;
  if (r == rows) {
    len[r] = 0;  // slot for the exclusive scan's total
    return;
  }
  const int32_t b = rp[r], e = rp[r + 1];
  int32_t lo = b, hi = e;  // first entry with col >= lo_col
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (col[mid] < lo_col) lo = mid + 1; else hi = mid;
  }
  const int32_t f = lo;
  hi = e;  // first entry with col >= hi_col
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (col[mid] < hi_col) lo = mid + 1; else hi = mid;
  }
  first[r] = f;
  len[r] = lo - f;
}
__global__ void slab_copy_kernel(const int32_t* __restrict__ col, const float* __restrict__ val,
                                 const int32_t* __restrict__ first, const int32_t* __restrict__ srp, int32_t rows,
                                 int64_t nnz, int64_t padded, int32_t* __restrict__ ocol, float* __restrict__ oval) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= padded) return;
  if (j >= nnz) {
    ocol[j] = 0;
    oval[j] = 0.0f;
    return;
  }
  int32_t lo = 0, hi = rows;  // row of output entry j: last r with srp[r] <= j
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if ((int64_t)srp[mid] <= j) lo = mid; else hi = mid;
  }
  const int64_t src = (int64_t)first[lo] + (j - srp[lo]);
  ocol[j] = col[src];
  oval[j] = val[src];
}
}  // namespace

int csr_column_slab_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t rows,
                           int32_t lo_col, int32_t hi_col, int32_t** o_row_ptr, int32_t** o_col, float** o_val,
                           int64_t* o_nnz, cudaStream_t stream) {
  *o_row_ptr = nullptr;
  *o_col = nullptr;
  *o_val = nullptr;
  *o_nnz = 0;
  DevBuf first, len, tmp;
  int st;
  if ((st = first.alloc(((size_t)rows + 1) * 4)) || (st = len.alloc(((size_t)rows + 1) * 4))) return st;
  HISPMV_CUDA(cudaMalloc((void**)o_row_ptr, ((size_t)rows + 1 + 4) * sizeof(int32_t)));
  slab_count_kernel<<<blocks_for((int64_t)rows + 1, 256), 256, 0, stream>>>(d_row_ptr, d_col, rows, lo_col, hi_col,
                                                                           first.as<int32_t>(), len.as<int32_t>());
  size_t tb = 0;
  HISPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.as<int32_t>(), *o_row_ptr, (int64_t)rows + 1, stream));
  st = tmp.alloc(tb);
  if (st == HISPMV_OK)
    st = check_cuda(cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.as<int32_t>(), *o_row_ptr, (int64_t)rows + 1, stream),
                    "scan", __FILE__, __LINE__);
  int32_t total = 0;
  if (st == HISPMV_OK)
    st = check_cuda(cudaMemcpyAsync(&total, *o_row_ptr + rows, 4, cudaMemcpyDeviceToHost, stream), "D2H", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamSynchronize(stream), "sync", __FILE__, __LINE__);
  const int64_t pad = padded_nnz(total);
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)o_col, pad * 4), "cudaMalloc(slab col)", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)o_val, pad * 4), "cudaMalloc(slab val)", __FILE__, __LINE__);
  if (st == HISPMV_OK) {
    slab_copy_kernel<<<blocks_for(pad, 256), 256, 0, stream>>>(d_col, d_val, first.as<int32_t>(), *o_row_ptr, rows,
                                                               total, pad, *o_col, *o_val);
    st = check_cuda(cudaGetLastError(), "slab_copy", __FILE__, __LINE__);
  }
  if (st == HISPMV_OK) st = check_cuda(cudaStreamSynchronize(stream), "sync", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(*o_row_ptr);
    cudaFree(*o_col);
    cudaFree(*o_val);
    *o_row_ptr = nullptr;
    *o_col = nullptr;
    *o_val = nullptr;
    return st;
  }
  *o_nnz = total;
  return HISPMV_OK;
}

// Slab width: x beyond kSlabMinBytes does not stay in L2 under random gathers (measured: 275 G gathers/s for a 40 MB
// table, 101 G/s at 160 MB, 60 G/s at 400 MB), so the columns are cut into equal slabs of at most kSlabMaxCols.
// C5 on one GPU (100 M columns): 17.5 ms whole, 7.5 ms as 8 slabs -- each pass pays row_ptr + y read + y write again.
int32_t select_slab_cols(int32_t cols, int64_t nnz, const ColProbe& probe) {
  const bool banded = probe.cmp >= 64 && probe.near * 4 >= probe.cmp * 3;
  if (banded || (int64_t)cols * 4 <= kSlabMinBytes || nnz < 4000000) return 0;
  const int64_t n = ((int64_t)cols + kSlabMaxCols - 1) / kSlabMaxCols;
  const int64_t w = ((int64_t)cols + n - 1) / n;
  return (int32_t)((w + 31) & ~(int64_t)31);
}

int tile_desc_device(const int32_t* d_row_ptr, const int32_t* d_tile_row, const int32_t* d_tile_chunk,
                     int64_t num_tiles, int32_t chunk_nnz, TileDesc** d_desc, cudaStream_t stream) {
  *d_desc = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)d_desc, (size_t)std::max<int64_t>(num_tiles, 1) * sizeof(TileDesc)));
  if (num_tiles > 0) {
    tile_desc_kernel<<<blocks_for(num_tiles, 256), 256, 0, stream>>>(d_row_ptr, d_tile_row, d_tile_chunk, num_tiles,
                                                                     chunk_nnz, *d_desc);
    HISPMV_CUDA(cudaGetLastError());
  }
  return HISPMV_OK;
}

// ------------------------------------------------------------------------------------------------
// Runtime selector.  Pure integer arithmetic on the row-length histogram and the column-locality probe, so that the
// CPU restatement (oracle/oracle.c: oracle_select_kernel) reproduces the decision bit for bit.
//
// The FPGA design space (channels A:B:C, pre-accumulator, row-distribution network; dse.py:23-95) collapses on a GPU
// to two questions: may rows be split across CTAs at all (row_dist_net), and in which order should lanes walk the
// nonzeros so that the x gathers of one instruction share cache lines (the measured ceiling on gather-heavy matrices
// is the ~0.95 L1-miss requests per clock an SM can issue)?
//   mean    = ceil(nnz / rows);  lanes = largest power of two <= mean, clamped to [2, 32]
//   regular = rows of >= 4*mean nonzeros hold < 1/8 of all nonzeros (estimated from the power-of-two histogram:
//             a row in bin k, 2^(k-1) <= len < 2^k, counts 3 * 2^(k-2) nonzeros, 1 for k = 1)
//   banded  = the probe compared >= 64 entry pairs and >= 3/4 of them were within 32 columns of the row above
//   EMPTY       no nonzeros
//   CSR_SCALAR / CSR_VECTOR(lanes)   row splitting not allowed (row_dist_net off): mean <= 2 / otherwise
//   ROWSTAGE    regular, banded and mean <= 64: row-major walk behind a TMA-staged stream (longer rows already
//               give a nnz-major warp 32 column-sorted neighbours of one row, and would all be LONG tiles here)
//   ADAPTIVE    everything else: nnz-major tiles, heavy rows chunked across CTAs
// ------------------------------------------------------------------------------------------------
void select_kernel(const RowStats& st, const ColProbe& probe, int allow_split_rows, int* kernel, int* lanes) {
  *lanes = 0;
  if (st.rows <= 0 || st.nnz <= 0) {
    *kernel = HISPMV_KERNEL_EMPTY;
    return;
  }
  const int64_t mean = (st.nnz + st.rows - 1) / st.rows;
  int l = 2;
  while (l < 32 && (int64_t)l * 2 <= mean) l *= 2;
  if (!allow_split_rows) {
    if (mean <= 2) {
      *kernel = HISPMV_KERNEL_CSR_SCALAR;
    } else {
      *kernel = HISPMV_KERNEL_CSR_VECTOR;
      *lanes = l;
    }
    return;
  }
  int64_t heavy_nnz = 0;
  for (int k = 1; k < HISPMV_HIST_BINS; ++k) {
    const int64_t lo = (int64_t)1 << (k - 1);
    if (lo >= 4 * mean) heavy_nnz += st.hist[k] * (k == 1 ? 1 : 3 * ((int64_t)1 << (k - 2)));
  }
  const bool regular = heavy_nnz * 8 < st.nnz;
  const bool banded = probe.cmp >= 64 && probe.near * 4 >= probe.cmp * 3;
  *kernel = (regular && banded && mean <= 64) ? HISPMV_KERNEL_ROWSTAGE : HISPMV_KERNEL_ADAPTIVE;
}

// ROWSTAGE: CTAs of kRowstageThreads = 128 threads, `lanes` lanes per row, R = 128 / lanes rows per pass; the STREAM
// budget is R rows of average length (row ends count as items), so a regular matrix gets tiles of about R rows and
// every lane group has a row.  Small tiles on purpose: 16 KB of staged stream per CTA lets ~13 CTAs share an SM, i.e.
// 13 independent load -> gather -> reduce chains in flight (C4: 0.84 of HBM peak with 256 threads and 3584-item
// tiles, 0.92 with 128 threads and 1792).
//   items = ceil((nnz + rows) / rows);  lanes = smallest power of two with (128 / lanes) * items <= 1792
//   B = max(128, (128 / lanes) * items);  T = 256 (longer rows become LONG tiles);  CH = 4096
void rowstage_params(const RowStats& st, int lanes_in, int* lanes, int32_t* stream_items, int32_t* long_threshold,
                     int32_t* chunk_nnz) {
  const int64_t rows = std::max<int64_t>(st.rows, 1);
  const int64_t items = (st.nnz + rows + rows - 1) / rows;
  int l = lanes_in;
  if (l <= 0) {
    l = 1;
    while (l < 32 && (128 / l) * items > 1792) l *= 2;
  }
  int64_t b = (128 / l) * items;
  b = std::max<int64_t>(128, std::min<int64_t>(b, kRowstageMaxCap - 256));
  *lanes = l;
  *stream_items = (int32_t)b;
  *long_threshold = 256;
  *chunk_nnz = 4096;
}

int merge_tile_items_for(const RowStats& st) {
  // small problems: smaller tiles so that there are enough CTAs to cover 148 SMs
  const int64_t total = (int64_t)st.rows + st.nnz;
  if (total < 148LL * 4 * 1792) return 128 * 7;
  return 256 * 7;
}

}  // namespace hispmv
