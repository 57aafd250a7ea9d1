// Deterministic synthetic CSR generators running in HBM (see include/hispmv_synth.h).
// Benchmark / test support: these build BASELINE.json's C2 / C4 / C5 matrices without touching the host.
// Every entry is a pure function of (seed, row, k); oracle/oracle.c restates the same functions on the CPU.
#include <cub/cub.cuh>

#include <vector>

#include "hispmv_synth.h"
#include "internal.h"

namespace hispmv {
namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash_row(uint64_t seed, int64_t r) {
  return mix64(mix64(seed) ^ (uint64_t)r);
}
__host__ __device__ __forceinline__ uint64_t hash_entry(uint64_t seed, int64_t r, int64_t k) {
  return mix64(hash_row(seed ^ 0xA5A5A5A5A5A5A5A5ull, r) + (uint64_t)k * 0xD1342543DE82EF95ull);
}

struct SynthParams {
  int kind;
  uint64_t seed;
  int32_t rows, cols;
  int64_t p0, p1, p2;
};

__device__ __forceinline__ int32_t synth_row_len(const SynthParams& sp, int64_t r) {
  switch (sp.kind) {
    case HISPMV_SYNTH_POWERLAW: {
      const int64_t period = sp.p2 >> 8;   // > 0: the row-length sequence repeats (weak-scaling stacks of one shape)
      const uint64_t u = hash_row(sp.seed, period > 0 ? r % period : r) >> 32;
      const uint64_t len = (uint64_t)sp.p0 / (u + 1);
      return (int32_t)(len < (uint64_t)sp.p1 ? len : (uint64_t)sp.p1);
    }
    case HISPMV_SYNTH_UNIFORM: {
      const uint64_t h = hash_row(sp.seed, r);
      return (int32_t)(sp.p0 + __popcll(h & (uint64_t)sp.p1));
    }
    case HISPMV_SYNTH_STENCIL27: {
      const int64_t nx = sp.p0, ny = sp.p1, nz = sp.p2;
      const int64_t ix = r % nx, iy = (r / nx) % ny, iz = r / (nx * ny);
      const int cx = 1 + (ix > 0) + (ix < nx - 1);
      const int cy = 1 + (iy > 0) + (iy < ny - 1);
      const int cz = 1 + (iz > 0) + (iz < nz - 1);
      return cx * cy * cz;
    }
  }
  return 0;
}

__device__ __forceinline__ void synth_entry(const SynthParams& sp, int64_t r, int32_t k, int32_t len, int32_t* col,
                                            float* val) {
  const uint64_t h = hash_entry(sp.seed, r, k);
  // value: uniform in [-1, 1) on a 2^-23 grid, exact in fp32
  *val = (float)((int32_t)(h & 0xFFFFFF) - 0x800000) * (1.0f / 8388608.0f);
  if (sp.kind == HISPMV_SYNTH_STENCIL27) {
    const int64_t nx = sp.p0, ny = sp.p1, nz = sp.p2;
    const int64_t ix = r % nx, iy = (r / nx) % ny, iz = r / (nx * ny);
    const int cx = 1 + (ix > 0) + (ix < nx - 1);
    const int cy = 1 + (iy > 0) + (iy < ny - 1);
    const int a = k / (cy * cx), b = (k / cx) % cy, c = k % cx;
    const int64_t dz = a - (iz > 0), dy = b - (iy > 0), dx = c - (ix > 0);
    *col = (int32_t)(r + dz * nx * ny + dy * nx + dx);
    return;
  }
  // stratified draw: q in [k/len, (k+1)/len), then a monotone map to a column => sorted within the row
  const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  const double q = __ddiv_rn(__dadd_rn((double)k, u), (double)len);
  double w = q;
  const int gamma = sp.kind == HISPMV_SYNTH_POWERLAW ? (int)(sp.p2 & 0xFF) : 1;
  for (int g = 1; g < gamma; ++g) w = __dmul_rn(w, q);
  int64_t c = (int64_t)__dmul_rn(w, (double)sp.cols);
  if (c >= sp.cols) c = sp.cols - 1;
  *col = (int32_t)c;
}

__global__ void row_len_kernel(SynthParams sp, int32_t row_begin, int32_t n, int64_t* __restrict__ len64,
                               int32_t* __restrict__ len32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int32_t l = synth_row_len(sp, (int64_t)row_begin + i);
    if (len64) len64[i] = l;
    if (len32) len32[i] = l;
  }
  if (i == n && len32) len32[n] = 0;  // slot for the exclusive scan's total
  if (i == n && len64) len64[n] = 0;
}

__global__ void fill_entries_kernel(SynthParams sp, int32_t row_begin, int32_t n_rows,
                                    const int32_t* __restrict__ row_ptr, int64_t nnz, int64_t padded,
                                    int32_t* __restrict__ col, float* __restrict__ val) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= padded) return;
  if (j >= nnz) {
    col[j] = 0;
    val[j] = 0.0f;
    return;
  }
  // row of entry j: last r with row_ptr[r] <= j
  int32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if ((int64_t)row_ptr[mid] <= j) lo = mid; else hi = mid;
  }
  const int32_t b = row_ptr[lo], e = row_ptr[lo + 1];
  int32_t c;
  float v;
  synth_entry(sp, (int64_t)row_begin + lo, (int32_t)(j - b), e - b, &c, &v);
  col[j] = c;
  val[j] = v;
}

__global__ void bounds_from_prefix_kernel(const int64_t* __restrict__ prefix, int32_t rows, int n_parts,
                                          int32_t* __restrict__ bounds) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_parts) return;
  if (k == 0) { bounds[0] = 0; return; }
  if (k == n_parts) { bounds[k] = rows; return; }
  const int64_t nnz = prefix[rows];
  const int64_t target = (nnz * (int64_t)k) / n_parts;
  int64_t lo = 0, hi = rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (prefix[mid] < target) lo = mid + 1; else hi = mid;
  }
  bounds[k] = (int32_t)lo;
}

int make_params(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, SynthParams* sp) {
  if (!params || rows < 0 || cols <= 0) {
    set_error("synth: bad arguments");
    return HISPMV_ERR_ARG;
  }
  sp->kind = kind;
  sp->seed = seed;
  sp->rows = rows;
  sp->cols = cols;
  sp->p0 = params[0];
  sp->p1 = params[1];
  sp->p2 = kind == HISPMV_SYNTH_UNIFORM ? 0 : params[2];
  if (kind == HISPMV_SYNTH_STENCIL27) {
    if (sp->p0 < 1 || sp->p1 < 1 || sp->p2 < 1 || sp->p0 * sp->p1 * sp->p2 != rows || rows != cols) {
      set_error("synth: stencil needs rows == cols == nx*ny*nz");
      return HISPMV_ERR_ARG;
    }
  } else if (kind == HISPMV_SYNTH_POWERLAW) {
    if (sp->p0 < 1 || sp->p1 < 0 || (sp->p2 & 0xFF) < 1 || (sp->p2 & 0xFF) > 16 || sp->p2 < 0) {
      set_error("synth: powerlaw needs K >= 1, clip >= 0, 1 <= gamma <= 16 (params[2] = gamma | period << 8)");
      return HISPMV_ERR_ARG;
    }
  } else if (kind != HISPMV_SYNTH_UNIFORM) {
    set_error("synth: unknown kind");
    return HISPMV_ERR_ARG;
  }
  return HISPMV_OK;
}

// exclusive prefix (64-bit) of the row lengths of [row_begin,row_end); d_prefix has n+1 entries
int prefix64(const SynthParams& sp, int32_t row_begin, int32_t n, int64_t** d_prefix) {
  *d_prefix = nullptr;
  int64_t* d_len = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)&d_len, ((size_t)n + 1) * 8));
  int st = check_cuda(cudaMalloc((void**)d_prefix, ((size_t)n + 1) * 8), "cudaMalloc", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    cudaFree(d_len);
    return st;
  }
  row_len_kernel<<<(int)(((int64_t)n + 1 + 255) / 256), 256>>>(sp, row_begin, n, d_len, nullptr);
  size_t tb = 0;
  void* tmp = nullptr;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, *d_prefix, (int64_t)n + 1);
  st = check_cuda(cudaMalloc(&tmp, tb ? tb : 16), "cudaMalloc", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cub::DeviceScan::ExclusiveSum(tmp, tb, d_len, *d_prefix, (int64_t)n + 1), "scan", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaDeviceSynchronize(), "sync", __FILE__, __LINE__);
  cudaFree(tmp);
  cudaFree(d_len);
  if (st != HISPMV_OK) {
    cudaFree(*d_prefix);
    *d_prefix = nullptr;
  }
  return st;
}

}  // namespace
}  // namespace hispmv

using namespace hispmv;

extern "C" {

int hispmv_synth_count(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int32_t row_begin,
                       int32_t row_end, int64_t* nnz) {
  SynthParams sp;
  int st = make_params(kind, seed, rows, cols, params, &sp);
  if (st != HISPMV_OK) return st;
  if (!nnz || row_begin < 0 || row_end < row_begin || row_end > rows) {
    set_error("synth_count: bad row range");
    return HISPMV_ERR_ARG;
  }
  int64_t* d_prefix = nullptr;
  const int32_t n = row_end - row_begin;
  st = prefix64(sp, row_begin, n, &d_prefix);
  if (st != HISPMV_OK) return st;
  st = check_cuda(cudaMemcpy(nnz, d_prefix + n, 8, cudaMemcpyDeviceToHost), "D2H", __FILE__, __LINE__);
  cudaFree(d_prefix);
  return st;
}

int hispmv_synth_shard_bounds(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int n_parts,
                              int32_t* bounds, int64_t* total_nnz) {
  SynthParams sp;
  int st = make_params(kind, seed, rows, cols, params, &sp);
  if (st != HISPMV_OK) return st;
  if (!bounds || n_parts < 1) return HISPMV_ERR_ARG;
  int64_t* d_prefix = nullptr;
  st = prefix64(sp, 0, rows, &d_prefix);
  if (st != HISPMV_OK) return st;
  int32_t* d_b = nullptr;
  st = check_cuda(cudaMalloc((void**)&d_b, (n_parts + 1) * 4), "cudaMalloc", __FILE__, __LINE__);
  if (st == HISPMV_OK) {
    bounds_from_prefix_kernel<<<(n_parts + 64) / 64, 64>>>(d_prefix, rows, n_parts, d_b);
    st = check_cuda(cudaMemcpy(bounds, d_b, (n_parts + 1) * 4, cudaMemcpyDeviceToHost), "D2H", __FILE__, __LINE__);
    if (st == HISPMV_OK && total_nnz)
      st = check_cuda(cudaMemcpy(total_nnz, d_prefix + rows, 8, cudaMemcpyDeviceToHost), "D2H", __FILE__, __LINE__);
  }
  cudaFree(d_b);
  cudaFree(d_prefix);
  return st;
}

int hispmv_synth_csr(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int32_t row_begin,
                     int32_t row_end, int32_t** d_row_ptr, int32_t** d_col, float** d_val, int64_t* nnz_out) {
  SynthParams sp;
  int st = make_params(kind, seed, rows, cols, params, &sp);
  if (st != HISPMV_OK) return st;
  if (!d_row_ptr || !d_col || !d_val || !nnz_out || row_begin < 0 || row_end < row_begin || row_end > rows) {
    set_error("synth_csr: bad arguments");
    return HISPMV_ERR_ARG;
  }
  *d_row_ptr = nullptr;
  *d_col = nullptr;
  *d_val = nullptr;
  const int32_t n = row_end - row_begin;
  int64_t total = 0;
  st = hispmv_synth_count(kind, seed, rows, cols, params, row_begin, row_end, &total);
  if (st != HISPMV_OK) return st;
  if (total >= (int64_t)INT32_MAX) {
    set_error("synth_csr: row block holds 2^31 or more nonzeros; shard it");
    return HISPMV_ERR_ARG;
  }
  int32_t* d_len = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)&d_len, ((size_t)n + 1) * 4));
  st = check_cuda(cudaMalloc((void**)d_row_ptr, ((size_t)n + 1) * 4), "cudaMalloc", __FILE__, __LINE__);
  const int64_t padded = ((total + 3) & ~(int64_t)3) + 4;
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)d_col, padded * 4), "cudaMalloc(col)", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)d_val, padded * 4), "cudaMalloc(val)", __FILE__, __LINE__);
  void* tmp = nullptr;
  if (st == HISPMV_OK) {
    row_len_kernel<<<(int)(((int64_t)n + 1 + 255) / 256), 256>>>(sp, row_begin, n, nullptr, d_len);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, d_len, *d_row_ptr, (int64_t)n + 1);
    st = check_cuda(cudaMalloc(&tmp, tb ? tb : 16), "cudaMalloc", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cub::DeviceScan::ExclusiveSum(tmp, tb, d_len, *d_row_ptr, (int64_t)n + 1), "scan", __FILE__, __LINE__);
  }
  if (st == HISPMV_OK) {
    fill_entries_kernel<<<(int)((padded + 255) / 256), 256>>>(sp, row_begin, n, *d_row_ptr, total, padded, *d_col, *d_val);
    st = check_cuda(cudaGetLastError(), "fill_entries", __FILE__, __LINE__);
  }
  if (st == HISPMV_OK) st = check_cuda(cudaDeviceSynchronize(), "sync", __FILE__, __LINE__);
  cudaFree(tmp);
  cudaFree(d_len);
  if (st != HISPMV_OK) {
    cudaFree(*d_row_ptr);
    cudaFree(*d_col);
    cudaFree(*d_val);
    *d_row_ptr = nullptr;
    *d_col = nullptr;
    *d_val = nullptr;
    return st;
  }
  *nnz_out = total;
  return HISPMV_OK;
}

void hispmv_synth_free(void* p) { cudaFree(p); }

}  // extern "C"
