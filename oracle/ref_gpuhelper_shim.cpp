// TEST INFRASTRUCTURE (oracle/) -- C-ABI doorway into the UNMODIFIED reference helper
// /root/reference/gpu/src/spmvHelper.cpp (compiled where it lies into oracle/_ref/libref_gpuhelper.so):
//   cooToCsr   gpu/src/spmvHelper.cpp:117-156   (the flat COO -> CSR the bit-exact contract is pinned to)
//   cpuSpMV    gpu/src/spmvHelper.cpp:223-233
//   loadMtx    gpu/src/spmvHelper.cpp:4-115
#include <cstdint>
#include <cstring>

#include "spmvHelper.h"  // reference header (gpu/include)

extern "C" {

int ref_gpu_coo_to_csr(int rows, int cols, int64_t nnz, const int* r, const int* c, const float* v, int* row_ptr,
                       int* col, float* val) {
  COOMatrix_t coo;
  coo.rows.assign(r, r + nnz);
  coo.cols.assign(c, c + nnz);
  coo.values.assign(v, v + nnz);
  coo.rows_count = rows;
  coo.cols_count = cols;
  coo.nnz = (int)nnz;
  CSRMatrix_t csr = cooToCsr(coo);
  std::memcpy(row_ptr, csr.row_offsets.data(), sizeof(int) * (rows + 1));
  std::memcpy(col, csr.col_indices.data(), sizeof(int) * nnz);
  std::memcpy(val, csr.values.data(), sizeof(float) * nnz);
  return 0;
}

void ref_gpu_cpu_spmv(int rows, int64_t nnz, const int* r, const int* c, const float* v, int cols, const float* x,
                      float* y, float alpha, float beta) {
  std::vector<int> rv(r, r + nnz), cv(c, c + nnz);
  std::vector<float> vv(v, v + nnz), xv(x, x + cols), yv(y, y + rows);
  cpuSpMV(rows, (int)nnz, rv, cv, vv, xv, yv, alpha, beta);
  std::memcpy(y, yv.data(), sizeof(float) * rows);
}

static COOMatrix_t g_coo;
int ref_gpu_load_mtx(const char* path, int* rows, int* cols, int64_t* nnz) {
  try {
    g_coo = loadMtx(path);
  } catch (const std::exception&) {
    return -1;
  }
  *rows = g_coo.rows_count;
  *cols = g_coo.cols_count;
  *nnz = (int64_t)g_coo.rows.size();
  return 0;
}
void ref_gpu_load_mtx_fetch(int* r, int* c, float* v) {
  std::memcpy(r, g_coo.rows.data(), sizeof(int) * g_coo.rows.size());
  std::memcpy(c, g_coo.cols.data(), sizeof(int) * g_coo.cols.size());
  std::memcpy(v, g_coo.values.data(), sizeof(float) * g_coo.values.size());
  g_coo = COOMatrix_t();
}
}
