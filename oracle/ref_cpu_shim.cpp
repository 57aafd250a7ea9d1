// TEST INFRASTRUCTURE (oracle/) -- not product code; only tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library this file is linked into.
//
// C-ABI doorway into the UNMODIFIED reference CPU path.  The Makefile compiles
// /root/reference/cpu/src/main.cpp (with -Dmain=hispmv_ref_cpu_main) and
// /root/reference/cpu/src/helper_functions.cpp where they lie and links them with this file into
// oracle/_ref/libref_cpu.so.  Nothing of the reference is copied here: we only declare the
// signatures of its non-static functions and forward plain pointers to them.
//
//   cpu_spmv            /root/reference/cpu/src/main.cpp:11-23
//   mkl_spmv            /root/reference/cpu/src/main.cpp:26-49
//   naive_gemv          /root/reference/cpu/src/main.cpp:53-71
//   mkl_gemv            /root/reference/cpu/src/main.cpp:74-96
//   readMatrixCSC       /root/reference/cpu/src/helper_functions.cpp:148-210
//   convertCSCtoCSR     /root/reference/cpu/src/helper_functions.cpp:212-241
#include <cstdint>
#include <cstring>
#include <vector>

#include "helper_functions.h"  // reference header: struct CSRMatrix, readMatrixCSC, convertCSCtoCSR
#include <mkl.h>               // oracle/mkl_shim/mkl.h

void cpu_spmv(const CSRMatrix& A, const std::vector<float> B, std::vector<float>& C, const float alpha,
              const float beta, const int rp_time);
double mkl_spmv(const CSRMatrix& matrix, const int rows, const int cols, const int nnz, const std::vector<float>& x,
                std::vector<float>& y, float alpha, float beta, int rp_time);
double naive_gemv(int rows, int cols, const std::vector<float>& A, const std::vector<float>& x, std::vector<float>& y,
                  float alpha, float beta, int rp_time);
double mkl_gemv(int rows, int cols, const std::vector<float>& A, const std::vector<float>& x, std::vector<float>& y,
                float alpha, float beta, int rp_time);

namespace {
CSRMatrix make_csr(const int* row_ptr, const int* col, const float* val, int rows, int64_t nnz) {
  CSRMatrix m;
  m.row_offsets.assign(row_ptr, row_ptr + rows + 1);
  m.col_indices.assign(col, col + nnz);
  m.values.assign(val, val + nnz);
  return m;
}
}  // namespace

extern "C" {

void ref_set_threads(int n) { mkl_set_num_threads(n); }
int ref_get_threads(void) { return mkl_get_max_threads(); }

// y (in/out) <- alpha*A*x + beta*y, single-thread naive loop of the reference.
void ref_cpu_spmv(const int* row_ptr, const int* col, const float* val, int rows, int cols, int64_t nnz, const float* x,
                  float* y, float alpha, float beta) {
  CSRMatrix m = make_csr(row_ptr, col, val, rows, nnz);
  std::vector<float> xv(x, x + cols), yv(y, y + rows);
  cpu_spmv(m, xv, yv, alpha, beta, 1);
  std::memcpy(y, yv.data(), sizeof(float) * rows);
}

// y (in/out) <- alpha*A*x + beta*y via mkl_sparse_s_mv, exactly as the reference calls it.
// Returns the reference's own timing (ns per call) of the rp_time loop.  rp_time must be 1 when
// y is used as a result (the loop is in place).
double ref_mkl_spmv(const int* row_ptr, const int* col, const float* val, int rows, int cols, int64_t nnz,
                    const float* x, float* y, float alpha, float beta, int rp_time) {
  CSRMatrix m = make_csr(row_ptr, col, val, rows, nnz);
  std::vector<float> xv(x, x + cols), yv(y, y + rows);
  double ns = mkl_spmv(m, rows, cols, (int)nnz, xv, yv, alpha, beta, rp_time);
  std::memcpy(y, yv.data(), sizeof(float) * rows);
  return ns;
}

double ref_naive_gemv(const float* A, int rows, int cols, const float* x, float* y, float alpha, float beta) {
  std::vector<float> Av(A, A + (size_t)rows * cols), xv(x, x + cols), yv(y, y + rows);
  double ns = naive_gemv(rows, cols, Av, xv, yv, alpha, beta, 1);
  std::memcpy(y, yv.data(), sizeof(float) * rows);
  return ns;
}

double ref_mkl_gemv(const float* A, int rows, int cols, const float* x, float* y, float alpha, float beta,
                    int rp_time) {
  std::vector<float> Av(A, A + (size_t)rows * cols), xv(x, x + cols), yv(y, y + rows);
  double ns = mkl_gemv(rows, cols, Av, xv, yv, alpha, beta, rp_time);
  std::memcpy(y, yv.data(), sizeof(float) * rows);
  return ns;
}

// Matrix Market -> CSR through the reference reader (readMatrixCSC + convertCSCtoCSR).
// Two-call protocol: first with null outputs to get sizes, then with buffers.
static CSRMatrix g_last;
static int g_rows, g_cols, g_nnz;
int ref_read_mtx(const char* path, int* rows, int* cols, int64_t* nnz) {
  std::vector<float> v;
  std::vector<int> ri, co;
  readMatrixCSC(const_cast<char*>(path), v, ri, co, g_rows, g_cols, g_nnz);
  g_last = CSRMatrix();
  convertCSCtoCSR(v, ri, co, g_last.values, g_last.col_indices, g_last.row_offsets, g_rows, g_cols, g_nnz);
  *rows = g_rows;
  *cols = g_cols;
  *nnz = g_nnz;
  return 0;
}
void ref_read_mtx_fetch(int* row_ptr, int* col, float* val) {
  std::memcpy(row_ptr, g_last.row_offsets.data(), sizeof(int) * (g_rows + 1));
  std::memcpy(col, g_last.col_indices.data(), sizeof(int) * g_nnz);
  std::memcpy(val, g_last.values.data(), sizeof(float) * g_nnz);
  g_last = CSRMatrix();
}

}  // extern "C"
