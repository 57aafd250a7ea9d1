/* TEST INFRASTRUCTURE (oracle/) -- not product code.
 *
 * Hand-declared subset of Intel MKL's <mkl.h>, resolved at link time against the MKL that
 * PyTorch's libtorch_cpu.so bundles (oneAPI MKL 2024.2, LP64).  Standalone MKL (libmkl_rt)
 * is not installed in this image; this header lets the reference's unmodified
 * /root/reference/cpu/src/main.cpp (which does `#include <mkl.h>`, main.cpp:2) compile.
 * Only what main.cpp:26-49,74-96,136-137 calls is declared.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
struct sparse_matrix;
typedef struct sparse_matrix* sparse_matrix_t;
typedef enum { SPARSE_STATUS_SUCCESS = 0 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50 } sparse_diag_type_t;
struct matrix_descr {
  sparse_matrix_type_t type;
  sparse_fill_mode_t mode;
  sparse_diag_type_t diag;
};
sparse_status_t mkl_sparse_s_create_csr(sparse_matrix_t*, sparse_index_base_t, int, int, int*, int*, int*, float*);
sparse_status_t mkl_sparse_s_mv(sparse_operation_t, float, const sparse_matrix_t, struct matrix_descr, const float*,
                                float, float*);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t);
int MKL_Get_Max_Threads(void);
int MKL_Set_Num_Threads_Local(int);
void sgemv_(const char*, const int*, const int*, const float*, const float*, const int*, const float*, const int*,
            const float*, float*, const int*);
#ifdef __cplusplus
}
#endif

#define mkl_get_max_threads MKL_Get_Max_Threads
static inline void mkl_set_num_threads(int n) { MKL_Set_Num_Threads_Local(n); }

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112 } CBLAS_TRANSPOSE;
/* libtorch_cpu exports the Fortran sgemv_ but not cblas_sgemv: row-major NoTrans == col-major Trans. */
static inline void cblas_sgemv(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE trans, int m, int n, float alpha, const float* A,
                               int lda, const float* x, int incx, float beta, float* y, int incy) {
  (void)layout;
  (void)trans;
  sgemv_("T", &n, &m, &alpha, A, &lda, x, &incx, &beta, y, &incy);
}
