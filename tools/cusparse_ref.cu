// COMPARATOR, not product code.  The reference's GPU baseline is cuSPARSE's generic SpMV
// (/root/reference/gpu/src/spmv.cu:83-103: cusparseCreateCsr with 32-bit indices, CUDA_R_32F,
// CUSPARSE_SPMV_ALG_DEFAULT, one external buffer).  This file makes exactly that call sequence on a CSR that already
// lives in HBM and times it with CUDA events, so bench.py and tools/sweep.py can print the vendor library's time on the
// same matrix, on the same GPU, in the same run (`vs_cusparse`).  Nothing under hispmv_b200/ links or loads it.
//
//   nvcc -O2 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a tools/cusparse_ref.cu \
//        -o tools/libcusparse_ref.so -lcusparse
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdint.h>
#include <stdio.h>

#define CK(call)                                                                 \
  do {                                                                           \
    auto _s = (call);                                                            \
    if ((int)_s != 0) {                                                          \
      snprintf(g_err, sizeof(g_err), "%s failed with %d (%s:%d)", #call, (int)_s, __FILE__, __LINE__); \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

static char g_err[256];

extern "C" {

const char* cusparse_ref_error(void) { return g_err; }
int cusparse_ref_version(void) {
  int v = 0;
  cusparseHandle_t h;
  if (cusparseCreate(&h) != CUSPARSE_STATUS_SUCCESS) return -1;
  cusparseGetVersion(h, &v);
  cusparseDestroy(h);
  return v;
}

// y = alpha * A x + beta * y, `warmup` untimed calls then `steps` calls between two CUDA events on `stream`.
// Device pointers throughout.  *ms_per_call gets the mean; d_y holds the result of the last call (with beta != 0 it
// accumulates from call to call exactly as the reference's loop does -- pass beta = 0 to read a meaningful y).
int cusparse_ref_spmv(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int64_t rows, int64_t cols,
                      int64_t nnz, const float* d_x, float* d_y, float alpha, float beta, int warmup, int steps,
                      void* stream, float* ms_per_call, int64_t* buffer_bytes) {
  cusparseHandle_t handle = nullptr;
  cusparseSpMatDescr_t matA;
  cusparseDnVecDescr_t vecX, vecY;
  void* dBuffer = nullptr;
  size_t bufferSize = 0;
  cudaStream_t s = (cudaStream_t)stream;
  CK(cusparseCreate(&handle));
  CK(cusparseSetStream(handle, s));
  CK(cusparseCreateCsr(&matA, rows, cols, nnz, (void*)d_row_ptr, (void*)d_col, (void*)d_val, CUSPARSE_INDEX_32I,
                       CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
  CK(cusparseCreateDnVec(&vecX, cols, (void*)d_x, CUDA_R_32F));
  CK(cusparseCreateDnVec(&vecY, rows, (void*)d_y, CUDA_R_32F));
  CK(cusparseSpMV_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_32F,
                             CUSPARSE_SPMV_ALG_DEFAULT, &bufferSize));
  CK(cudaMalloc(&dBuffer, bufferSize ? bufferSize : 16));
  if (buffer_bytes) *buffer_bytes = (int64_t)bufferSize;
  for (int i = 0; i < warmup; ++i)
    CK(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_32F,
                    CUSPARSE_SPMV_ALG_DEFAULT, dBuffer));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < steps; ++i)
    CK(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, vecX, &beta, vecY, CUDA_R_32F,
                    CUSPARSE_SPMV_ALG_DEFAULT, dBuffer));
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_per_call) *ms_per_call = steps > 0 ? ms / steps : 0.0f;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CK(cusparseDestroySpMat(matA));
  CK(cusparseDestroyDnVec(vecX));
  CK(cusparseDestroyDnVec(vecY));
  CK(cusparseDestroy(handle));
  cudaFree(dBuffer);
  return 0;
}

}  // extern "C"
