/*
 * hispmv_synth.h -- deterministic synthetic matrices generated directly in HBM (benchmark / test support).
 *
 * BASELINE.json's configs are synthetic (no network, no SuiteSparse files): power-law graph-like (C2),
 * banded 27-point stencil (C4), uniform random (C5).  At 10^8..10^9 nonzeros they cannot be built on the
 * host and pushed over PCIe in reasonable time, so they are generated on the device from a counter-based
 * hash: every entry (row r, k-th nonzero of the row) is a pure function of (seed, r, k), which the CPU
 * oracle restates exactly (oracle/oracle.c: oracle_synth_*) for bit-exact checks of row_ptr / col / val.
 * Columns are produced already sorted within each row (stratified sampling through a monotone inverse
 * CDF), duplicates allowed -- the same CSR invariants the COO path produces.
 *
 * Not part of the reference's surface: the reference reads .mtx files (get_tb_matrices.py:57-82).
 */
#ifndef HISPMV_SYNTH_H_
#define HISPMV_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  /* row length = min(clip, K / (u32+1)) with u32 = hash(seed,row)>>32: P(len >= L) ~ (K/2^32)/L  (alpha = 2)
   * column     = floor(cols * q^gamma), q = (k + u)/len stratified  (gamma = 1 uniform, 5 ~ Zipf s = 0.8)
   * params = { K, clip, gamma | period << 8 }: period > 0 makes row r as long as row r % period (the entries still
   * depend on r), so N stacked blocks of `period` rows hold the same number of nonzeros each (weak scaling) */
  HISPMV_SYNTH_POWERLAW = 1,
  /* row length = base + popcount(hash & mask), columns uniform stratified.  params = { base, mask } */
  HISPMV_SYNTH_UNIFORM = 2,
  /* 27-point stencil on an nx*ny*nz grid (rows = cols = nx*ny*nz).  params = { nx, ny, nz } */
  HISPMV_SYNTH_STENCIL27 = 3
};

/* Row lengths of rows [row_begin,row_end) summed: the nonzero count of that block. */
int hispmv_synth_count(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int32_t row_begin,
                       int32_t row_end, int64_t* nnz);
/* nnz-balanced split points over the whole matrix (same rule as hispmv_shard_bounds, on 64-bit prefix sums). */
int hispmv_synth_shard_bounds(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int n_parts,
                              int32_t* bounds, int64_t* total_nnz);
/* Generate rows [row_begin,row_end) as CSR on the current device.  row_ptr is rebased to 0; col/val carry the
 * library's 128-bit padding.  Free the three arrays with hispmv_synth_free. */
int hispmv_synth_csr(int kind, uint64_t seed, int32_t rows, int32_t cols, const int64_t* params, int32_t row_begin,
                     int32_t row_end, int32_t** d_row_ptr, int32_t** d_col, float** d_val, int64_t* nnz);
void hispmv_synth_free(void* d_ptr);

#ifdef __cplusplus
}
#endif
#endif
