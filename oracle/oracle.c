/*
 * oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's algorithms on the SpMV/GeMV hot
 * path, plus single-threaded restatements of the GPU partitioner's integer artefacts.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  The product (hispmv_b200/) never links, imports or calls anything in oracle/.
 *
 * Pinning: the floating-point routines here are checked (tests/test_oracle_pins.py) against the
 * UNMODIFIED reference code compiled into oracle/_ref/ (cpu_spmv, mkl_spmv, naive_gemv, mkl_gemv,
 * cooToCsr, cpuSpMV, HiSpmvHandle::cpuSequential / tileAndPad) and against golden vectors generated from
 * those binaries (tests/golden/).  The reference holds no golden vectors of its own; MKL's
 * summation order is not specified, so parity against mkl_sparse_s_mv / sgemv is tolerance-based:
 * |y - y_ref| / (sum_j |a_ij||x_j| + |beta*y0_i|) <= 1e-5.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * COO -> CSR.  gpu/src/spmvHelper.cpp:117-156 (cooToCsr) and common/src/spmv-helper.cpp:139-227
 * (tileAndPad): group by row, sort each row's (col, val) pairs lexicographically (std::sort on
 * std::pair<int,float>), duplicates kept as separate entries.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int col;
  float val;
} colval_t;

static int cmp_colval(const void* a, const void* b) {
  const colval_t* x = (const colval_t*)a;
  const colval_t* y = (const colval_t*)b;
  if (x->col != y->col) return x->col < y->col ? -1 : 1;
  if (x->val < y->val) return -1;
  if (y->val < x->val) return 1;
  return 0;
}

int oracle_coo_to_csr(int rows, int64_t nnz, const int* coo_r, const int* coo_c, const float* coo_v, int* row_ptr,
                      int* col, float* val) {
  int64_t i;
  int r;
  int* fill;
  colval_t* tmp;
  memset(row_ptr, 0, sizeof(int) * ((size_t)rows + 1));
  for (i = 0; i < nnz; ++i) {
    if (coo_r[i] < 0 || coo_r[i] >= rows) return -1;
    row_ptr[coo_r[i] + 1]++;
  }
  for (r = 0; r < rows; ++r) row_ptr[r + 1] += row_ptr[r];
  fill = (int*)malloc(sizeof(int) * ((size_t)rows + 1));
  tmp = (colval_t*)malloc(sizeof(colval_t) * (size_t)(nnz ? nnz : 1));
  memcpy(fill, row_ptr, sizeof(int) * ((size_t)rows + 1));
  for (i = 0; i < nnz; ++i) {
    const int p = fill[coo_r[i]]++;
    tmp[p].col = coo_c[i];
    tmp[p].val = coo_v[i];
  }
  for (r = 0; r < rows; ++r) qsort(tmp + row_ptr[r], (size_t)(row_ptr[r + 1] - row_ptr[r]), sizeof(colval_t), cmp_colval);
  for (i = 0; i < nnz; ++i) {
    col[i] = tmp[i].col;
    val[i] = tmp[i].val;
  }
  free(fill);
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * fp32 SpMV in the reference's naive orders.
 * ---------------------------------------------------------------------------------------------- */
/* cpu/src/main.cpp:11-23 (cpu_spmv): C[i] *= beta; C[i] += alpha*value*B[col] in CSR order, in place. */
void oracle_spmv_csr_f32(int rows, const int* row_ptr, const int* col, const float* val, const float* x, float* y,
                         float alpha, float beta) {
  int i, j;
  for (i = 0; i < rows; ++i) {
    float c = y[i] * beta;
    for (j = row_ptr[i]; j < row_ptr[i + 1]; ++j) c += alpha * val[j] * x[col[j]];
    y[i] = c;
  }
}

/* common/src/spmv-helper.cpp:812-833 (cpuSequential, sparse branch): Cout[row] += value*B[col] in COO
 * order starting from Cout as given (the host passes zeros), then Cout = alpha*Cout + beta*Cin. */
void oracle_spmv_coo_f32(int rows, int64_t nnz, const int* coo_r, const int* coo_c, const float* coo_v, const float* x,
                         const float* c_in, float alpha, float beta, float* c_out) {
  int64_t i;
  int r;
  for (i = 0; i < nnz; ++i) c_out[coo_r[i]] += coo_v[i] * x[coo_c[i]];
  for (r = 0; r < rows; ++r) c_out[r] = (alpha * c_out[r]) + (beta * c_in[r]);
}

/* gpu/src/spmvHelper.cpp:223-233 (cpuSpMV): y *= beta; y[row] += alpha*value*x[col] in COO order. */
void oracle_spmv_coo_inplace_f32(int rows, int64_t nnz, const int* coo_r, const int* coo_c, const float* coo_v,
                                 const float* x, float* y, float alpha, float beta) {
  int64_t i;
  int r;
  for (r = 0; r < rows; ++r) y[r] *= beta;
  for (i = 0; i < nnz; ++i) y[coo_r[i]] += alpha * coo_v[i] * x[coo_c[i]];
}

/* cpu/src/main.cpp:53-71 (naive_gemv): acc += A[i][j]*x[j]; y = alpha*acc + beta*y. */
void oracle_gemv_f32(int rows, int cols, const float* a, const float* x, float* y, float alpha, float beta) {
  int i, j;
  for (i = 0; i < rows; ++i) {
    float acc = 0.0f;
    const float* ai = a + (size_t)i * cols;
    for (j = 0; j < cols; ++j) acc += ai[j] * x[j];
    y[i] = alpha * acc + beta * y[i];
  }
}

/* ------------------------------------------------------------------------------------------------
 * float64 reference + the normaliser of the stated tolerance:
 *   y64[i]   = alpha * sum_j a_ij x_j + beta * y0_i          (accumulated in double)
 *   scale[i] = |alpha| * sum_j |a_ij||x_j| + |beta*y0_i|
 * (semantics: automation_tool/assets/base_functions.cpp:535, y = beta*c_in + alpha*(A x)).
 * ---------------------------------------------------------------------------------------------- */
void oracle_spmv_csr_f64(int rows, const int* row_ptr, const int* col, const float* val, const float* x,
                         const float* y0, float alpha, float beta, double* y64, double* scale) {
  int i;
#pragma omp parallel for schedule(dynamic, 4096)
  for (i = 0; i < rows; ++i) {
    double s = 0.0, a = 0.0;
    int j;
    for (j = row_ptr[i]; j < row_ptr[i + 1]; ++j) {
      const double p = (double)val[j] * (double)x[col[j]];
      s += p;
      a += fabs(p);
    }
    {
      const double b = y0 ? (double)beta * (double)y0[i] : 0.0;
      y64[i] = (double)alpha * s + b;
      if (scale) scale[i] = fabs((double)alpha) * a + fabs(b);
    }
  }
}

void oracle_gemv_f64(int rows, int cols, const float* a, const float* x, const float* y0, float alpha, float beta,
                     double* y64, double* scale) {
  int i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < rows; ++i) {
    double s = 0.0, ab = 0.0;
    const float* ai = a + (size_t)i * cols;
    int j;
    for (j = 0; j < cols; ++j) {
      const double p = (double)ai[j] * (double)x[j];
      s += p;
      ab += fabs(p);
    }
    {
      const double b = y0 ? (double)beta * (double)y0[i] : 0.0;
      y64[i] = (double)alpha * s + b;
      if (scale) scale[i] = fabs((double)alpha) * ab + fabs(b);
    }
  }
}

/* max_i |y[i]-y64[i]| / scale[i]  (rows with scale == 0 must match exactly, else +inf) */
double oracle_max_scaled_error(int rows, const float* y, const double* y64, const double* scale, int* argmax) {
  double worst = 0.0;
  int i, at = -1;
  for (i = 0; i < rows; ++i) {
    const double d = fabs((double)y[i] - y64[i]);
    double e;
    if (scale[i] > 0.0) e = d / scale[i];
    else e = d == 0.0 ? 0.0 : INFINITY;
    if (!(e <= worst)) {  /* also catches NaN */
      worst = e;
      at = i;
    }
  }
  if (argmax) *argmax = at;
  return worst;
}

/* ------------------------------------------------------------------------------------------------
 * Partitioner restatements (single-threaded).  These restate hispmv_b200/csrc/partition.cu, the GPU
 * replacement for common/src/spmv-helper.cpp:265-347 (balanceWorkload) and :429-638 (schedule), and
 * automation_tool/src/dse.py:23-95 (config choice).  Integer-exact by construction.
 * ---------------------------------------------------------------------------------------------- */
#define ORACLE_HIST_BINS 33

/* hist[0]: empty rows; hist[k]: rows with 2^(k-1) <= nnz < 2^k */
void oracle_row_stats(int rows, const int* row_ptr, int64_t* hist, int* max_row_nnz, int* empty_rows) {
  int i, mx = 0;
  memset(hist, 0, sizeof(int64_t) * ORACLE_HIST_BINS);
  for (i = 0; i < rows; ++i) {
    const int len = row_ptr[i + 1] - row_ptr[i];
    int bin = 0, l = len;
    while (l > 0) {
      ++bin;
      l >>= 1;
    }
    hist[bin]++;
    if (len > mx) mx = len;
  }
  *max_row_nnz = mx;
  *empty_rows = (int)hist[0];
}

/* kernel ids as in include/hispmv.h */
enum { K_SCALAR = 1, K_VECTOR = 2, K_MERGE = 3, K_EMPTY = 5, K_ADAPTIVE = 6, K_ROWSTAGE = 7 };

/* Column-locality probe (partition.cu: col_probe_kernel): up to 8192 evenly spaced rows r >= 1, entries of rows r and
 * r-1 compared position by position over the first min(len(r), len(r-1), 32) positions. */
void oracle_col_probe(int rows, const int* rp, const int* col, int64_t* near, int64_t* cmp) {
  int64_t nr = 0, cp = 0;
  if (rows >= 2) {
    const int samples = rows - 1 < 8192 ? rows - 1 : 8192;
    int j, k;
    for (j = 0; j < samples; ++j) {
      const int64_t r = 1 + ((int64_t)j * (rows - 1)) / samples;
      const int a0 = rp[r - 1], a1 = rp[r], a2 = rp[r + 1];
      int n = a1 - a0 < a2 - a1 ? a1 - a0 : a2 - a1;
      if (n > 32) n = 32;
      for (k = 0; k < n; ++k) {
        const int d = col[a1 + k] - col[a0 + k];
        nr += (d <= 32 && d >= -32);
      }
      if (n > 0) cp += n;
    }
  }
  *near = nr;
  *cmp = cp;
}

/* partition.cu: select_kernel.  hist: 33 power-of-two bins as produced by oracle_row_stats. */
void oracle_select_kernel(int rows, int64_t nnz, const int64_t* hist, int64_t probe_near, int64_t probe_cmp,
                          int allow_split_rows, int* kernel, int* lanes) {
  int64_t mean, heavy_nnz = 0;
  int l = 2, k;
  *lanes = 0;
  if (rows <= 0 || nnz <= 0) {
    *kernel = K_EMPTY;
    return;
  }
  mean = (nnz + rows - 1) / rows;
  while (l < 32 && (int64_t)l * 2 <= mean) l *= 2;
  if (!allow_split_rows) {
    if (mean <= 2) {
      *kernel = K_SCALAR;
    } else {
      *kernel = K_VECTOR;
      *lanes = l;
    }
    return;
  }
  for (k = 1; k < 33; ++k) {
    const int64_t lo = (int64_t)1 << (k - 1);
    if (lo >= 4 * mean) heavy_nnz += hist[k] * (k == 1 ? 1 : 3 * ((int64_t)1 << (k - 2)));
  }
  {
    const int regular = heavy_nnz * 8 < nnz;
    const int banded = probe_cmp >= 64 && probe_near * 4 >= probe_cmp * 3;
    *kernel = (regular && banded && mean <= 64) ? K_ROWSTAGE : K_ADAPTIVE;
  }
}

int oracle_merge_tile_items(int rows, int64_t nnz) {
  const int64_t total = (int64_t)rows + nnz;
  if (total < 148LL * 4 * 1792) return 128 * 7;
  return 256 * 7;
}

/* Column slab width (partition.cu: select_slab_cols); 0 = no slabs. */
int oracle_select_slab_cols(int cols, int64_t nnz, int64_t probe_near, int64_t probe_cmp) {
  const int banded = probe_cmp >= 64 && probe_near * 4 >= probe_cmp * 3;
  int64_t n, w;
  if (banded || (int64_t)cols * 4 <= (64ll << 20) || nnz < 4000000) return 0;
  n = ((int64_t)cols + 12500000 - 1) / 12500000;
  w = ((int64_t)cols + n - 1) / n;
  return (int)((w + 31) & ~(int64_t)31);
}

/* ROWSTAGE plan parameters (partition.cu: rowstage_params): lanes per row, STREAM budget B, long threshold T,
 * chunk size CH.  lanes_in = 0 lets the rule choose lanes. */
void oracle_rowstage_params(int rows, int64_t nnz, int lanes_in, int* lanes, int* stream_items, int* long_threshold,
                            int* chunk_nnz) {
  const int64_t rws = rows > 1 ? rows : 1;
  const int64_t items = (nnz + rws + rws - 1) / rws;
  int l = lanes_in;
  int64_t b;
  if (l <= 0) {
    l = 1;
    while (l < 32 && (128 / l) * items > 1792) l *= 2;
  }
  b = (128 / l) * items;
  if (b > 8192 - 256) b = 8192 - 256;
  if (b < 128) b = 128;
  *lanes = l;
  *stream_items = (int)b;
  *long_threshold = 256;
  *chunk_nnz = 4096;
}

/* Merge-path tile start coordinates: tile t starts at diagonal min(t*tile_items, rows+nnz) of the merge
 * of the row-end list (row_ptr[1..rows]) with the nonzero indices 0..nnz-1.  num_tiles+1 entries. */
int64_t oracle_merge_tiles(int rows, const int* row_ptr, int tile_items, int* tile_row, int64_t* tile_nnz) {
  const int64_t nnz = row_ptr[rows];
  const int64_t total = (int64_t)rows + nnz;
  int64_t nt = (total + tile_items - 1) / tile_items, t;
  if (nt < 1) nt = 1;
  if (!tile_row) return nt;
  for (t = 0; t <= nt; ++t) {
    int64_t diag = t * (int64_t)tile_items, lo, hi;
    if (diag > total) diag = total;
    lo = diag > nnz ? diag - nnz : 0;
    hi = diag < rows ? diag : rows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)row_ptr[mid + 1] <= diag - mid - 1) lo = mid + 1;
      else hi = mid;
    }
    tile_row[t] = (int)lo;
    tile_nnz[t] = diag - lo;
  }
  return nt;
}

/* Rows cut by a tile boundary after at least one of their nonzeros, ascending, no repeats: the rows whose
 * partial sums travel through the carry-out array (analogue of balanceWorkload's shared-row list). */
int64_t oracle_split_rows(int rows, const int* row_ptr, int64_t num_tiles, const int* tile_row,
                          const int64_t* tile_nnz, int* out) {
  int64_t t, n = 0;
  int last = -1;
  for (t = 1; t < num_tiles; ++t) {
    const int r = tile_row[t];
    if (r < rows && tile_nnz[t] > (int64_t)row_ptr[r] && r != last) {
      if (out) out[n] = r;
      ++n;
      last = r;
    }
  }
  return n;
}

/* Adaptive row-aligned tiles (hispmv_b200/csrc/partition.cu: adaptive_tiles_device).  Short rows
 * (len < T) are packed, in order, into STREAM tiles of about B merge items (1 + len each); a tile is cut
 * where the running item count crosses a multiple of B, and around every long row.  A long row becomes
 * ceil(len / CH) LONG tiles.  tile_row gets num_tiles+1 entries (last = rows), tile_chunk num_tiles entries
 * (-1 for STREAM, chunk index for LONG).  Call with tile_row == NULL to count. */
int64_t oracle_adaptive_tiles(int rows, const int* row_ptr, int B, int T, int CH, int* tile_row, int* tile_chunk) {
  int64_t nt = 0, S = 0, S_prev = 0;
  int r, prev_long = 0;
  for (r = 0; r < rows; ++r) {
    const int len = row_ptr[r + 1] - row_ptr[r];
    if (len >= T) {
      const int c = (len + CH - 1) / CH;
      int k;
      for (k = 0; k < c; ++k) {
        if (tile_row) {
          tile_row[nt] = r;
          tile_chunk[nt] = k;
        }
        ++nt;
      }
      prev_long = 1;
      S_prev = S; /* w_r = 0 */
    } else {
      if (r == 0 || prev_long || (S / B != S_prev / B)) {
        if (tile_row) {
          tile_row[nt] = r;
          tile_chunk[nt] = -1;
        }
        ++nt;
      }
      prev_long = 0;
      S_prev = S;
      S += 1 + (int64_t)len;
    }
  }
  if (tile_row) tile_row[nt] = rows;
  return nt;
}

/* rows split across several LONG tiles (len >= T and more than one chunk), ascending */
int64_t oracle_adaptive_split_rows(int rows, const int* row_ptr, int T, int CH, int* out) {
  int64_t n = 0;
  int r;
  for (r = 0; r < rows; ++r) {
    const int len = row_ptr[r + 1] - row_ptr[r];
    if (len >= T && (len + CH - 1) / CH >= 2) {
      if (out) out[n] = r;
      ++n;
    }
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Blocked (column-slab x row-panel) plan, restating hispmv_b200/csrc/blocked.cu: pb_build_device,
 * pb_make_work and select_blocked.  The reference's own column tiling is tileAndPad
 * (common/src/spmv-helper.cpp:139-227: a tile covers as many columns as the on-chip x buffers hold);
 * this is the same idea with the GPU engine's sizes, single-threaded.
 *   slab s        = columns [s*W, (s+1)*W)
 *   blocked order : slab-major, CSR order inside a slab; every slab starts at a multiple of G (the group
 *                   size, 512 = 32 lanes x 16 entries); the gaps are padding (val 0, lcol 0, no flag)
 *   per entry     : val, lcol = col - s*W, and an end flag (16 per word, bit j = entry 16*i+j) on the last
 *                   entry of every PIECE = maximal run of consecutive entries of one row inside one group
 *   storage       : inside a group, entry e = 16*lane + w is stored where a warp's coalesced vector loads hand
 *                   lane `lane` its 16 consecutive entries: val at ((w/4)*32 + lane)*4 + w%4, lcol at
 *                   ((w/8)*32 + lane)*8 + w%8 (pb_phys_val / pb_phys_lcol); flags and piece numbers follow the
 *                   logical order
 *   pieces        : numbered in blocked order; group_base[g] = pieces ending before group g;
 *                   prow_ptr = CSR-style offsets of the pieces of every row (a row's pieces in slab order);
 *                   piece_pcsr[q] = position of piece q in that per-row order, piece_slab[q] its slab
 * oracle_pb_order returns the padded length (sizes with o_val == NULL); *num_pieces gets the piece count.
 * ---------------------------------------------------------------------------------------------- */
static int64_t pb_phys_val(int64_t k, int G) {
  const int64_t g = k / G, e = k % G;
  const int per = G / 32;
  const int64_t lane = e / per, w = e % per;
  return g * G + ((w / 4) * 32 + lane) * 4 + w % 4;
}
static int64_t pb_phys_lcol(int64_t k, int G) {
  const int64_t g = k / G, e = k % G;
  const int per = G / 32;
  const int64_t lane = e / per, w = e % per;
  return g * G + ((w / 8) * 32 + lane) * 8 + w % 8;
}

int64_t oracle_pb_order(int rows, int cols, const int* row_ptr, const int* col, const float* val, int W, int G,
                        int* slab_ptr, float* o_val, uint16_t* o_lcol, uint16_t* o_flags, int* group_base,
                        int* prow_ptr, int* piece_pcsr, int* piece_slab, int64_t* num_pieces) {
  const int64_t nnz = row_ptr[rows];
  const int S = (int)(((int64_t)cols + W - 1) / W);
  int64_t* cnt = (int64_t*)calloc((size_t)S + 1, sizeof(int64_t));
  int64_t* start = (int64_t*)calloc((size_t)S + 1, sizeof(int64_t));
  int64_t j, pos = 0, k, np = 0;
  int* brow; /* row of every blocked position, -1 for padding */
  int* fill;
  int s, r;
  for (j = 0; j < nnz; ++j) cnt[col[j] / W]++;
  for (s = 0; s < S; ++s) {
    start[s] = pos;
    if (slab_ptr) slab_ptr[s] = (int)pos;
    pos += cnt[s];
    pos = (pos + G - 1) / G * G;
    cnt[s] = 0; /* becomes the fill counter */
  }
  if (slab_ptr) slab_ptr[S] = (int)pos;
  brow = (int*)malloc(sizeof(int) * (size_t)(pos + 1));
  for (k = 0; k <= pos; ++k) brow[k] = -1;
  if (o_val) {
    memset(o_val, 0, sizeof(float) * (size_t)pos);
    memset(o_lcol, 0, sizeof(uint16_t) * (size_t)pos);
    memset(o_flags, 0, sizeof(uint16_t) * (size_t)(pos / 16));
  }
  for (r = 0; r < rows; ++r)
    for (j = row_ptr[r]; j < row_ptr[r + 1]; ++j) {
      const int sl = col[j] / W;
      const int64_t dst = start[sl] + cnt[sl]++;
      brow[dst] = r;
      if (o_val) {
        o_val[pb_phys_val(dst, G)] = val[j];
        o_lcol[pb_phys_lcol(dst, G)] = (uint16_t)(col[j] - sl * W);
      }
    }
  /* pieces: an entry ends one when the next position is another row, padding, or the next group */
  if (prow_ptr) memset(prow_ptr, 0, sizeof(int) * ((size_t)rows + 1));
  for (k = 0; k < pos; ++k) {
    if (k % G == 0 && group_base) group_base[k / G] = (int)np;
    if (brow[k] < 0) continue;
    if ((k + 1) % G == 0 || brow[k + 1] != brow[k]) {
      if (o_flags) o_flags[k / 16] |= (uint16_t)(1u << (k % 16));
      if (prow_ptr) prow_ptr[brow[k] + 1]++;
      ++np;
    }
  }
  if (group_base) group_base[pos / G] = (int)np;
  if (num_pieces) *num_pieces = np;
  if (prow_ptr && piece_pcsr) {
    int64_t q = 0;
    for (r = 0; r < rows; ++r) prow_ptr[r + 1] += prow_ptr[r];
    fill = (int*)calloc((size_t)rows + 1, sizeof(int));
    s = 0;
    for (k = 0; k < pos; ++k) {
      while (s + 1 < S && k >= start[s + 1]) ++s;
      if (brow[k] < 0) continue;
      if ((k + 1) % G == 0 || brow[k + 1] != brow[k]) {
        piece_pcsr[q] = prow_ptr[brow[k]] + fill[brow[k]]++;
        piece_slab[q] = s;
        ++q;
      }
    }
    free(fill);
  } else if (prow_ptr) {
    for (r = 0; r < rows; ++r) prow_ptr[r + 1] += prow_ptr[r];
  }
  free(brow);
  free(cnt);
  free(start);
  return pos;
}

static void pb_panel_extent(const int* row_ptr, const int* tile_row, const int* tile_chunk, int CH, int64_t t,
                            int* n0, int* n1) {
  if (tile_chunk[t] >= 0) {
    const int r = tile_row[t];
    const int b = row_ptr[r] + tile_chunk[t] * CH;
    const int e = row_ptr[r + 1];
    *n0 = b;
    *n1 = b + CH < e ? b + CH : e;
  } else {
    *n0 = row_ptr[tile_row[t]];
    *n1 = row_ptr[tile_row[t + 1]];
  }
}

/* Panels are the adaptive tiles over prow_ptr (pieces instead of nonzeros).  For every piece: perm = its per-row-order
 * position minus the first position of its panel.  Segments = the non-empty (panel, slab) runs of consecutive piece
 * ids, listed panel-major then by slab, as (first piece id, number of the panel's pieces in earlier slabs).
 * Returns the segment count (sizes with seg == NULL). */
int64_t oracle_pb_segments(int64_t num_pieces, const int* piece_pcsr, const int* piece_slab, int S, const int* prow_ptr,
                           int64_t num_panels, const int* tile_row, const int* tile_chunk, int CH, uint16_t* perm,
                           int* panel_seg, int* seg) {
  /* panel of every per-row-order position */
  const int64_t total = num_pieces;
  int* pan_of = (int*)malloc(sizeof(int) * (size_t)(total + 1));
  int* n0s = (int*)malloc(sizeof(int) * (size_t)(num_panels + 1));
  int64_t t, q, nseg = 0;
  for (t = 0; t < num_panels; ++t) {
    int n0, n1, i;
    pb_panel_extent(prow_ptr, tile_row, tile_chunk, CH, t, &n0, &n1);
    n0s[t] = n0;
    for (i = n0; i < n1; ++i) pan_of[i] = (int)t;
  }
  for (q = 0; q < total; ++q)
    if (perm) perm[q] = (uint16_t)(piece_pcsr[q] - n0s[pan_of[piece_pcsr[q]]]);
  /* runs of equal (slab, panel) in piece-id order; collect them, then order by (panel, slab) with a counting pass */
  {
    int64_t nrun = 0, i;
    int* run_start;
    int* run_pan;
    int* run_slab;
    int* run_len;
    int64_t* pcount = (int64_t*)calloc((size_t)num_panels + 1, sizeof(int64_t));
    for (q = 0; q < total; ++q)
      if (q == 0 || piece_slab[q] != piece_slab[q - 1] || pan_of[piece_pcsr[q]] != pan_of[piece_pcsr[q - 1]]) ++nrun;
    run_start = (int*)malloc(sizeof(int) * (size_t)(nrun + 1));
    run_pan = (int*)malloc(sizeof(int) * (size_t)(nrun + 1));
    run_slab = (int*)malloc(sizeof(int) * (size_t)(nrun + 1));
    run_len = (int*)malloc(sizeof(int) * (size_t)(nrun + 1));
    nrun = 0;
    for (q = 0; q < total; ++q) {
      if (q == 0 || piece_slab[q] != piece_slab[q - 1] || pan_of[piece_pcsr[q]] != pan_of[piece_pcsr[q - 1]]) {
        run_start[nrun] = (int)q;
        run_pan[nrun] = pan_of[piece_pcsr[q]];
        run_slab[nrun] = piece_slab[q];
        run_len[nrun] = 0;
        ++nrun;
      }
      run_len[nrun - 1]++;
    }
    nseg = nrun;
    /* runs come slab-major and, inside a slab, by panel: a stable counting sort by panel gives (panel, slab) order */
    for (i = 0; i < nrun; ++i) pcount[run_pan[i] + 1]++;
    for (t = 0; t < num_panels; ++t) pcount[t + 1] += pcount[t];
    if (panel_seg)
      for (t = 0; t <= num_panels; ++t) panel_seg[t] = (int)pcount[t];
    if (seg) {
      int64_t* at = (int64_t*)malloc(sizeof(int64_t) * ((size_t)num_panels + 1));
      int* off = (int*)calloc((size_t)num_panels + 1, sizeof(int));
      memcpy(at, pcount, sizeof(int64_t) * ((size_t)num_panels + 1));
      for (i = 0; i < nrun; ++i) {
        const int p = run_pan[i];
        const int64_t d = at[p]++;
        seg[2 * d] = run_start[i];
        seg[2 * d + 1] = off[p];
        off[p] += run_len[i];
      }
      free(at);
      free(off);
    }
    free(run_start);
    free(run_pan);
    free(run_slab);
    free(run_len);
    free(pcount);
  }
  free(pan_of);
  free(n0s);
  return nseg;
}

/* Pass-1 work ranges (blocked.cu: pb_make_work): n_cta contiguous runs of groups of the blocked order, balanced by
 * entries (align per group) + piece_cost16/16 per piece a group ends + slab_cost for every slab a range is the first
 * to touch.  group_base has total/align + 1 entries (oracle_pb_order).  work gets 2*n_cta ints. */
void oracle_pb_work(int S, const int* slab_ptr, int align, const int* group_base, int n_cta, int64_t slab_cost,
                    int64_t piece_cost16, int* work) {
  const int64_t ngroups = slab_ptr[S] / align;
  int64_t used = 0, remaining, g = 0, i;
  int s = 0, b;
  for (b = 0; b < S; ++b) used += slab_ptr[b + 1] > slab_ptr[b];
  remaining = slab_cost * used;
  for (i = 0; i < ngroups; ++i) remaining += align + piece_cost16 * (group_base[i + 1] - group_base[i]) / 16;
  for (b = 0; b < n_cta; ++b) {
    const int64_t g0 = g;
    int64_t budget = (remaining + (n_cta - b) - 1) / (n_cta - b), spent = 0;
    int fresh = 1;
    while (g < ngroups && (budget > 0 || b == n_cta - 1)) {
      const int64_t k = g * align;
      int64_t c;
      while (s < S && slab_ptr[s + 1] <= k) {
        ++s;
        fresh = 1;
      }
      if (s >= S) break;
      if (fresh) {
        if (k == slab_ptr[s]) {
          budget -= slab_cost;
          spent += slab_cost;
        }
        fresh = 0;
        if (budget <= 0 && g > g0 && b != n_cta - 1) break;
      }
      c = align + piece_cost16 * (group_base[g + 1] - group_base[g]) / 16;
      ++g;
      budget -= c;
      spent += c;
    }
    if (b == n_cta - 1) g = ngroups;
    work[2 * b] = (int)(g0 * align);
    work[2 * b + 1] = (int)(g * align);
    remaining -= spent;
    if (remaining < 0) remaining = 0;
  }
}

/* blocked.cu: pb_count_runs_device -- (row, slab) runs: the piece count before group boundaries split any */
int64_t oracle_pb_count_runs(int rows, const int* row_ptr, const int* col, int W) {
  int64_t runs = 0;
  int r, j;
  for (r = 0; r < rows; ++r)
    for (j = row_ptr[r]; j < row_ptr[r + 1]; ++j)
      if (j == row_ptr[r] || col[j] / W != col[j - 1] / W) ++runs;
  return runs;
}

/* blocked.cu: select_blocked -- scattered columns over an x beyond an SM's L1 (>= 100 000 columns), a matrix big enough for two launches, and rows
 * concentrated enough that the (row, slab) runs are at most 0.6 of the nonzeros */
int oracle_select_blocked(int rows, int cols, int64_t nnz, int64_t slab_runs, int64_t probe_near, int64_t probe_cmp,
                          int allow_split_rows) {
  const int banded = probe_cmp >= 64 && probe_near * 4 >= probe_cmp * 3;
  if (!allow_split_rows || banded || rows <= 0) return 0;
  if ((int64_t)cols < 100000 || nnz < 16000000) return 0;
  if (((int64_t)cols + 49152 - 1) / 49152 > 4096) return 0;
  return slab_runs * 5 <= nnz * 3 ? 1 : 0;
}

/* nnz-balanced contiguous row blocks: bounds[k] = first row r with row_ptr[r] >= k*nnz/n_parts. */
void oracle_shard_bounds(int rows, const int* row_ptr, int n_parts, int* bounds) {
  const int64_t nnz = row_ptr[rows];
  int k;
  bounds[0] = 0;
  bounds[n_parts] = rows;
  for (k = 1; k < n_parts; ++k) {
    const int64_t target = (nnz * (int64_t)k) / n_parts;
    int64_t lo = 0, hi = rows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)row_ptr[mid] < target) lo = mid + 1;
      else hi = mid;
    }
    bounds[k] = (int)lo;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic matrix generators (restating hispmv_b200/csrc/synth.cu; see include/hispmv_synth.h).
 * ---------------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint64_t hash_row(uint64_t seed, int64_t r) { return mix64(mix64(seed) ^ (uint64_t)r); }
static uint64_t hash_entry(uint64_t seed, int64_t r, int64_t k) {
  return mix64(hash_row(seed ^ 0xA5A5A5A5A5A5A5A5ull, r) + (uint64_t)k * 0xD1342543DE82EF95ull);
}

enum { SYNTH_POWERLAW = 1, SYNTH_UNIFORM = 2, SYNTH_STENCIL27 = 3 };

int oracle_synth_row_len(int kind, uint64_t seed, const int64_t* p, int64_t r) {
  if (kind == SYNTH_POWERLAW) {
    const int64_t period = p[2] >> 8; /* > 0: the row-length sequence repeats every `period` rows */
    const uint64_t u = hash_row(seed, period > 0 ? r % period : r) >> 32;
    const uint64_t len = (uint64_t)p[0] / (u + 1);
    return (int)(len < (uint64_t)p[1] ? len : (uint64_t)p[1]);
  }
  if (kind == SYNTH_UNIFORM) {
    const uint64_t h = hash_row(seed, r) & (uint64_t)p[1];
    return (int)(p[0] + __builtin_popcountll(h));
  }
  if (kind == SYNTH_STENCIL27) {
    const int64_t nx = p[0], ny = p[1], nz = p[2];
    const int64_t ix = r % nx, iy = (r / nx) % ny, iz = r / (nx * ny);
    const int cx = 1 + (ix > 0) + (ix < nx - 1);
    const int cy = 1 + (iy > 0) + (iy < ny - 1);
    const int cz = 1 + (iz > 0) + (iz < nz - 1);
    return cx * cy * cz;
  }
  return 0;
}

void oracle_synth_entry(int kind, uint64_t seed, int cols, const int64_t* p, int64_t r, int k, int len, int* col,
                        float* val) {
  const uint64_t h = hash_entry(seed, r, k);
  *val = (float)((int32_t)(h & 0xFFFFFF) - 0x800000) * (1.0f / 8388608.0f);
  if (kind == SYNTH_STENCIL27) {
    const int64_t nx = p[0], ny = p[1];
    const int64_t ix = r % nx, iy = (r / nx) % ny, iz = r / (nx * ny);
    const int cx = 1 + (ix > 0) + (ix < nx - 1);
    const int cy = 1 + (iy > 0) + (iy < ny - 1);
    const int a = k / (cy * cx), b = (k / cx) % cy, c = k % cx;
    const int64_t dz = a - (iz > 0), dy = b - (iy > 0), dx = c - (ix > 0);
    *col = (int)(r + dz * nx * ny + dy * nx + dx);
    return;
  }
  {
    /* -ffp-contract=off (oracle/Makefile): every operation below is a single IEEE double operation */
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    const double q = ((double)k + u) / (double)len;
    double w = q;
    const int gamma = kind == SYNTH_POWERLAW ? (int)(p[2] & 0xFF) : 1;
    int g;
    int64_t c;
    for (g = 1; g < gamma; ++g) w = w * q;
    c = (int64_t)(w * (double)cols);
    if (c >= cols) c = cols - 1;
    *col = (int)c;
  }
}

/* whole row block [row_begin,row_end) as CSR (row_ptr rebased to 0).  Call with col == NULL to size. */
int64_t oracle_synth_csr(int kind, uint64_t seed, int cols, const int64_t* p, int row_begin, int row_end, int* row_ptr,
                         int* col, float* val) {
  int64_t nnz = 0;
  int r;
  for (r = row_begin; r < row_end; ++r) {
    const int len = oracle_synth_row_len(kind, seed, p, r);
    if (row_ptr) row_ptr[r - row_begin] = (int)nnz;
    if (col) {
      int k;
      for (k = 0; k < len; ++k) oracle_synth_entry(kind, seed, cols, p, r, k, len, col + nnz + k, val + nnz + k);
    }
    nnz += len;
  }
  if (row_ptr) row_ptr[row_end - row_begin] = (int)nnz;
  return nnz;
}

/* ------------------------------------------------------------------------------------------------
 * Matrix Market reader, restating HiSpmvHandle::loadMtx (common/src/spmv-helper.cpp:34-136): banner
 * check, comment skip, 1-based indices, pattern -> 1.0f, zeros dropped, symmetric / skew-symmetric
 * expansion (mirror entry right after the original).  Two-pass: call with outputs NULL to count.
 * Returns the entry count or a negative error.
 * ---------------------------------------------------------------------------------------------- */
int64_t oracle_load_mtx(const char* path, int* rows, int* cols, int* coo_r, int* coo_c, float* coo_v) {
  FILE* f = fopen(path, "r");
  char line[1024], h0[64], h1[64], h2[64], h3[64], h4[64];
  int pattern, symm, skew;
  long long nr, nc, nz;
  int64_t n = 0;
  if (!f) return -1;
  if (!fgets(line, sizeof(line), f)) { fclose(f); return -2; }
  h0[0] = h1[0] = h2[0] = h3[0] = h4[0] = 0;
  sscanf(line, "%63s %63s %63s %63s %63s", h0, h1, h2, h3, h4);
  if (strcmp(h0, "%%MatrixMarket") || strcmp(h1, "matrix")) { fclose(f); return -2; }
  if (strcmp(h2, "coordinate")) { fclose(f); return -3; }
  pattern = !strcmp(h3, "pattern");
  if (strcmp(h3, "real") && strcmp(h3, "integer") && !pattern) { fclose(f); return -4; }
  symm = !strcmp(h4, "symmetric");
  skew = !strcmp(h4, "skew-symmetric");
  if (strcmp(h4, "general") && !symm && !skew) { fclose(f); return -5; }
  do {
    if (!fgets(line, sizeof(line), f)) { fclose(f); return -6; }
  } while (line[0] == '%');
  if (sscanf(line, "%lld %lld %lld", &nr, &nc, &nz) != 3) { fclose(f); return -6; }
  *rows = (int)nr;
  *cols = (int)nc;
  while (fgets(line, sizeof(line), f)) {
    long r, c;
    float v = 1.0f;
    char* q = line;
    char* e;
    r = strtol(q, &e, 10);
    if (e == q) continue;
    q = e;
    c = strtol(q, &e, 10);
    if (e == q) continue;
    q = e;
    if (!pattern) {
      v = strtof(q, &e);
      if (e == q) continue;
    }
    if (v == 0) continue;
    if (coo_r) { coo_r[n] = (int)r - 1; coo_c[n] = (int)c - 1; coo_v[n] = v; }
    ++n;
    if ((symm || skew) && r != c) {
      if (coo_r) { coo_r[n] = (int)c - 1; coo_c[n] = (int)r - 1; coo_v[n] = skew ? -v : v; }
      ++n;
    }
  }
  fclose(f);
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Packed PEG stream decoder / emulator.  The reference host library packs a matrix into
 * num_ch_A streams of 64-bit words, pes_per_ch PEs interleaved per channel
 *     word = rowEnd[63] row15[62:48] tileEnd[47] shared[46] col14[45:32] val32[31:0]
 * (encode, common/include/spmv-helper.h:45-60; layout written by prepareTile, common/src/spmv-helper.cpp:517-638:
 * word k of PE p of tile t sits at channel p / pes_per_ch, address tile_offset + k * pes_per_ch + p % pes_per_ch).
 * The accelerator consumes it as restated here:
 *   ComputeAB       (automation_tool/assets/base_functions.cpp:228-241)  product = val * x_window[col14]; the flags
 *                   of a PE pair come from the EVEN PE's word
 *   PreAccumulator  (:295-305,330)  shared pair: destination bank = row field of the even PE (low bits), row16 = row
 *                   field of the odd PE; a word is a dummy when its rowEnd bit is clear (every real entry has it set)
 *   ADD / SWB / SSW (:356-436)      partials of one shared row from all PEs are tree-added, then routed to the bank
 *                   (restated as a pairwise tree over adjacent PEs; the exact stage order is the crossbar
 *                   generator's, so the emulation is held to the tolerance, not to the bit)
 *   AccumBuffer     (:483-485)      BUFF_C[row16] += val in stream order, one buffer per PE, dumped per row tile
 *   Compute_C       (:535)          y = beta * c_in + alpha * acc
 * `stream` is num_ch * words_per_ch words, channel-major.  Local row of a private word of PE p: row15 * num_pes + p;
 * of a shared pair: row15(odd) * num_pes + rowfield(even) (prepareTile :560-563,612); global = tile index * tile size.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int valid, shared, row, col;
  float val;
} peg_item_t;

static inline uint64_t peg_word(const uint64_t* stream, int64_t words_per_ch, int pes_per_ch, int pe, int64_t slot) {
  return stream[(int64_t)(pe / pes_per_ch) * words_per_ch + slot * pes_per_ch + pe % pes_per_ch];
}

/* Decodes slot `slot` of PE pair g into items[0..1]; returns the pair's tileEnd flag. */
static int peg_decode_pair(const uint64_t* stream, int64_t words_per_ch, int pes_per_ch, int num_pes, int g,
                           int64_t slot, int row_base, int col_base, peg_item_t* items) {
  int log2pes = 0;
  while ((1 << log2pes) < num_pes) ++log2pes;
  const uint64_t w0 = peg_word(stream, words_per_ch, pes_per_ch, 2 * g, slot);
  const uint64_t w1 = peg_word(stream, words_per_ch, pes_per_ch, 2 * g + 1, slot);
  const int shared = (int)((w0 >> 46) & 1), tile_end = (int)((w0 >> 47) & 1);
  const uint64_t w[2] = {w0, w1};
  const int rf0 = (int)((w0 >> 48) & 0xFFFF), rf1 = (int)((w1 >> 48) & 0xFFFF);
  for (int p = 0; p < 2; ++p) {
    const uint32_t bits = (uint32_t)(w[p] & 0xFFFFFFFFu);
    const int rf = (int)((w[p] >> 48) & 0xFFFF);
    peg_item_t* it = &items[p];
    memcpy(&it->val, &bits, 4);
    it->col = col_base + (int)((w[p] >> 32) & 0x3FFF);
    it->shared = shared;
    if (shared) {
      it->valid = (rf1 >> 15) & 1;
      it->row = row_base + (rf1 & 0x7FFF) * num_pes + (rf0 & ((1 << log2pes) - 1));
    } else {
      it->valid = (rf >> 15) & 1;
      it->row = row_base + (rf & 0x7FFF) * num_pes + (2 * g + p);
    }
  }
  return tile_end;
}

/* Every real (non-dummy) word as (global row, global col, value, shared flag), in slot-major / PE-minor order.
 * Returns the count (arrays may be NULL to count only), -2 if the PE pairs disagree on a tile boundary. */
int64_t oracle_peg_decode(const uint64_t* stream, int num_ch, int pes_per_ch, int64_t words_per_ch, int tile_rows,
                          int tile_cols, int col_tiles, int32_t* row, int32_t* col, float* val, uint8_t* shared) {
  const int num_pes = num_ch * pes_per_ch;
  const int64_t slots = words_per_ch / pes_per_ch;
  int64_t n = 0;
  int tile = 0;
  for (int64_t s = 0; s < slots; ++s) {
    int ends = 0;
    for (int g = 0; g < num_pes / 2; ++g) {
      peg_item_t it[2];
      ends += peg_decode_pair(stream, words_per_ch, pes_per_ch, num_pes, g, s, (tile / col_tiles) * tile_rows,
                              (tile % col_tiles) * tile_cols, it);
      for (int p = 0; p < 2; ++p) {
        if (!it[p].valid) continue;
        if (row) { row[n] = it[p].row; col[n] = it[p].col; val[n] = it[p].val; shared[n] = (uint8_t)it[p].shared; }
        ++n;
      }
    }
    if (ends != 0 && ends != num_pes / 2) return -2;
    if (ends) ++tile;
  }
  return n;
}

/* y = beta * c_in + alpha * (A x) computed the way the accelerator walks the stream (fp32, -ffp-contract=off).
 * Returns the number of tiles seen, or -2 / -3 on a malformed stream (tile flags disagree / row out of range). */
int oracle_peg_spmv(const uint64_t* stream, int num_ch, int pes_per_ch, int64_t words_per_ch, int tile_rows,
                    int tile_cols, int col_tiles, int rows, int cols, const float* x, const float* c_in, float alpha,
                    float beta, float* y) {
  const int num_pes = num_ch * pes_per_ch;
  const int64_t slots = words_per_ch / pes_per_ch;
  float* acc = (float*)calloc((size_t)(rows > 0 ? rows : 1), sizeof(float));
  float* part = (float*)malloc(sizeof(float) * (size_t)num_pes);
  int* prow = (int*)malloc(sizeof(int) * (size_t)num_pes);
  int tile = 0, status = 0;
  for (int64_t s = 0; s < slots && status == 0; ++s) {
    int ends = 0;
    for (int pe = 0; pe < num_pes; ++pe) prow[pe] = -1;
    for (int g = 0; g < num_pes / 2; ++g) {
      peg_item_t it[2];
      ends += peg_decode_pair(stream, words_per_ch, pes_per_ch, num_pes, g, s, (tile / col_tiles) * tile_rows,
                              (tile % col_tiles) * tile_cols, it);
      for (int p = 0; p < 2; ++p) {
        if (!it[p].valid) continue;
        if (it[p].row < 0 || it[p].row >= rows || it[p].col < 0 || it[p].col >= cols) {
          if (it[p].val == 0.0f) continue; /* zero fill of a shared row's last stripe / padded rows */
          status = -3;
          break;
        }
        const float prod = it[p].val * x[it[p].col];                     /* ComputeAB */
        if (it[p].shared) {
          part[2 * g + p] = prod;                                        /* goes through the ADD tree below */
          prow[2 * g + p] = it[p].row;
        } else {
          acc[it[p].row] = prod + acc[it[p].row];                        /* AccumBuffer: val + BUFF_C[row] */
        }
      }
    }
    for (int stride = 1; stride < num_pes; stride *= 2)                  /* ADD stages: adjacent partners */
      for (int a = 0; a + stride < num_pes; a += 2 * stride)
        if (prow[a] >= 0 && prow[a + stride] == prow[a]) {
          part[a] = part[a] + part[a + stride];
          prow[a + stride] = -1;
        }
    for (int pe = 0; pe < num_pes; ++pe)
      if (prow[pe] >= 0) acc[prow[pe]] = part[pe] + acc[prow[pe]];
    if (ends != 0 && ends != num_pes / 2) status = -2;
    if (ends) ++tile;
  }
  if (status == 0)
    for (int i = 0; i < rows; ++i) y[i] = (beta * c_in[i]) + (alpha * acc[i]);   /* Compute_C :535 */
  free(acc);
  free(part);
  free(prow);
  return status ? status : tile;
}
