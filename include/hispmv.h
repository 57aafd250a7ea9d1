/*
 * hispmv.h -- C-ABI of libhispmv_cuda.so: the B200 (sm_100a) SpMV / GeMV engine behind HiSpMV's plugin surface.
 *
 * This is the drop-in boundary.  Everything above it (the `pyhispmv` pybind11 module, the apps/ scripts,
 * a cgo/ctypes/JNI binding) sees plain pointers and sizes; everything below it is hand-written CUDA.
 * There is no CPU fallback and no backend dispatch: every entry point fails with HISPMV_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo):
 *
 *   hispmv_create            FpgaHandle::FpgaHandle                 pyhispmv/src/fpga_handle.cpp:40-154
 *   hispmv_destroy           (reference never frees; no dtor)       pyhispmv/include/fpga_handle.h:9-74
 *   hispmv_add_sparse_coo    FpgaHandle::createSparseMtxHandle      pyhispmv/src/fpga_handle.cpp:156-207
 *                            -> HiSpmvHandle::prepareSparseMtxForFPGA common/src/spmv-helper.cpp:648-715
 *   hispmv_add_sparse_csr    same, for callers that already hold CSR (cpu/src/main.cpp:26-32)
 *   hispmv_add_dense         FpgaHandle::createDenseMtxHandle       pyhispmv/src/fpga_handle.cpp:209-250
 *                            -> HiSpmvHandle::prepareDenseMtxForFPGA common/src/spmv-helper.cpp:717-750
 *   hispmv_commit            FpgaHandle::loadMatrices               pyhispmv/src/fpga_handle.cpp:252-264
 *   hispmv_select            FpgaHandle::selectMatrix               pyhispmv/src/fpga_handle.cpp:266-283
 *   hispmv_run               FpgaHandle::runKernel                  pyhispmv/src/fpga_handle.cpp:286-321
 *   hispmv_linear            FpgaHandle::runLinear                  pyhispmv/src/fpga_handle.cpp:323-388
 *   hispmv_run_dev           the SpMV() kernel invocation itself    automation_tool/assets/top_function.cpp:1-47,
 *                            y = beta*c_in + alpha*(A x)            automation_tool/assets/base_functions.cpp:535
 *   hispmv_plan_* / hispmv_matrix_info
 *                            balanceWorkload's shared-row list and the tiling facts the reference prints
 *                            common/src/spmv-helper.cpp:242-347; automation_tool/src/dse.py:23-95 (selector)
 *   hispmv_load_mtx          HiSpmvHandle::loadMtx                  common/src/spmv-helper.cpp:34-136
 *
 * Semantics kept from the reference: COO input is unsorted, duplicates are kept as separate entries and
 * therefore summed, explicit zeros are kept, indices are int32 and values fp32; handle indices are handed
 * out in creation order; -1 means "device memory full" (fpga_handle.cpp:192-195,235-238);
 * run computes y = alpha * A x + beta * bias (base_functions.cpp:535); linear uses alpha = beta = 1
 * (fpga_handle.cpp:351-352) on len(x)/cols vectors one after the other.
 * Differences (documented in INTEGRATION.md): errors are returned, never exit(); commit is idempotent and
 * handles may be added after it; destroy frees device memory.
 */
#ifndef HISPMV_H_
#define HISPMV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hispmv_ctx hispmv_ctx;

/* status codes (negative = failure).  HISPMV_FULL keeps the reference's -1 sentinel. */
enum {
  HISPMV_OK = 0,
  HISPMV_FULL = -1,         /* device memory exhausted while adding a matrix */
  HISPMV_ERR_ARG = -2,      /* bad argument / misuse (reference: assert) */
  HISPMV_ERR_INDEX = -3,    /* matrix index out of range (reference: exit(1), fpga_handle.cpp:267-270) */
  HISPMV_ERR_CUDA = -4,     /* CUDA runtime failure, no usable device */
  HISPMV_ERR_STATE = -5,    /* run before select, dense handle while dense_overlay is off, ... */
  HISPMV_ERR_IO = -6        /* Matrix Market file problems */
};

/* kernel strategies the runtime selector chooses between (hispmv_matrix_info.kernel) */
enum {
  HISPMV_KERNEL_AUTO = 0,        /* let the row-length histogram decide */
  HISPMV_KERNEL_CSR_SCALAR = 1,  /* one thread per row: very short, regular rows */
  HISPMV_KERNEL_CSR_VECTOR = 2,  /* a sub-warp of 2..32 lanes per row: regular rows */
  HISPMV_KERNEL_MERGE = 3,       /* merge-path tiles, heavy rows split across CTAs, carry-out fix-up */
  HISPMV_KERNEL_GEMV = 4,        /* dense overlay: streaming row-major GeMV */
  HISPMV_KERNEL_EMPTY = 5,       /* nnz == 0: y = beta * bias */
  HISPMV_KERNEL_ADAPTIVE = 6,    /* row-aligned nnz-balanced tiles: short rows streamed through shared memory,
                                    long rows chunked across CTAs with carry-out (default for imbalanced rows) */
  HISPMV_KERNEL_ROWSTAGE = 7,    /* the same tiles with the col/val stream staged in shared memory by TMA bulk copies
                                    and rows walked by 1..32 lanes each: regular rows with column locality
                                    (banded / stencil / FEM) */
  HISPMV_KERNEL_BLOCKED = 8      /* two streaming passes for scattered columns over a large x: the nonzeros are kept a
                                    second time in column-slab-major order, pass 1 multiplies them with the slab's piece
                                    of x held in shared memory (the reference's column tiles and on-chip x buffers,
                                    common/src/spmv-helper.cpp:139-227), pass 2 sums the products per row panel */
};

/* ctor flags: the reference's hardware switches that still mean something on a GPU */
enum {
  HISPMV_FLAG_DENSE_OVERLAY = 1, /* dense_overlay: allow hispmv_add_dense (spmv-helper.cpp:718 asserts it) */
  HISPMV_FLAG_ROW_DIST_NET = 2   /* row_dist_net: allow heavy rows to be split across CTAs (merge kernel) */
};

#define HISPMV_HIST_BINS 33

typedef struct hispmv_matrix_info {
  int32_t rows, cols;       /* global shape as given by the caller */
  int32_t row_begin, row_end; /* the row block this context holds (whole matrix unless sharded) */
  int64_t nnz;              /* nonzeros held locally (dense: local_rows*cols) */
  int32_t is_dense;
  int32_t kernel;           /* HISPMV_KERNEL_* actually planned */
  int32_t vector_lanes;     /* lanes per row when kernel == CSR_VECTOR or ROWSTAGE */
  int32_t tile_items;       /* merge items (row ends + nonzeros) per CTA: MERGE tile size / ADAPTIVE stream budget */
  int64_t num_tiles;        /* merge tiles (CTAs) */
  int64_t num_split_rows;   /* rows whose nonzeros span more than one tile (the "shared rows") */
  int32_t max_row_nnz;
  int32_t empty_rows;
  int64_t hist[HISPMV_HIST_BINS]; /* hist[0]: empty rows; hist[k]: rows with 2^(k-1) <= nnz < 2^k */
  int64_t device_bytes;     /* HBM held by this matrix, incl. plan metadata */
  int64_t probe_near;       /* column-locality probe: sampled entry pairs within 32 columns of the row above ... */
  int64_t probe_cmp;        /* ... out of this many compared (selector input: banded when near/cmp >= 3/4) */
  int32_t x_window_cols;    /* > 0 only under the HISPMV_PERSIST research switch (persistent x-window kernel) */
  int32_t long_threshold;   /* ADAPTIVE / ROWSTAGE: rows with at least this many nonzeros become LONG tiles */
  int32_t chunk_nnz;        /* ADAPTIVE / ROWSTAGE: nonzeros per LONG tile (chunk of a heavy row) */
  int32_t num_slabs;        /* > 0: x is larger than L2 and the matrix runs as this many column slabs, one launch each */
  int32_t slab_cols;        /* columns per slab */
  int32_t reserved_;
  int64_t slab_runs;        /* selector input for BLOCKED: (row, 49152-column slab) runs of the matrix; 0 when the cheaper
                               conditions (size, scattered columns) already ruled the strategy out */
} hispmv_matrix_info;

const char* hispmv_last_error(void);
int hispmv_version(void);   /* 200 = this ABI; +1 when built with the research kernels (make EXPERIMENTAL=1) */

/* device_id: CUDA ordinal.  flags: HISPMV_FLAG_*.  Fails (no fallback) if the device is not sm_100. */
int hispmv_create(hispmv_ctx** out, int device_id, int flags);
/* One handle, n_gpus devices (first_device .. first_device + n_gpus - 1) in ONE process -- the single-handle model of
 * the reference's FpgaHandle (pyhispmv/src/fpga_handle.cpp:40-154) over several GPUs: every matrix added from host
 * arrays is cut into n_gpus nnz-balanced row blocks (the split points of hispmv_set_shard), one per GPU; hispmv_run /
 * hispmv_linear send the whole x and each GPU's block of bias to its GPU, run the blocks concurrently (one host thread
 * per GPU) and write the blocks of y side by side into the caller's vector.  The host-buffer surface (add_*, load_mtx,
 * commit, select, run, linear, matrix_info, force_kernel, sync) works on such a handle; the device-pointer and plan
 * calls answer HISPMV_ERR_STATE -- reach the per-GPU contexts with hispmv_multi_child.  The one-process-per-GPU model
 * (hispmv_set_shard + a collective library outside this ABI) remains the one for device-resident pipelines. */
int hispmv_create_multi(hispmv_ctx** out, int first_device, int n_gpus, int flags);
int hispmv_multi_gpus(hispmv_ctx* ctx);                    /* 0 for a plain handle */
hispmv_ctx* hispmv_multi_child(hispmv_ctx* ctx, int k);    /* owned by ctx; NULL when out of range */
void hispmv_destroy(hispmv_ctx* ctx);

/* Row-block sharding for one-process-per-GPU runs: after this call every matrix added keeps only part
 * `part` of `n_parts` -- contiguous, nnz-balanced row blocks for sparse matrices (split points =
 * lower_bound(row_ptr, k*nnz/n_parts); rows are never split across GPUs), equal row blocks for dense.
 * y and bias passed to run/linear then have row_end-row_begin entries; x stays full length. */
int hispmv_set_shard(hispmv_ctx* ctx, int part, int n_parts);
/* The split points themselves (n_parts+1 entries, bounds[0]=0, bounds[n_parts]=rows) from a HOST row_ptr. */
int hispmv_shard_bounds(const int32_t* row_ptr, int32_t rows, int n_parts, int32_t* bounds);
/* Optional cap on the HBM the context may hold (bytes; 0 = device capacity).  The reference's cap is
 * 256 MiB per HBM channel (fpga_handle.h:12). */
int hispmv_set_memory_limit(hispmv_ctx* ctx, int64_t bytes);

/* ---- adding matrices: return handle index >= 0, HISPMV_FULL (-1), or another negative status ----
 * COO: unsorted, duplicates kept; an index outside [0,rows) x [0,cols) is refused (HISPMV_ERR_ARG).
 * CSR: row_ptr[0] = 0, non-decreasing, row_ptr[rows] = nnz, and 0 <= col < cols -- refused otherwise (HISPMV_ERR_ARG);
 *      a row whose columns are not in non-decreasing order is accepted and re-sorted by (column, value), the order the
 *      COO path produces (the reference's per-row std::sort, common/src/spmv-helper.cpp:216). */
int hispmv_add_sparse_coo(hispmv_ctx* ctx, const int32_t* coo_rows, const int32_t* coo_cols, const float* coo_vals,
                          int64_t nnz, int32_t rows, int32_t cols);
int hispmv_add_sparse_csr(hispmv_ctx* ctx, const int32_t* row_ptr, const int32_t* col_idx, const float* vals,
                          int32_t rows, int32_t cols);
int hispmv_add_dense(hispmv_ctx* ctx, const float* a_rowmajor, int32_t rows, int32_t cols);
/* device-resident inputs (benchmarks and generators that must not cross PCIe); arrays are copied. */
int hispmv_add_sparse_coo_dev(hispmv_ctx* ctx, const int32_t* d_rows, const int32_t* d_cols, const float* d_vals,
                              int64_t nnz, int32_t rows, int32_t cols);
int hispmv_add_sparse_csr_dev(hispmv_ctx* ctx, const int32_t* d_row_ptr, const int32_t* d_col_idx,
                              const float* d_vals, int32_t rows, int32_t cols);
int hispmv_add_dense_dev(hispmv_ctx* ctx, const float* d_a_rowmajor, int32_t rows, int32_t cols);

int hispmv_commit(hispmv_ctx* ctx);
int hispmv_num_matrices(hispmv_ctx* ctx);
int hispmv_select(hispmv_ctx* ctx, uint32_t idx);

/* Force a strategy for matrix idx (HISPMV_KERNEL_*; lanes only for CSR_VECTOR, 0 = auto) and re-plan.
 * Used by the selector tests and the per-kernel benchmarks. */
int hispmv_force_kernel(hispmv_ctx* ctx, int idx, int kernel, int lanes);

/* ---- host-buffer calls (synchronous; copies inside, like the reference's BO syncs) ---- */
int hispmv_run(hispmv_ctx* ctx, const float* x, const float* bias, float* y, float alpha, float beta);
int hispmv_linear(hispmv_ctx* ctx, int idx, const float* x, int64_t x_len, const float* bias, float* y_out);
/* hispmv_run with x already in HBM (d_x is complete once the work queued on x_stream -- a cudaStream_t -- so far has
 * run); bias and y are host buffers, copied and pipelined as in hispmv_run.  The multi-GPU host path uses it: each
 * rank sends only its slice of x across PCIe and the slices meet over NVLink (hispmv_multicast_copy). */
int hispmv_run_xdev(hispmv_ctx* ctx, const float* d_x, void* x_stream, const float* bias, float* y, float alpha,
                    float beta);

/* ---- device-buffer calls: asynchronous on `stream`, a cudaStream_t.  As everywhere in CUDA, NULL is the
 *      default stream; hispmv_stream() returns the context's own non-blocking stream. ---- */
void* hispmv_stream(hispmv_ctx* ctx);
int hispmv_run_dev(hispmv_ctx* ctx, int idx, const float* d_x, const float* d_bias, float* d_y, float alpha,
                   float beta, void* stream);
/* hispmv_run_dev one phase at a time, for callers that overlap the phases with their own copies or exchanges (the
 * reference overlaps host-side fills with the running kernel, pyhispmv/src/fpga_handle.cpp:366-379).  phases: 1 = the
 * products (needs only x; BLOCKED matrices, a no-op for the one-pass strategies), 2 = row sums + alpha/beta epilogue
 * (needs bias; the whole operation for the one-pass strategies), 3 = both = hispmv_run_dev. */
int hispmv_run_dev_phase(hispmv_ctx* ctx, int idx, const float* d_x, const float* d_bias, float* d_y, float alpha,
                         float beta, int phases, void* stream);
/* y = relu?(A x + bias) for chained layers that stay on the device. */
int hispmv_linear_dev(hispmv_ctx* ctx, int idx, const float* d_x, const float* d_bias, float* d_y, int relu,
                      void* stream);

/* The all-gather of a chained layer fused into the kernel that produces y: mc_y is the NVSwitch MULTICAST address of
 * this rank's first row inside a vector replicated on every GPU of a multicast group (symmetric-memory rendezvous), and
 * every result is stored with multimem.st, so the switch delivers this rank's y block to all replicas -- the next
 * layer's x -- with no separate collective (reference: the host-side copy of y into the next layer's input,
 * apps/fpga_layer_manager.py:58-67).  y = relu?(alpha*A x + beta*bias).  Peers need a barrier on the group before
 * they read.  Strategies that read y back (MERGE, column slabs) refuse with HISPMV_ERR_STATE. */
int hispmv_run_dev_mc(hispmv_ctx* ctx, int idx, const float* d_x, const float* d_bias, float* mc_y, float alpha,
                      float beta, int relu, void* stream);

/* Several right-hand sides in one pass ("SpMM-lite"): d_x is [num_vecs][cols], d_y [num_vecs][local rows], both
 * row-major in HBM; y_k = relu?(alpha*A x_k + beta*bias).  The reference runs a batch vector by vector (runLinear's
 * loop, pyhispmv/src/fpga_handle.cpp:336,366-379); here up to eight vectors are interleaved so that col/val are read
 * once and one 32-byte sector gather serves all of them.  Dense matrices and sparse matrices whose longest row has at
 * most 65536 nonzeros take this path, everything else falls back to one launch per vector (same results contract).
 * hispmv_linear uses it for num_vecs >= 2. */
int hispmv_run_dev_batch(hispmv_ctx* ctx, int idx, const float* d_x, const float* d_bias, float* d_y, int64_t num_vecs,
                         float alpha, float beta, int relu, void* stream);

int hispmv_sync(hispmv_ctx* ctx);
/* number of kernels one hispmv_run_dev of matrix idx launches (for bench.py's gpu_launches) */
int hispmv_launches_per_run(hispmv_ctx* ctx, int idx);

/* ---- plan introspection: the bit-exact integer contract ---- */
int hispmv_matrix_info_get(hispmv_ctx* ctx, int idx, hispmv_matrix_info* out);
/* Copy the device CSR of the local row block back to the host (row_ptr rebased to 0).  Any pointer may be
 * NULL to skip that array.  Sizes: local_rows+1, nnz, nnz. */
int hispmv_plan_csr(hispmv_ctx* ctx, int idx, int32_t* row_ptr, int32_t* col_idx, float* vals);
/* Tile start coordinates, num_tiles+1 entries each (last = (local_rows, nnz)).
 * MERGE: merge-path coordinates.  ADAPTIVE: first row of the tile and the offset of its first nonzero. */
int hispmv_plan_tiles(hispmv_ctx* ctx, int idx, int32_t* tile_row, int64_t* tile_nnz);
/* Sorted ids of the rows split across tiles (num_split_rows entries). */
int hispmv_plan_split_rows(hispmv_ctx* ctx, int idx, int32_t* rows_out);
/* Column slab `slab` of matrix idx as CSR over the local rows (sizes: local_rows+1, and the slab's nnz, which
 * hispmv_plan_slab_nnz returns).  Any pointer may be NULL. */
int64_t hispmv_plan_slab_nnz(hispmv_ctx* ctx, int idx, int slab);
int hispmv_plan_slab_csr(hispmv_ctx* ctx, int idx, int slab, int32_t* row_ptr, int32_t* col_idx, float* vals);
/* ADAPTIVE / BLOCKED: per tile (panel), -1 for a STREAM tile or the chunk index of a LONG tile (num_tiles entries). */
int hispmv_plan_tile_chunks(hispmv_ctx* ctx, int idx, int32_t* chunk_out);
/* BLOCKED only (the column tiling of tileAndPad, common/src/spmv-helper.cpp:139-227, as this engine lays it out).
 * out8 = { slab_cols, num_slabs, padded_nnz, num_segments, num_chunks, pass-1 work ranges, panels, num_pieces }.
 * hispmv_plan_blocked copies the plan to the host (any pointer may be NULL):
 *   slab_ptr[num_slabs+1]      slab starts in the slab-major order (multiples of 512)
 *   vals / lcol [padded_nnz]   value and column - slab*slab_cols of every entry (padding entries are zero); inside a
 *                              512-entry group entry e = 16*lane + w is stored at ((w/4)*32 + lane)*4 + w%4 (vals) and
 *                              ((w/8)*32 + lane)*8 + w%8 (lcol), the order a warp's vector loads consume
 *   flags[padded_nnz/16]       bit j of word i: entry 16i+j is the last of its PIECE (consecutive entries of one row
 *                              inside one 512-entry group); pieces are numbered in slab-major order
 *   group_base[padded_nnz/512+1]  pieces that end before each group
 *   prow_ptr[local_rows+1]     CSR-style offsets of every row's pieces (slab order inside a row); the panels
 *                              (hispmv_plan_tiles / hispmv_plan_tile_chunks) are cut over these, not over nonzeros
 *   perm[num_pieces]           a piece's position in that per-row order minus the first position of its panel
 *   panel_seg[panels+1], seg_start_off[2*num_segments]   per non-empty (panel, slab) segment, panel-major:
 *                              (first piece id, pieces of the same panel in earlier slabs)
 *   panel_chunk[panels+1], chunk_start_count[2*num_chunks]   the same segments cut into runs of at most 32 pieces,
 *                              (first piece id, count): what a warp of pass 2 fetches per step
 *   work[2*ranges]             pass-1 [begin, end) per resident CTA
 *   (hispmv_plan_blocked_stage) how pass 2 fetches a STREAM panel: the 16-byte-aligned quads of partial sums that
 *   cover its segments, listed panel by panel:
 *   out4 = { positions (4 per quad) in total, end-mark words in total, shared-memory words of the largest panel, segments }
 *   seg_copy[2*num_segments]   (first piece id of the aligned range, panel-relative first quad | quads << 16); no quads
 *                              for segments of LONG panels (those are summed through the chunk table)
 *   chunk_src[positions/4]     the first piece id of every quad
 *   perm2[positions]           the slot (perm) of every position, 0xFFFF = alignment padding
 *   panel_aux[2*(panels+1)]    (first position, first end-mark word) of every panel
 *   end_bits[words]            per STREAM panel, bit j of its words: slot j ends a row */
int hispmv_plan_blocked_info(hispmv_ctx* ctx, int idx, int64_t* out8);
int hispmv_plan_blocked(hispmv_ctx* ctx, int idx, int32_t* slab_ptr, float* vals, uint16_t* lcol, uint16_t* flags,
                        int32_t* group_base, int32_t* prow_ptr, uint16_t* perm, int32_t* panel_seg,
                        int32_t* seg_start_off, int32_t* panel_chunk, int32_t* chunk_start_count, int32_t* work);
int hispmv_plan_blocked_stage(hispmv_ctx* ctx, int idx, int64_t* out4, int32_t* seg_copy, uint16_t* perm2,
                              int32_t* panel_aux, uint32_t* end_bits, int32_t* chunk_src);

/* ---- x exchange over NVSwitch multicast: store n floats from d_src to a multicast address (every GPU of the
 *      multicast group receives them).  mc_dst comes from a symmetric-memory rendezvous; sm_budget > 0 = that many
 *      CTAs of multimem.st stores, 0 = 32 CTAs, < 0 = a copy engine writes to the multicast address (no SM used).
 *      Asynchronous on `stream`; peers need a barrier on the same group before they read. ---- */
int hispmv_multicast_copy(void* mc_dst, const float* d_src, int64_t n, int sm_budget, void* stream);
/* The same n floats stored into n_peers peer buffers (device pointers of this or other GPUs, peer-mapped: the replicas
 * of a symmetric-memory rendezvous) with plain stores over NVLink, ctas_per_peer CTAs each (0 = 2).  The all-gather of
 * a distributed x: every rank sends its block to every replica, its own included. */
int hispmv_peer_copy(void* const* peer_dst, int n_peers, const float* d_src, int64_t n, int ctas_per_peer, void* stream);

/* ---- Matrix Market ingest: real/integer/pattern x general/symmetric/skew-symmetric ---- */
int hispmv_load_mtx(hispmv_ctx* ctx, const char* path);
/* The parser behind hispmv_load_mtx on its own (host only, no GPU needed): the entry lines are parsed by one thread per
 * ~1 MB chunk and come out as 0-based COO in file order, exactly what loadMtx (common/src/spmv-helper.cpp:34-136) and
 * its two twins (gpu/src/spmvHelper.cpp:4-115, cpu/src/helper_functions.cpp:91-210) return.  Arrays are malloc'ed by the
 * callee; release them with hispmv_parse_mtx_free. */
int hispmv_parse_mtx(const char* path, int32_t* rows, int32_t* cols, int64_t* nnz, int32_t** coo_rows,
                     int32_t** coo_cols, float** coo_vals);
void hispmv_parse_mtx_free(int32_t* coo_rows, int32_t* coo_cols, float* coo_vals);

#ifdef __cplusplus
}
#endif
#endif /* HISPMV_H_ */
