// Internal declarations shared by the translation units of libhispmv_cuda.so.
// Nothing here is part of the C-ABI (see include/hispmv.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "hispmv.h"

namespace hispmv {

void set_error(const std::string& msg);
// returns HISPMV_OK or records the CUDA error text and returns HISPMV_ERR_CUDA / HISPMV_FULL (OOM)
int check_cuda(cudaError_t e, const char* what, const char* file, int line);
#define HISPMV_CUDA(expr)                                                    \
  do {                                                                       \
    int _st = ::hispmv::check_cuda((expr), #expr, __FILE__, __LINE__);       \
    if (_st != HISPMV_OK) return _st;                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Device-side matrix as the kernels see it.
// ---------------------------------------------------------------------------------------------
struct CsrDev {
  int32_t rows = 0;       // local rows (row block held by this GPU)
  int32_t cols = 0;
  int64_t nnz = 0;
  const int32_t* row_ptr = nullptr;  // rows+1, rebased to 0
  const int32_t* col = nullptr;      // nnz, padded to a multiple of 4 plus 4 (zeros)
  const float* val = nullptr;        // nnz, same padding
};

struct MergePlan {
  int32_t tile_items = 0;        // merge items per CTA (THREADS * ITEMS_PER_THREAD of the instantiation used)
  int64_t num_tiles = 0;
  const int32_t* tile_row = nullptr;   // num_tiles+1
  const int64_t* tile_nnz = nullptr;   // num_tiles+1
  float* carry = nullptr;              // num_tiles: partial sum of the row left open at each tile's end
};

// Row-aligned adaptive tiling (the default for imbalanced matrices):
//   STREAM tile  consecutive rows, each shorter than long_threshold, together about stream_items merge items
//                (row ends + nonzeros; never more than stream_items + long_threshold)
//   LONG tile    one chunk of at most chunk_nnz nonzeros of a row with >= long_threshold nonzeros; a row
//                with several chunks is "split": its partial sums meet in carry[] (one slot per tile)
// One tile of an AdaptivePlan, 32 bytes (two 128-bit loads).
struct TileDesc {
  int32_t r0, r1;      // rows [r0, r1) (LONG: the one row r0)
  int32_t n0, n1;      // nonzeros [n0, n1) (LONG: the chunk)
  int32_t chunk;       // -1 for a STREAM tile, else the chunk index of a LONG tile
  int32_t nchunks;     // LONG: number of chunks of the row
  int32_t tile;        // the tile's own index
  int32_t pad;
};

struct AdaptivePlan {
  int32_t stream_items = 0, long_threshold = 0, chunk_nnz = 0;
  int64_t num_tiles = 0;
  const int32_t* tile_row = nullptr;    // num_tiles+1: first row of every tile, then `rows`
  const int32_t* tile_chunk = nullptr;  // num_tiles: -1 for STREAM tiles, chunk index for LONG tiles
  const TileDesc* desc = nullptr;       // num_tiles: the same facts resolved against row_ptr, one record per tile
  float* carry = nullptr;               // num_tiles partial sums (LONG tiles of split rows)
  unsigned int* counter = nullptr;      // num_tiles arrival counters, indexed by a split row's first tile
  int32_t hot_cols = 0x7fffffff;        // x[c] with c < hot_cols is kept in L1 (persistent kernel: in shared
                                        // memory), the rest bypasses L1 allocation
  unsigned int* sched = nullptr;        // persistent kernel: [0] next tile, [1] groups that ran dry
  int64_t tile_begin = 0;               // one-CTA-per-tile kernels: launch only tiles [tile_begin, tile_begin + tile_count)
  int64_t tile_count = -1;              // (-1: all) -- the host-buffer call pipelines row ranges against PCIe copies
  int32_t ahead = 0;                    // > 0: every CTA asks L2 to fetch the col/val range of tile t + ahead (the tile
                                        // the next wave runs in its place), so DRAM keeps streaming while this wave
                                        // gathers and reduces; 0 = off; < 0: every CTA prefetches its own tile
  int32_t ahead_all = 0;                // 0: look ahead only for LONG tiles (chunks of long rows: four dependent
                                        // load -> gather rounds per thread, the part that is DRAM-latency-bound)
};

struct RowStats {
  int64_t hist[HISPMV_HIST_BINS];
  int64_t nnz;
  int32_t rows;
  int32_t max_row_nnz;
  int32_t empty_rows;
};

// ---- partition.cu ---------------------------------------------------------------------------
// COO (device) -> CSR (device).  Entries are ordered by (row, col, value) exactly like the reference's
// per-row std::sort over (col, val) pairs (common/src/spmv-helper.cpp:216; gpu/src/spmvHelper.cpp:139).
// Outputs are cudaMalloc'ed by the callee; col/val carry the 128-bit padding described in CsrDev.
int coo_to_csr_device(const int32_t* d_rows, const int32_t* d_cols, const float* d_vals, int64_t nnz, int32_t rows,
                      int32_t cols, int32_t** d_row_ptr, int32_t** d_col, float** d_val, cudaStream_t stream);
// Row-length histogram (power-of-two bins) + max / empty counts, computed on the device.
int row_stats_device(const int32_t* d_row_ptr, int32_t rows, RowStats* out, cudaStream_t stream);
// Merge-path tile start coordinates for `tile_items` items per tile; arrays cudaMalloc'ed by the callee.
int merge_tiles_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int32_t tile_items, int64_t* num_tiles,
                       int32_t** d_tile_row, int64_t** d_tile_nnz, cudaStream_t stream);
// Rows whose nonzeros span more than one tile, ascending.  d_out cudaMalloc'ed by the callee (may be null if 0).
int split_rows_device(const int32_t* d_row_ptr, int32_t rows, const int32_t* d_tile_row, const int64_t* d_tile_nnz,
                      int64_t num_tiles, int32_t** d_out, int64_t* count, cudaStream_t stream);
// Slice [row_begin,row_end) out of a device CSR into fresh, padded arrays with row_ptr rebased to 0.
int csr_slice_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t row_begin,
                     int32_t row_end, int32_t** o_row_ptr, int32_t** o_col, float** o_val, int64_t* o_nnz,
                     cudaStream_t stream);
// Pad-copy col/val (device->device or host->device, given by `kind`) into freshly allocated padded arrays.
int alloc_padded_nnz_arrays(const int32_t* src_col, const float* src_val, int64_t nnz, cudaMemcpyKind kind,
                            int32_t** d_col, float** d_val, cudaStream_t stream);
// nnz-balanced split points from a device row_ptr: bounds[k] = lower_bound(row_ptr, k*nnz/n_parts).
int shard_bounds_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int n_parts, int32_t* h_bounds,
                        cudaStream_t stream);

// Caller-supplied CSR: h_flags3 = { row_ptr not monotone within [0, nnz], a column outside [0, cols), a row whose columns
// are not non-decreasing }.  csr_expand_rows_device writes the row index of every entry (for the re-sort).
int csr_validate_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, int32_t cols, int64_t nnz,
                        int* h_flags3, cudaStream_t stream);
int csr_expand_rows_device(const int32_t* d_row_ptr, int32_t rows, int64_t nnz, int32_t* d_rows_out,
                           cudaStream_t stream);

// Adaptive tiles; arrays cudaMalloc'ed by the callee.  d_split_rows: rows with more than one LONG chunk.
int adaptive_tiles_device(const int32_t* d_row_ptr, int32_t rows, int32_t stream_items, int32_t long_threshold,
                          int32_t chunk_nnz, int64_t* num_tiles, int32_t** d_tile_row, int32_t** d_tile_chunk,
                          int32_t** d_split_rows, int64_t* num_split, cudaStream_t stream);

// Column-locality probe: for up to kProbeSamples evenly spaced rows r, the first min(len(r), len(r-1), 32) entries of
// rows r and r-1 are compared position by position; `near` counts pairs whose columns differ by at most 32 (same or
// adjacent 128-byte line of x), `cmp` the pairs compared.  Banded / stencil / FEM matrices score near 1.
struct ColProbe {
  int64_t near = 0, cmp = 0;
};
constexpr int32_t kProbeSamples = 8192;
int col_probe_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, ColProbe* out, cudaStream_t stream);

// Column slab [lo_col, hi_col) of a device CSR as a CSR over the same rows (arrays cudaMalloc'ed by the callee, padded).
int csr_column_slab_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t rows,
                           int32_t lo_col, int32_t hi_col, int32_t** o_row_ptr, int32_t** o_col, float** o_val,
                           int64_t* o_nnz, cudaStream_t stream);
constexpr int64_t kSlabMinBytes = 64ll << 20;   // x up to 64 MB is left whole
constexpr int32_t kSlabMaxCols = 12500000;      // 50 MB of x per slab (C5 sweep: 10 M 8.1 ms, 12.5 M 7.5, 20 M 7.9, 25 M 9.0)
// 0 = no slabs, else the slab width in columns (restated in oracle/: oracle_select_slab_cols)
int32_t select_slab_cols(int32_t cols, int64_t nnz, const ColProbe& probe);
// carry[] slots of split rows hold this NaN payload until their chunk has written its partial (adaptive.cu)
constexpr uint32_t kCarryEmptyBits = 0x7fc0dead;
cudaError_t fill_u32_device(uint32_t* p, uint32_t value, size_t n, cudaStream_t stream);
// TileDesc records from tile_row / tile_chunk / row_ptr (one thread per tile); d_desc cudaMalloc'ed by the callee.
int tile_desc_device(const int32_t* d_row_ptr, const int32_t* d_tile_row, const int32_t* d_tile_chunk,
                     int64_t num_tiles, int32_t chunk_nnz, TileDesc** d_desc, cudaStream_t stream);

// The runtime selector (pure host integer arithmetic over RowStats; restated in oracle/).
void select_kernel(const RowStats& st, const ColProbe& probe, int allow_split_rows, int* kernel, int* lanes);
// tile_items the merge kernel instantiation uses for a matrix with these stats
int merge_tile_items_for(const RowStats& st);
// ROWSTAGE plan parameters from the row statistics (lanes_in = 0: choose lanes); restated in oracle/.
void rowstage_params(const RowStats& st, int lanes_in, int* lanes, int32_t* stream_items, int32_t* long_threshold,
                     int32_t* chunk_nnz);

// ---- spmv.cu --------------------------------------------------------------------------------
struct Epilogue {
  float alpha, beta;
  const float* bias;  // may be null when beta == 0
  int relu;           // fused max(.,0) (device-resident chained layers only)
  int y_mc = 0;       // y is a multicast address: results are stored with multimem.st and land on every GPU of the group
};
int launch_csr_scalar(const CsrDev& A, const float* x, float* y, Epilogue ep, cudaStream_t s);
int launch_csr_vector(const CsrDev& A, int lanes, const float* x, float* y, Epilogue ep, cudaStream_t s);
int launch_merge(const CsrDev& A, const MergePlan& P, const float* x, float* y, Epilogue ep, cudaStream_t s);
int launch_adaptive(const CsrDev& A, const AdaptivePlan& P, int threads, const float* x, float* y, Epilogue ep,
                    cudaStream_t s);
constexpr int kAdaptiveStreamItems = 2048, kAdaptiveLongThreshold = 1024, kAdaptiveChunkNnz = 4096;
// Persistent nnz-major kernel: one CTA per SM, x[0, hot_cols) held in shared memory, tiles pulled from P.sched.
int launch_adaptive_persistent(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep,
                               int sm_count, cudaStream_t s);
constexpr int kPersistentMaxHot = 40960;  // floats of x the window may hold next to the four group buffers
// experimental.cu: research kernels, compiled only with -DHISPMV_EXPERIMENTAL (the default build has refusing stubs)
bool experimental_kernels_built();
// One warp per tile (tiles of a few hundred items): stream_items + long_threshold <= kWarpTileCap.
int launch_warptile(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, cudaStream_t s);
constexpr int kWarpTileCap = 512;
// Warp-specialised persistent pipeline (TMA producer / gather teams / reduce warps).  A stage holds one tile:
// stream_items + long_threshold <= kPipelineCap, chunk_nnz <= kPipelineCap, stream_items <= kPipelineRows.
constexpr int kPipelineCap = 1920, kPipelineRows = 1408;  // CAP = 6 gathers x 320 team threads
int launch_pipeline(const CsrDev& A, const AdaptivePlan& P, const float* x, float* y, Epilogue ep, int sm_count,
                    cudaStream_t s);
// Row-major STREAM tiles behind a TMA-staged stream (regular rows with column locality); `lanes` lanes per row.
int launch_rowstage(const CsrDev& A, const AdaptivePlan& P, int lanes, int threads, const float* x, float* y,
                    Epilogue ep, cudaStream_t s);
constexpr int kRowstageMaxCap = 8192;
int launch_empty(int32_t rows, float* y, Epilogue ep, cudaStream_t s);
bool merge_tile_items_supported(int tile_items);

// ---- blocked.cu: column-blocked two-pass SpMV (x gathers served from shared memory) --------------------------------
// For matrices whose x gathers have no locality (power-law / uniform columns, x far larger than L1) a scattered 4-byte
// gather costs a whole 32-byte L2 sector, and the L2's sector rate -- not HBM -- bounds the one-pass kernels (DESIGN.md
// section 4; cuSPARSE sits on the same ceiling).  The blocked strategy takes the gathers out of L2 altogether:
//   expand  (pass 1)  the nonzeros are stored a second time in SLAB-MAJOR order (slab = slab_cols consecutive columns,
//                     inside a slab in CSR order, every slab padded to a multiple of kPbGroup entries); a CTA keeps the
//                     slab's piece of x in shared memory (TMA bulk loads) and streams val / 16-bit local column.
//                     Consecutive entries of one row inside one 128-entry group form a PIECE; a warp multiplies a group,
//                     adds up every piece with a segmented shuffle scan and writes one partial sum per piece --
//                     a long row costs one partial per slab and group, a hypersparse row one per nonzero.
//   reduce  (pass 2)  rows are cut into PANELS (the adaptive tiles over the per-row piece counts: STREAM panels of rows
//                     with few pieces, LONG panels = chunks of a row with very many); a panel's partials are the
//                     (panel, slab) SEGMENTS -- runs of consecutive piece ids -- which a CTA walks in slab order,
//                     dropping each partial at its place in the panel's per-row order (16-bit perm) in shared memory,
//                     then sums the rows in that order (deterministic) and applies alpha / beta / ReLU.
// Traffic: 4 + 2 + 0.25 bytes per nonzero plus 4 per piece in pass 1, 4 + 2 per piece in pass 2 -- all of it streaming.
struct PbSeg {
  int32_t start;  // first piece id of the segment (piece ids follow the slab-major order)
  int32_t off;    // number of pieces of the same panel in earlier slabs (the segment's offset in the panel's walk)
};
struct PbPlan {
  int32_t slab_cols = 0, num_slabs = 0;
  int64_t padded_nnz = 0;              // length of the blocked arrays: every slab starts at a multiple of kPbGroup
  const int32_t* slab_ptr = nullptr;   // num_slabs+1 starts in blocked order (device)
  const float* val = nullptr;          // blocked order; padding entries are 0
  const uint16_t* lcol = nullptr;      // column - slab * slab_cols
  const uint16_t* flags = nullptr;     // padded_nnz/16: bit j of word i = entry 16i+j ends a piece
  const int32_t* group_base = nullptr; // padded_nnz/kPbGroup + 1: pieces that end before each group
  const uint16_t* perm = nullptr;      // per piece: position in its row's piece order - the panel's first position
  const int32_t* prow_ptr = nullptr;   // rows+1: CSR-style offsets of the pieces of every row
  float* part = nullptr;               // one partial sum per piece: pass 1 output / pass 2 input
  int64_t num_pieces = 0;
  int64_t num_panels = 0;
  const TileDesc* desc = nullptr;      // the panels (adaptive tiles over prow_ptr)
  const int32_t* panel_seg = nullptr;  // num_panels+1 offsets into seg[]
  const PbSeg* seg = nullptr;          // non-empty (panel, slab) segments, panel-major then slab
  int32_t max_panel_segs = 0;         // the most runs (chunk[] entries) any panel has: sizes pass 2's shared-memory copy
  const int32_t* panel_chunk = nullptr;  // num_panels+1 offsets into chunk[]
  const int2* chunk = nullptr;           // the segments cut into runs of at most kPbChunk pieces: (first piece id, count)
  // the gather of pass 2 (STREAM panels): the 16-byte-aligned quads of partial sums that cover the panel's segments,
  // panel-major
  const int32_t* chunk_src = nullptr;  // per quad: its first piece id (a multiple of 4)
  const int2* seg_copy = nullptr;      // per segment: {first piece of the aligned range, panel-relative quad | quads << 16}
  const uint16_t* perm2 = nullptr;     // per position of those quads: slot in the panel, 0xFFFF = alignment padding
  const int2* panel_aux = nullptr;     // num_panels+1: {first staged position, first word of end_bits}
  const uint32_t* end_bits = nullptr;  // per STREAM panel, one bit per slot: this slot ends a row
  int32_t reduce_words = 0;            // shared-memory words of the largest STREAM panel (skewed slots)
  const int2* work = nullptr;          // pass 1: [k0, k1) in blocked order per CTA, cost-balanced
  int32_t work_begin = 0;              // first entry of work[] this launch uses (0: whole order; num_work / 2 num_work:
                                       // the head / tail part of the two-part host-buffer schedule)
  long long* dbg = nullptr;            // development (HISPMV_PB_DEBUG): per pass-1 CTA {ns busy, slab loads, groups}
  int32_t num_work = 0;
  int32_t cap_words = 0;               // upper bound on the slots of a STREAM panel (panel items + long threshold)
  int64_t panel_begin = 0, panel_count = -1;  // pass 2: launch only these panels (host-buffer pipeline)
  float* carry = nullptr;              // split LONG rows: as in AdaptivePlan
  unsigned int* counter = nullptr;
};
constexpr int32_t kPbChunk = 32;           // pieces a warp of pass 2 fetches per step
constexpr int32_t kPbGroup = 512;          // entries a warp handles per step: 16 consecutive ones per lane
constexpr int32_t kPbMaxSlabCols = 49152;  // 192 KB of x next to the 32 KB in which pass 1's sixteen warps stage their pieces
// owned device arrays of a blocked plan (built by pb_order_device + pb_segments_device, freed by pb_free)
struct PbArrays {
  int32_t slab_cols = 0, num_slabs = 0;
  int64_t padded_nnz = 0, num_pieces = 0, num_seg = 0;
  int32_t max_panel_segs = 0;
  int32_t* d_slab_ptr = nullptr;
  float* d_val = nullptr;
  uint16_t* d_lcol = nullptr;
  uint16_t* d_flags = nullptr;
  int32_t* d_group_base = nullptr;
  int32_t* d_prow_ptr = nullptr;
  uint16_t* d_perm = nullptr;
  int32_t* d_panel_seg = nullptr;
  PbSeg* d_seg = nullptr;
  int32_t* d_panel_chunk = nullptr;
  int2* d_chunk = nullptr;
  int64_t num_chunks = 0;
  int2* d_seg_copy = nullptr;
  uint16_t* d_perm2 = nullptr;
  int32_t* d_chunk_src = nullptr;
  int2* d_panel_aux = nullptr;
  uint32_t* d_end_bits = nullptr;
  int64_t stage_total = 0, bit_words = 0;
  int32_t reduce_words = 0;
  int2* d_work = nullptr;              // 3 * num_work ranges: whole order, head part, tail part (pb_make_work)
  int32_t num_work = 0;
  int64_t head_cols = 0;               // columns of x the head part needs (0: no two-part schedule)
  float* d_part[2] = {nullptr, nullptr};  // one per stream lane, the second allocated on first use
  int32_t* h_slab_ptr = nullptr;          // host copy (num_slabs+1), new[]
  int32_t* d_piece_pcsr = nullptr;        // between the two build stages only
  int32_t* d_piece_slab = nullptr;
};
void pb_free(PbArrays* a);
// Stage 1: the blocked copy of a device CSR, its pieces and prow_ptr.  Stage 2 (after the panels have been cut over
// prow_ptr with adaptive_tiles_device / tile_desc_device): perm and the segment table.  Restated in oracle/
// (oracle_pb_order, oracle_pb_segments).
int pb_order_device(const int32_t* d_row_ptr, const int32_t* d_col, const float* d_val, int32_t rows, int32_t cols,
                    int64_t nnz, int32_t slab_cols, PbArrays* out, cudaStream_t stream);
int pb_segments_device(PbArrays* a, const TileDesc* d_desc, int64_t num_panels, int32_t rows, cudaStream_t stream);
// pass-1 work ranges for `n_cta` resident CTAs: contiguous, balanced by entries + piece_cost16/16 per piece + slab_cost
// per slab (re)load
int pb_make_work(PbArrays* a, int n_cta, int64_t slab_cost, int64_t piece_cost16, cudaStream_t stream);
int launch_pb_expand(const PbPlan& P, int32_t cols, const float* x, cudaStream_t s);
int launch_pb_reduce(const CsrDev& A, const PbPlan& P, float* y, Epilogue ep, cudaStream_t s);
// Slab width / panel parameters of the blocked strategy, and whether the selector prefers it (restated in oracle/).
constexpr int32_t kPbSlabCols = 49152, kPbPanelItems = 7168, kPbLongThreshold = 1024, kPbChunkNnz = 8192;
// pass-1 balance, fitted to the CTAs' busy times on C2 (HISPMV_PB_DEBUG): a piece is worth 6/16 of an entry (97 ns per
// group, 0.073 ns per piece), staging a slab's x slice 15 000 entries (2.9 us)
constexpr int64_t kPbPieceCost16 = 6, kPbSlabCost = 15000;
// (row, slab) runs of a device CSR for slabs of slab_cols columns: the selector's estimate of the piece count
int pb_count_runs_device(const int32_t* d_row_ptr, const int32_t* d_col, int32_t rows, int32_t slab_cols, int64_t* runs,
                         cudaStream_t stream);
int select_blocked(int32_t rows, int32_t cols, int64_t nnz, int64_t slab_runs, const ColProbe& probe,
                   int allow_split_rows);

// ---- batch.cu: several right-hand sides in one pass over A (x interleaved as xi[c * K + k], K = batch_width(nv)) ----
int batch_width(int nv);  // 2, 4 or 8
// x [nv][n] -> xi [n_pad][K] (rows n..n_pad-1 and vectors nv..K-1 zero)
int launch_interleave(const float* x, int nv, int64_t n, int64_t n_pad, float* xi, cudaStream_t s);
// y [nv][rows] = alpha * A x_k + beta * bias (+ReLU); `lanes` lanes walk each row
// rows longer than long_len (the list long_rows, built once per matrix by batch_long_rows_device) get one CTA each
int launch_spmm_csr(const CsrDev& A, int lanes, const int32_t* long_rows, int64_t n_long, int long_len, const float* xi,
                    float* y, int nv, Epilogue ep, cudaStream_t s);
int batch_long_rows_device(const int32_t* d_row_ptr, int32_t rows, int32_t min_len, int64_t max_count, int32_t** d_out,
                           int64_t* count, cudaStream_t stream);

// ---- gemv.cu --------------------------------------------------------------------------------
struct DenseDev {
  int32_t rows = 0, cols = 0;
  int64_t ld = 0;           // leading dimension in floats (multiple of 4 so every row is 16-byte aligned)
  const float* a = nullptr;
};
int launch_gemv(const DenseDev& A, const float* x, float* y, Epilogue ep, int sm_count, cudaStream_t s);
// batch.cu: y [nv][rows] for nv <= 8 vectors in the panel layout written by launch_interleave_panels
int64_t panel_floats(int64_t ld);  // buffer size of that layout for eight vectors
int launch_interleave_panels(const float* x, int nv, int64_t n, int64_t ld, float* xp, cudaStream_t s);
int launch_gemm_lite(const DenseDev& A, const float* xp, float* y, int nv, Epilogue ep, cudaStream_t s);

}  // namespace hispmv
