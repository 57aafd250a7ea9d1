// TEST INFRASTRUCTURE (oracle/) -- C-ABI doorway into the UNMODIFIED reference host library
// /root/reference/common/src/spmv-helper.cpp (compiled where it lies, against the stub TAPA/XRT headers in
// oracle/stubs/, into oracle/_ref/libref_common.so):
//   HiSpmvHandle::cpuSequential             common/src/spmv-helper.cpp:812-833
//   HiSpmvHandle::tileAndPad(COO)           common/src/spmv-helper.cpp:139-227   (private)
//   HiSpmvHandle::prepareSparseMtxForFPGA   common/src/spmv-helper.cpp:648-715   (shared-row list)
//   HiSpmvHandle::loadMtx                   common/src/spmv-helper.cpp:34-136    (private)
//   HiSpmvHandle::getPreparedMtx            the packed 64-bit PEG streams (computeTileSize / prepareTile :429-638)
//   HiSpmvHandle::printErrorStats           common/src/spmv-helper.cpp:835-895   (the host program's error report)
// The private members are reached by re-declaring access for this translation unit only; the reference
// sources themselves are compiled untouched.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>

#define private public
#include "spmv-helper.h"  // reference header (common/include)
#undef private
#include "fpga-power.h"

// The XRT power sampler (common/src/fpga-power.cpp) is not built; give its symbols empty bodies so that
// spmv-helper.cpp's fpgaRun links.  fpgaRun is never called by the oracle.
FpgaPowerMonitor::FpgaPowerMonitor() : isMonitoring(false), debug(false) {}
FpgaPowerMonitor::~FpgaPowerMonitor() {}
void FpgaPowerMonitor::startMonitoring(const int, bool) {}
void FpgaPowerMonitor::stopMonitoring() {}
float FpgaPowerMonitor::getAveragePower(size_t& n) const { n = 0; return 0.f; }
float FpgaPowerMonitor::getMaxPower() const { return 0.f; }

namespace {
struct Quiet {  // the reference prints configuration banners on every call
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
HiSpmvHandle* make_handle(int a, int b, int c, int urams, int lat, int dense, int pre, int rdn) {
  return new HiSpmvHandle(a, b, c, 512, urams, lat, dense != 0, pre != 0, rdn != 0);
}
}  // namespace

extern "C" {

// Cout = alpha * (A*B) + beta * Cin, COO order, fp32 (the reference's own self-check oracle).
void ref_common_cpu_sequential(int rows, int cols, int64_t nnz, const int* r, const int* c, const float* v,
                               const float* x, const float* c_in, float alpha, float beta, float* c_out) {
  Quiet q;
  HiSpmvHandle* h = make_handle(24, 1, 1, 2, 5, 1, 0, 1);
  h->rows = rows;
  h->cols = cols;
  h->nnz = (int)nnz;
  h->coo_mtx.rows.assign(r, r + nnz);
  h->coo_mtx.cols.assign(c, c + nnz);
  h->coo_mtx.values.assign(v, v + nnz);
  std::vector<float> B(x, x + cols), Cin(c_in, c_in + rows), Cout(rows, 0.0f);
  h->cpuSequential(B, Cin, alpha, beta, Cout);
  std::memcpy(c_out, Cout.data(), sizeof(float) * rows);
  delete h;
}

// COO -> per-tile CSR through the reference's tileAndPad, flattened back to one global CSR
// (column tiles are visited in ascending order, so concatenating them keeps each row sorted).
int ref_common_coo_to_csr(int rows, int cols, int64_t nnz, const int* r, const int* c, const float* v, int* row_ptr,
                          int* col, float* val) {
  Quiet q;
  HiSpmvHandle* h = make_handle(24, 1, 1, 2, 5, 1, 0, 1);
  h->rows = rows;
  h->cols = cols;
  h->nnz = (int)nnz;
  COOMatrix_t coo;
  coo.rows.assign(r, r + nnz);
  coo.cols.assign(c, c + nnz);
  coo.values.assign(v, v + nnz);
  auto tiles = h->tileAndPad(coo);
  int64_t out = 0;
  row_ptr[0] = 0;
  for (int gr = 0; gr < rows; ++gr) {
    const int ti = gr / h->tile_rows, lr = gr % h->tile_rows;
    for (int tj = 0; tj < h->col_tiles; ++tj) {
      const CSRMatrix_t& t = tiles[ti][tj];
      for (int k = t.row_offsets[lr]; k < t.row_offsets[lr + 1]; ++k) {
        col[out] = tj * h->tile_cols + t.col_indices[k];
        val[out] = t.values[k];
        ++out;
      }
    }
    row_ptr[gr + 1] = (int)out;
  }
  delete h;
  return out == nnz ? 0 : -1;
}

// The reference's "shared rows" (rows processed by all PEs) for a given hardware configuration.
// Returns the count; ids (ascending) are written when out != NULL and capacity suffices.
int ref_common_shared_rows(int num_ch_a, int urams, int rows, int cols, int64_t nnz, const int* r, const int* c,
                           const float* v, int* out, int capacity) {
  Quiet q;
  HiSpmvHandle* h = make_handle(num_ch_a, 1, 1, urams, 5, 1, 0, 1);
  std::vector<int> rv(r, r + nnz), cv(c, c + nnz);
  std::vector<float> vv(v, v + nnz);
  h->prepareSparseMtxForFPGA(rows, cols, rv, cv, vv);
  int n = 0;
  for (int id : h->shared_row_indices) {
    if (out && n < capacity) out[n] = id;
    ++n;
  }
  delete h;
  return n;
}

static COOMatrix_t g_coo;
static int g_rows, g_cols;
int ref_common_load_mtx(const char* path, int* rows, int* cols, int64_t* nnz) {
  Quiet q;
  HiSpmvHandle* h = make_handle(24, 1, 1, 2, 5, 1, 0, 1);
  int st = 0;
  try {
    g_coo = h->loadMtx(path);
    g_rows = h->rows;
    g_cols = h->cols;
  } catch (const std::exception&) {
    st = -1;
  }
  delete h;
  if (st) return st;
  *rows = g_rows;
  *cols = g_cols;
  *nnz = (int64_t)g_coo.rows.size();
  return 0;
}
void ref_common_load_mtx_fetch(int* r, int* c, float* v) {
  std::memcpy(r, g_coo.rows.data(), sizeof(int) * g_coo.rows.size());
  std::memcpy(c, g_coo.cols.data(), sizeof(int) * g_coo.cols.size());
  std::memcpy(v, g_coo.values.data(), sizeof(float) * g_coo.values.size());
  g_coo = COOMatrix_t();
}

// The packed PEG streams of a COO matrix for one hardware configuration (prepareSparseMtxForFPGA -> prepareTile).
// meta = {num_pes, pes_per_ch, tile_rows, tile_cols, row_tiles, col_tiles, words_per_channel, shared rows};
// returns the number of channels.  ref_common_pack_fetch copies channel-major words and the shared-row ids.
static std::vector<uint64_t> g_stream;
static std::vector<int> g_shared;
int ref_common_pack(int num_ch_a, int urams, int fp_acc_latency, int pre_acc, int rows, int cols, int64_t nnz,
                    const int* r, const int* c, const float* v, int64_t* meta) {
  Quiet q;
  HiSpmvHandle* h = make_handle(num_ch_a, 1, 1, urams, fp_acc_latency, 1, pre_acc, 1);
  std::vector<int> rv(r, r + nnz), cv(c, c + nnz);
  std::vector<float> vv(v, v + nnz);
  h->prepareSparseMtxForFPGA(rows, cols, rv, cv, vv);
  const auto& prep = h->prep_mtx;
  const int64_t words = prep.empty() ? 0 : (int64_t)prep[0].size();
  g_stream.resize((size_t)(words * (int64_t)prep.size()));
  for (size_t ch = 0; ch < prep.size(); ++ch) std::memcpy(g_stream.data() + ch * words, prep[ch].data(), sizeof(uint64_t) * words);
  g_shared.assign(h->shared_row_indices.begin(), h->shared_row_indices.end());
  meta[0] = h->num_pes; meta[1] = h->pes_per_ch; meta[2] = h->tile_rows; meta[3] = h->tile_cols;
  meta[4] = h->row_tiles; meta[5] = h->col_tiles; meta[6] = words; meta[7] = (int64_t)g_shared.size();
  const int n_ch = (int)prep.size();
  delete h;
  return n_ch;
}
void ref_common_pack_fetch(uint64_t* stream, int* shared_rows) {
  std::memcpy(stream, g_stream.data(), sizeof(uint64_t) * g_stream.size());
  if (shared_rows && !g_shared.empty()) std::memcpy(shared_rows, g_shared.data(), sizeof(int) * g_shared.size());
  g_stream = std::vector<uint64_t>();
  g_shared = std::vector<int>();
}

// The text HiSpmvHandle::printErrorStats prints for (cpu_ref, fpga_out); returns its length (truncated to cap - 1).
int ref_common_error_stats(int n, const float* cpu_ref, const float* out, char* buf, int cap) {
  std::ostringstream sink;
  std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
  std::cout.unsetf(std::ios::floatfield);  // the reference leaves std::scientific set after a histogram: start clean
  std::cout.precision(6);
  HiSpmvHandle* h = nullptr;
  {
    std::ostringstream banner;  // the constructor prints the configuration
    std::streambuf* keep = std::cout.rdbuf(banner.rdbuf());
    h = make_handle(24, 1, 1, 2, 5, 1, 0, 1);
    std::cout.rdbuf(keep);
  }
  std::vector<float> a(cpu_ref, cpu_ref + n), b(out, out + n);
  h->printErrorStats(a, b);
  delete h;
  std::cout.rdbuf(old);
  const std::string text = sink.str();
  const int len = (int)std::min<size_t>(text.size(), (size_t)std::max(cap - 1, 0));
  if (cap > 0) {
    std::memcpy(buf, text.data(), (size_t)len);
    buf[len] = 0;
  }
  return len;
}
}
