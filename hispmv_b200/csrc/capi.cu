// C-ABI of libhispmv_cuda.so (see include/hispmv.h for the reference interfaces each entry replaces).
// Host-side state only: contexts, matrix handles, plans; all arithmetic happens in spmv.cu / gemv.cu /
// partition.cu.  No CPU compute path exists here on purpose.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <algorithm>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "internal.h"

namespace hispmv {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return HISPMV_OK;
  char buf[512];
  snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  g_last_error = buf;
  cudaGetLastError();  // clear the sticky-less error so later calls can proceed
  return e == cudaErrorMemoryAllocation ? HISPMV_FULL : HISPMV_ERR_CUDA;
}

constexpr int kMaxRunChunks = 16;  // row ranges the host-buffer call pipelines against PCIe copies

struct Matrix {
  bool dense = false;
  int32_t rows = 0, cols = 0;          // global shape
  int32_t row_begin = 0, row_end = 0;  // local row block
  int64_t nnz = 0;
  // sparse
  int32_t* d_row_ptr = nullptr;
  int32_t* d_col = nullptr;
  float* d_val = nullptr;
  RowStats stats{};
  int kernel = HISPMV_KERNEL_AUTO, lanes = 0;
  bool forced = false;
  int32_t tile_items = 0;
  int64_t num_tiles = 0;
  int32_t* d_tile_row = nullptr;
  int64_t* d_tile_nnz = nullptr;
  float* d_carry = nullptr;
  int32_t* d_split_rows = nullptr;
  int64_t num_split = 0;
  int32_t* d_tile_chunk = nullptr;      // ADAPTIVE
  unsigned int* d_counter = nullptr;    // ADAPTIVE (two lanes, like d_carry)
  TileDesc* d_desc = nullptr;           // ADAPTIVE / ROWSTAGE: resolved tile records
  std::vector<int32_t> h_tile_row, h_tile_chunk;  // host copies: row ranges for the pipelined host-buffer call
  // column slabs (x larger than L2): each slab is a CSR over the same rows with its own plan; y accumulates over them
  std::vector<Matrix*> slabs;
  int32_t slab_cols = 0;
  bool is_slab = false;
  int32_t long_threshold = 0, chunk_nnz = 0;
  int32_t hot_cols = 0x7fffffff;        // ADAPTIVE / ROWSTAGE: split L1 policy threshold for x gathers
  int32_t ahead = 0;                    // ADAPTIVE: L2 prefetch distance in tiles (0 = off), see AdaptivePlan::ahead
  int32_t ahead_all = 0;                // look ahead for STREAM tiles too (development sweeps)
  bool persistent = false;              // ADAPTIVE: one resident CTA per SM with x[0, hot_cols) in shared memory
  bool pipeline = false;                // ADAPTIVE: warp-specialised persistent pipeline (TMA ring)
  bool warptile = false;                // ADAPTIVE: one warp per (small) tile, no CTA barrier
  int adaptive_threads = 256;           // ADAPTIVE: threads per CTA (256; 128 through the development switch)
  int rowstage_threads = 128;           // ROWSTAGE: threads per CTA (128; 256 through the development switch)
  ColProbe probe{};                     // column-locality probe (selector input)
  int32_t* d_batch_long_rows = nullptr;  // batches: rows with more than kBatchMaxRowNnz nonzeros (one CTA each), built
  int64_t num_batch_long_rows = -1;      // on the first batch call (-1: not yet)
  PbArrays pb;                          // BLOCKED: the slab-major copy, segment table, pass-1 work ranges, products
  int64_t pb_slab_cost = 0;             // BLOCKED: entries one slab load is worth when pass-1 ranges are balanced
  int64_t pb_piece_cost16 = 0;          // ... and sixteenths of an entry one piece is worth
  int64_t slab_runs = 0;                // selector input: (row, slab) runs for slabs of kPbSlabCols columns (0: not counted)
  // dense
  float* d_a = nullptr;
  int64_t ld = 0;

  int32_t local_rows() const { return row_end - row_begin; }
  void free_plan() {
    cudaFree(d_tile_row);
    cudaFree(d_tile_nnz);
    cudaFree(d_carry);
    cudaFree(d_split_rows);
    cudaFree(d_tile_chunk);
    cudaFree(d_counter);
    cudaFree(d_desc);
    pb_free(&pb);
    d_desc = nullptr;
    d_tile_chunk = nullptr;
    d_counter = nullptr;
    d_tile_row = nullptr;
    d_tile_nnz = nullptr;
    d_carry = nullptr;
    d_split_rows = nullptr;
    num_tiles = 0;
    num_split = 0;
    tile_items = 0;
  }
  int64_t device_bytes() const {
    if (dense) return (int64_t)local_rows() * ld * 4;
    int64_t b = ((int64_t)local_rows() + 1) * 4 + (((nnz + 3) & ~3LL) + 4) * 8;
    if (d_tile_row) b += (num_tiles + 1) * 12 + num_tiles * 16 + num_split * 4 + (d_desc ? num_tiles * 32 : 0);
    if (pb.d_val)  // val + lcol + flags per entry; perm + one partial per stream lane in use per piece; tables
      b += pb.padded_nnz * 6 + pb.padded_nnz / 8 + pb.padded_nnz / kPbGroup * 4 + pb.num_pieces * (2 + 4 * (pb.d_part[1] ? 2 : 1)) +
           pb.num_seg * 16 + pb.num_chunks * 8 + pb.stage_total * 3 + pb.bit_words * 4 + (num_tiles + 1) * 8 + (pb.num_slabs + 1) * 4 + ((int64_t)local_rows() + 1) * 4;
    for (auto* sm : slabs) b += sm->device_bytes();
    return b;
  }
  ~Matrix() {
    for (auto* sm : slabs) delete sm;
    free_plan();
    cudaFree(d_row_ptr);
    cudaFree(d_col);
    cudaFree(d_val);
    cudaFree(d_a);
    cudaFree(d_batch_long_rows);
  }
};

}  // namespace hispmv

using namespace hispmv;

namespace {

// Parallel host memcpy for callers that hand in pageable memory (plain numpy arrays through pyhispmv): a few resident
// threads copy slices of the caller's vector into / out of the context's pinned ring while the previous slice crosses
// PCIe.  One copy at a time (the plugin is single-threaded, pyhispmv holds the GIL for the whole call).
// memcpy for megabyte pieces between pageable and pinned memory: non-temporal stores, so the destination lines are not
// read for ownership first (a plain memcpy of a 1 MB piece stays below glibc's own non-temporal threshold and moves
// three bytes over the memory bus for every byte copied)
static void stream_copy(void* dst, const void* src, size_t n) {
#if defined(__SSE2__)
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (n < 4096 + head) {
    memcpy(dst, src, n);
    return;
  }
  memcpy(d, s, head);
  d += head;
  s += head;
  n -= head;
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 0);
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 1);
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 2);
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s) + 3);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 0, a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 1, b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 2, c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d) + 3, e);
    s += 64;
    d += 64;
  }
  _mm_sfence();
  memcpy(d, s, n - blocks * 64);
#else
  memcpy(dst, src, n);
#endif
}

class HostCopier {
 public:
  explicit HostCopier(int n_threads) {
    for (int i = 0; i < n_threads; ++i) th_.emplace_back([this, i] { work(i); });
    tasks_.resize((size_t)n_threads);
  }
  ~HostCopier() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t parts = th_.size() + 1;
    if (bytes < (1u << 20) || th_.empty()) {
      memcpy(dst, src, bytes);
      return;
    }
    const size_t per = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
    {
      std::lock_guard<std::mutex> g(m_);
      for (size_t i = 0; i < th_.size(); ++i) {
        const size_t off = std::min(bytes, (i + 1) * per);
        const size_t end = std::min(bytes, (i + 2) * per);
        tasks_[i] = Task{static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, end - off};
      }
      pending_ = (int)th_.size();
      ++gen_;
    }
    cv_.notify_all();
    stream_copy(dst, src, std::min(bytes, per));
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  struct Task {
    char* d;
    const char* s;
    size_t n;
  };
  void work(int i) {
    uint64_t seen = 0;
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        t = tasks_[(size_t)i];
      }
      if (t.n) stream_copy(t.d, t.s, t.n);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::vector<Task> tasks_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  uint64_t gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

constexpr int kPinSlots = 3;
constexpr size_t kPinSlotBytes = 8u << 20;      // slice of a pageable vector in flight
constexpr int64_t kSmallCallBytes = 1 << 20;    // x + bias + y up to this size take the one-graph path

// One captured launch sequence of the small host-buffer call: H2D (x | bias) -> kernel(s) -> D2H y
struct SmallGraph {
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
};
struct SmallKey {
  const void* m;
  uint32_t alpha, beta;
  int has_bias;
  bool operator<(const SmallKey& o) const {
    if (m != o.m) return m < o.m;
    if (alpha != o.alpha) return alpha < o.alpha;
    if (beta != o.beta) return beta < o.beta;
    return has_bias < o.has_bias;
  }
};

}  // namespace

struct hispmv_ctx {
  int device = 0;
  int flags = 0;
  int sm_count = 148;
  // batches (several right-hand sides per pass): interleaved x per stream lane, and host-path staging for a group
  float* d_xi[2] = {nullptr, nullptr};
  int64_t cap_xi = 0;
  float *d_xb = nullptr, *d_yb = nullptr;
  int64_t cap_xb = 0, cap_yb = 0;
  int64_t l2_persist_max = 0, l2_window_max = 0;  // device limits for persisting L2 lines / access-policy windows
  bool l2_persist_on = false;                     // the persisting carve-out has been set aside (column slabs)
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second lane for pipelined linear(); H2D lane of the pipelined run()
  cudaStream_t stream3 = nullptr;  // D2H lane of the pipelined run()
  cudaEvent_t ev_pipe[2 * kMaxRunChunks + 2] = {};  // x ready, bias range ready x16, kernel range done x16, x head ready
  cudaEvent_t ev_bias = nullptr;
  int shard_part = 0, shard_parts = 1;
  int64_t mem_limit = 0;
  std::vector<Matrix*> mats;
  int selected = -1;
  bool committed = false;
  // device staging for the host-buffer calls (two lanes)
  float* d_x[2] = {nullptr, nullptr};
  float* d_y[2] = {nullptr, nullptr};
  float* d_bias = nullptr;
  int64_t cap_x = 0, cap_y = 0;
  // callers with pageable memory: pinned ring + copier threads (created on first use)
  char* h_pin[kPinSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_pin[kPinSlots] = {nullptr, nullptr, nullptr};
  HostCopier* copier = nullptr;
  // small calls (a DNN layer's vectors): pinned (x | bias | y) + device (x | bias), one CUDA graph per (matrix, alpha, beta)
  float* h_small = nullptr;
  float* d_small = nullptr;
  float* dh_small = nullptr;   // device view of h_small (mapped pinned memory)
  bool small_direct = false;
  int64_t cap_small = 0;            // floats in h_small / d_small
  std::map<SmallKey, SmallGraph> small_graphs;
  std::map<const void*, int> small_seen;   // eager runs before a matrix's calls are captured
  // single-process multi-GPU (hispmv_create_multi): this context owns one child per GPU, child k holding row block k of
  // every matrix; the parent holds no matrix of its own and only fans the host-buffer calls out
  std::vector<hispmv_ctx*> kids;

  void drop_small_graphs() {
    for (auto& kv : small_graphs) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    small_graphs.clear();
  }

  int64_t used_bytes() const {
    int64_t b = 0;
    for (auto* m : mats) b += m->device_bytes();
    return b;
  }
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess) ok = true;
    cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (ok) cudaSetDevice(prev);
  }
};

int ensure_staging(hispmv_ctx* c, int64_t n_x, int64_t n_y) {
  // the capacity is dropped to 0 before the old buffers go, so a failed grow can never leave a stale capacity behind
  // null (or mixed-size) buffers; scratch failures in run paths are CUDA errors, not the "matrix memory full" sentinel
  auto scratch = [](int st) { return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st; };
  if (n_x > c->cap_x) {
    c->cap_x = 0;
    for (int l = 0; l < 2; ++l) {
      cudaFree(c->d_x[l]);
      c->d_x[l] = nullptr;
    }
    for (int l = 0; l < 2; ++l) {
      int st = check_cuda(cudaMalloc((void**)&c->d_x[l], (size_t)n_x * 4), "cudaMalloc(x staging)", __FILE__, __LINE__);
      if (st != HISPMV_OK) return scratch(st);
    }
    c->cap_x = n_x;
  }
  if (n_y > c->cap_y) {
    c->cap_y = 0;
    for (int l = 0; l < 2; ++l) {
      cudaFree(c->d_y[l]);
      c->d_y[l] = nullptr;
    }
    cudaFree(c->d_bias);
    c->d_bias = nullptr;
    for (int l = 0; l < 2; ++l) {
      int st = check_cuda(cudaMalloc((void**)&c->d_y[l], (size_t)n_y * 4), "cudaMalloc(y staging)", __FILE__, __LINE__);
      if (st != HISPMV_OK) return scratch(st);
    }
    int st = check_cuda(cudaMalloc((void**)&c->d_bias, (size_t)n_y * 4), "cudaMalloc(bias staging)", __FILE__, __LINE__);
    if (st != HISPMV_OK) return scratch(st);
    c->cap_y = n_y;
  }
  return HISPMV_OK;
}

// ---- host-buffer plumbing ------------------------------------------------------------------------------------------
bool is_pageable(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int ensure_pin_ring(hispmv_ctx* c) {
  if (c->h_pin[0]) return HISPMV_OK;
  for (int i = 0; i < kPinSlots; ++i) {
    int st = check_cuda(cudaMallocHost((void**)&c->h_pin[i], kPinSlotBytes), "cudaMallocHost(pinned ring)", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaEventCreateWithFlags(&c->ev_pin[i], cudaEventDisableTiming), "event", __FILE__, __LINE__);
    if (st != HISPMV_OK) return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st;
  }
  if (!c->copier) {
    const unsigned hw = std::thread::hardware_concurrency();
    c->copier = new HostCopier((int)std::max(1u, std::min(7u, hw > 1 ? hw / 2 : 1u)));
  }
  return HISPMV_OK;
}

// pageable host -> device through the pinned ring: the copier threads fill slot k while slot k-1 crosses PCIe
int staged_h2d(hispmv_ctx* c, void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
  int st = ensure_pin_ring(c);
  if (st != HISPMV_OK) return st;
  int k = 0;
  for (size_t off = 0; off < bytes; off += kPinSlotBytes, ++k) {
    const int slot = k % kPinSlots;
    const size_t len = std::min(kPinSlotBytes, bytes - off);
    if (k >= kPinSlots) HISPMV_CUDA(cudaEventSynchronize(c->ev_pin[slot]));  // its previous slice has left the slot
    c->copier->copy(c->h_pin[slot], static_cast<const char*>(h_src) + off, len);
    HISPMV_CUDA(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, c->h_pin[slot], len, cudaMemcpyHostToDevice, s));
    HISPMV_CUDA(cudaEventRecord(c->ev_pin[slot], s));
  }
  // the ring is free again for whoever comes next (the slots still in flight are waited for here: at most kPinSlots)
  for (int i = 0; i < std::min(k, kPinSlots); ++i) HISPMV_CUDA(cudaEventSynchronize(c->ev_pin[i]));
  return HISPMV_OK;
}

// device -> pageable host: slice k+1 crosses PCIe while the copier threads move slice k out of its slot
int staged_d2h(hispmv_ctx* c, void* h_dst, const void* d_src, size_t bytes, cudaStream_t s) {
  int st = ensure_pin_ring(c);
  if (st != HISPMV_OK) return st;
  const int n = (int)((bytes + kPinSlotBytes - 1) / kPinSlotBytes);
  auto len_of = [&](int k) { return std::min(kPinSlotBytes, bytes - (size_t)k * kPinSlotBytes); };
  for (int k = 0; k <= n; ++k) {
    if (k < n) {
      const int slot = k % kPinSlots;
      HISPMV_CUDA(cudaMemcpyAsync(c->h_pin[slot], static_cast<const char*>(d_src) + (size_t)k * kPinSlotBytes, len_of(k),
                                  cudaMemcpyDeviceToHost, s));
      HISPMV_CUDA(cudaEventRecord(c->ev_pin[slot], s));
    }
    if (k >= 1) {
      const int slot = (k - 1) % kPinSlots;
      HISPMV_CUDA(cudaEventSynchronize(c->ev_pin[slot]));
      c->copier->copy(static_cast<char*>(h_dst) + (size_t)(k - 1) * kPinSlotBytes, c->h_pin[slot], len_of(k - 1));
    }
  }
  return HISPMV_OK;
}

int run_matrix(hispmv_ctx* c, Matrix* m, const float* d_x, const float* d_bias, float* d_y, float alpha, float beta,
               int relu, cudaStream_t s, int lane = 0, int64_t tile_begin = 0, int64_t tile_count = -1, int y_mc = 0,
               int phases = 3, int work_part = 0);

// The small host-buffer call (a DNN layer's vectors: tens of kilobytes).  Everything is latency here, so the call is
// one memcpy into pinned memory, H2D of (x | bias), the kernel(s), D2H of y, one stream synchronisation and one memcpy
// out.  With HISPMV_SMALL_GRAPH=1 the device sequence is captured once per (matrix, alpha, beta) and replayed as one
// graph launch (the first call on a matrix runs eagerly: lazy per-kernel attributes are set outside any capture).
int small_call(hispmv_ctx* c, Matrix* m, const float* x, const float* bias, float* y, float alpha, float beta) {
  const int64_t cols = m->cols, n_y = m->local_rows();
  const int64_t cpad = (cols + 3) & ~3LL, ypad = (n_y + 3) & ~3LL;
  const int64_t need = cpad + 2 * ypad + 4;
  if (need > c->cap_small) {
    c->drop_small_graphs();
    c->cap_small = 0;
    if (c->h_small) cudaFreeHost(c->h_small);
    cudaFree(c->d_small);
    c->h_small = nullptr;
    c->d_small = nullptr;
    const int64_t cap = std::max<int64_t>(need, 1 << 16);
    int st = check_cuda(cudaHostAlloc((void**)&c->h_small, (size_t)cap * 4, cudaHostAllocMapped), "cudaHostAlloc(small)",
                        __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)&c->d_small, (size_t)cap * 4), "cudaMalloc(small)", __FILE__, __LINE__);
    if (st != HISPMV_OK) return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st;
    c->dh_small = nullptr;  // the device's view of the pinned block (the same address under unified addressing)
    c->small_direct = cudaHostGetDevicePointer((void**)&c->dh_small, c->h_small, 0) == cudaSuccess && c->dh_small;
    if (!c->small_direct) cudaGetLastError();
    c->cap_small = cap;
  }
  float *hx = c->h_small, *hb = hx + cpad, *hy = hb + ypad;
  float *dx = c->d_small, *db = dx + cpad, *dy = db + ypad;
  if (cols > 0) memcpy(hx, x, (size_t)cols * 4);
  if (bias && n_y > 0) memcpy(hb, bias, (size_t)n_y * 4);
  cudaStream_t s = c->stream;
  // One-launch strategies that never read y back take bias straight from the pinned block and write y straight into it
  // (mapped host memory: coalesced reads, posted writes): the call is one small copy of x up, one kernel, one
  // synchronisation -- no copy engine on the way back.  HISPMV_SMALL_DIRECT=0 keeps the three-copy sequence.
  static const bool direct_ok = !(getenv("HISPMV_SMALL_DIRECT") && atoi(getenv("HISPMV_SMALL_DIRECT")) == 0);
  const bool one_launch = m->dense || (m->slabs.empty() && (m->kernel == HISPMV_KERNEL_ADAPTIVE || m->kernel == HISPMV_KERNEL_ROWSTAGE ||
                                                           m->kernel == HISPMV_KERNEL_CSR_SCALAR || m->kernel == HISPMV_KERNEL_CSR_VECTOR));
  const bool direct = direct_ok && one_launch && c->small_direct;
  const size_t up = (size_t)((bias && !direct) ? cpad + n_y : cols) * 4;
  auto enqueue = [&]() -> int {
    if (up > 0) HISPMV_CUDA(cudaMemcpyAsync(dx, hx, up, cudaMemcpyHostToDevice, s));
    int st = run_matrix(c, m, dx, bias ? (direct ? c->dh_small + cpad : db) : nullptr, direct ? c->dh_small + cpad + ypad : dy,
                        alpha, beta, 0, s, 0, 0, -1, 0, 3);
    if (st != HISPMV_OK) return st;
    if (!direct && n_y > 0) HISPMV_CUDA(cudaMemcpyAsync(hy, dy, (size_t)n_y * 4, cudaMemcpyDeviceToHost, s));
    return HISPMV_OK;
  };
  uint32_t ab, bb;
  memcpy(&ab, &alpha, 4);
  memcpy(&bb, &beta, 4);
  const SmallKey key{m, ab, bb, bias != nullptr};
  // measured on model_test's three layers: eager 173.7 us per pass, graph replay 184.5 (a graph launch with two memcpy
  // nodes costs more than three eager enqueues): opt-in
  const char* sg = getenv("HISPMV_SMALL_GRAPH");
  const bool no_graph = !(sg && atoi(sg) == 1);
  if (no_graph && !c->small_graphs.empty()) c->drop_small_graphs();
  auto it = c->small_graphs.find(key);
  if (it == c->small_graphs.end() && !no_graph && c->small_seen[m]++ >= 1) {
    SmallGraph g;
    HISPMV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int st = enqueue();
    const cudaError_t e = cudaStreamEndCapture(s, &g.graph);
    if (st == HISPMV_OK && e == cudaSuccess && g.graph &&
        cudaGraphInstantiate(&g.exec, g.graph, 0) == cudaSuccess) {
      it = c->small_graphs.emplace(key, g).first;
    } else {  // not capturable (should not happen): stay on the eager sequence for this matrix
      if (g.graph) cudaGraphDestroy(g.graph);
      cudaGetLastError();
      c->small_seen[m] = -(1 << 30);
    }
  }
  if (it != c->small_graphs.end()) {
    HISPMV_CUDA(cudaGraphLaunch(it->second.exec, s));
  } else {
    const int st = enqueue();
    if (st != HISPMV_OK) return st;
  }
  HISPMV_CUDA(cudaStreamSynchronize(s));
  if (n_y > 0) memcpy(y, hy, (size_t)n_y * 4);
  return HISPMV_OK;
}

// ---- single-process multi-GPU --------------------------------------------------------------------------------------
// fn(kid, k) on every child, one host thread per GPU (the calls block: copies inside); the first failure wins and its
// error text is carried over to the calling thread
template <class F>
int for_each_kid(hispmv_ctx* c, F fn) {
  const size_t n = c->kids.size();
  std::vector<int> st(n, HISPMV_OK);
  std::vector<std::string> err(n);
  std::vector<std::thread> th;
  for (size_t k = 1; k < n; ++k)
    th.emplace_back([&, k] {
      st[k] = fn(c->kids[k], (int)k);
      if (st[k] < 0) err[k] = g_last_error;
    });
  st[0] = fn(c->kids[0], 0);
  if (st[0] < 0) err[0] = g_last_error;
  for (auto& t : th) t.join();
  for (size_t k = 0; k < n; ++k)
    if (st[k] < 0) {
      g_last_error = err[k];
      return st[k];
    }
  return st[0];
}

// add a matrix from HOST arrays to every child (each keeps its own row block); on any failure the children that had
// taken it drop it again, so the handle indices stay aligned
template <class F>
int multi_add(hispmv_ctx* c, F add) {
  const int r = for_each_kid(c, [&](hispmv_ctx* kid, int) { return add(kid); });
  const size_t want = c->kids[0]->mats.size();
  bool aligned = true;
  for (auto* kid : c->kids) aligned = aligned && kid->mats.size() == want;
  if (r < 0 || !aligned) {
    size_t least = want;
    for (auto* kid : c->kids) least = std::min(least, kid->mats.size());
    for (auto* kid : c->kids)
      while (kid->mats.size() > least) {
        DeviceGuard g(kid->device);
        delete kid->mats.back();
        kid->mats.pop_back();
      }
    return r < 0 ? r : HISPMV_ERR_STATE;
  }
  return r;
}

int multi_refuse(const char* what) {
  set_error(std::string(what) + ": not available on a multi-GPU handle (hispmv_create_multi); use hispmv_multi_child "
            "for the per-GPU contexts");
  return HISPMV_ERR_STATE;
}

// Build (or rebuild) the execution plan of a sparse matrix: histogram -> selector -> merge tiles.
int plan_sparse(hispmv_ctx* c, Matrix* m) {
  m->free_plan();
  int st = row_stats_device(m->d_row_ptr, m->local_rows(), &m->stats, c->stream);
  if (st != HISPMV_OK) return st;
  st = col_probe_device(m->d_row_ptr, m->d_col, m->local_rows(), &m->probe, c->stream);
  if (st != HISPMV_OK) return st;
  if (!m->forced) {
    select_kernel(m->stats, m->probe, (c->flags & HISPMV_FLAG_ROW_DIST_NET) != 0, &m->kernel, &m->lanes);
  } else if (m->nnz == 0 || m->local_rows() == 0) {
    m->kernel = HISPMV_KERNEL_EMPTY;
  }
  for (auto* sm : m->slabs) delete sm;
  m->slabs.clear();
  m->slab_cols = 0;
  if (!m->forced && m->kernel == HISPMV_KERNEL_ADAPTIVE && !m->is_slab) {
    // scattered gathers over a large x: two streaming passes with x served from shared memory (blocked.cu)
    // (the run count costs a pass over the column indices, so it is taken only where the cheap conditions already hold)
    m->slab_runs = 0;
    int want = select_blocked(m->local_rows(), m->cols, m->nnz, 0, m->probe, (c->flags & HISPMV_FLAG_ROW_DIST_NET) != 0);
    if (want) {
      st = pb_count_runs_device(m->d_row_ptr, m->d_col, m->local_rows(), kPbSlabCols, &m->slab_runs, c->stream);
      if (st != HISPMV_OK) return st;
      want = select_blocked(m->local_rows(), m->cols, m->nnz, m->slab_runs, m->probe,
                            (c->flags & HISPMV_FLAG_ROW_DIST_NET) != 0);
    }
    if (const char* e = getenv("HISPMV_BLOCKED_AUTO")) want = want && atoi(e) != 0;  // "0": keep the one-pass kernels
    if (want) m->kernel = HISPMV_KERNEL_BLOCKED;
  }
  if (m->kernel == HISPMV_KERNEL_ADAPTIVE && !m->is_slab) {
    int32_t w = select_slab_cols(m->cols, m->nnz, m->probe);
    if (const char* e = getenv("HISPMV_SLAB_COLS")) w = std::max(0, atoi(e));  // tests / sweeps
    if (w > 0 && w < m->cols) {
      m->slab_cols = w;
      for (int64_t lo = 0; lo < m->cols; lo += w) {
        Matrix* sm = new Matrix();
        sm->is_slab = true;
        sm->forced = true;
        sm->kernel = HISPMV_KERNEL_ADAPTIVE;
        if (const char* e = getenv("HISPMV_SLAB_KERNEL")) {  // development: "scalar" / "vector2" / "vector4"
          if (!strcmp(e, "scalar")) sm->kernel = HISPMV_KERNEL_CSR_SCALAR;
          if (!strcmp(e, "vector2")) { sm->kernel = HISPMV_KERNEL_CSR_VECTOR; sm->lanes = 2; }
          if (!strcmp(e, "vector4")) { sm->kernel = HISPMV_KERNEL_CSR_VECTOR; sm->lanes = 4; }
        }
        sm->rows = m->rows;
        sm->cols = m->cols;
        sm->row_begin = m->row_begin;
        sm->row_end = m->row_end;
        m->slabs.push_back(sm);
        st = csr_column_slab_device(m->d_row_ptr, m->d_col, m->d_val, m->local_rows(), (int32_t)lo,
                                    (int32_t)std::min<int64_t>(lo + w, m->cols), &sm->d_row_ptr, &sm->d_col, &sm->d_val,
                                    &sm->nnz, c->stream);
        if (st == HISPMV_OK) st = plan_sparse(c, sm);
        if (st != HISPMV_OK) return st;
      }
      return HISPMV_OK;  // the parent keeps its CSR for introspection; the slabs carry the plans
    }
  }
  if (m->kernel == HISPMV_KERNEL_CSR_VECTOR && m->lanes < 2) {
    int k, l;
    select_kernel(m->stats, m->probe, 0, &k, &l);
    m->lanes = l ? l : 2;
  }
  if (m->kernel == HISPMV_KERNEL_MERGE) {
    m->tile_items = merge_tile_items_for(m->stats);
    if (const char* e = getenv("HISPMV_MERGE_TILE")) {
      const int v = atoi(e);
      if (merge_tile_items_supported(v)) m->tile_items = v;
    }
    st = merge_tiles_device(m->d_row_ptr, m->local_rows(), m->nnz, m->tile_items, &m->num_tiles, &m->d_tile_row,
                            &m->d_tile_nnz, c->stream);
    if (st != HISPMV_OK) return st;
    HISPMV_CUDA(cudaMalloc((void**)&m->d_carry, (size_t)m->num_tiles * 2 * sizeof(float)));
    st = split_rows_device(m->d_row_ptr, m->local_rows(), m->d_tile_row, m->d_tile_nnz, m->num_tiles,
                           &m->d_split_rows, &m->num_split, c->stream);
    if (st != HISPMV_OK) return st;
  }
  if (m->kernel == HISPMV_KERNEL_BLOCKED) {
    int32_t W = kPbSlabCols;
    m->tile_items = kPbPanelItems;
    m->long_threshold = kPbLongThreshold;
    m->chunk_nnz = kPbChunkNnz;
    m->pb_slab_cost = kPbSlabCost;
    m->pb_piece_cost16 = kPbPieceCost16;
    if (const char* e = getenv("HISPMV_BLOCKED")) {  // "W,B,T,CH[,SLABCOST[,PIECECOST16]]" (development sweeps)
      int w = 0, b = 0, t = 0, ch = 0, sc = -1, pc = -1;
      if (sscanf(e, "%d,%d,%d,%d,%d,%d", &w, &b, &t, &ch, &sc, &pc) >= 4 && w >= 1024 && w <= kPbMaxSlabCols && (w & 3) == 0 &&
          b >= 256 && t >= 16 && b + t <= 16384 && ch >= 128 && ch <= 65535) {
        W = w;
        m->tile_items = b;
        m->long_threshold = t;
        m->chunk_nnz = ch;
        if (sc >= 0) m->pb_slab_cost = sc;
        if (pc >= 0) m->pb_piece_cost16 = pc;
      }
    }
    if (m->nnz <= 0 || m->local_rows() <= 0) {
      m->kernel = HISPMV_KERNEL_EMPTY;
    } else {
      // slab-major copy and its pieces first: the panels are cut over the per-row PIECE counts (prow_ptr), not nonzeros
      st = pb_order_device(m->d_row_ptr, m->d_col, m->d_val, m->local_rows(), m->cols, m->nnz, W, &m->pb, c->stream);
      if (st != HISPMV_OK) return st;
      st = adaptive_tiles_device(m->pb.d_prow_ptr, m->local_rows(), m->tile_items, m->long_threshold, m->chunk_nnz,
                                 &m->num_tiles, &m->d_tile_row, &m->d_tile_chunk, &m->d_split_rows, &m->num_split,
                                 c->stream);
      if (st != HISPMV_OK) return st;
      st = tile_desc_device(m->pb.d_prow_ptr, m->d_tile_row, m->d_tile_chunk, m->num_tiles, m->chunk_nnz, &m->d_desc,
                            c->stream);
      if (st != HISPMV_OK) return st;
      m->h_tile_row.assign((size_t)m->num_tiles + 1, 0);
      m->h_tile_chunk.assign((size_t)std::max<int64_t>(m->num_tiles, 1), -1);
      HISPMV_CUDA(cudaMemcpyAsync(m->h_tile_row.data(), m->d_tile_row, (size_t)(m->num_tiles + 1) * 4,
                                  cudaMemcpyDeviceToHost, c->stream));
      HISPMV_CUDA(cudaMemcpyAsync(m->h_tile_chunk.data(), m->d_tile_chunk, (size_t)m->num_tiles * 4,
                                  cudaMemcpyDeviceToHost, c->stream));
      const size_t n = (size_t)std::max<int64_t>(m->num_tiles, 1) * 2;
      HISPMV_CUDA(cudaMalloc((void**)&m->d_carry, n * sizeof(float)));
      HISPMV_CUDA(fill_u32_device(reinterpret_cast<uint32_t*>(m->d_carry), kCarryEmptyBits, n, c->stream));
      HISPMV_CUDA(cudaMalloc((void**)&m->d_counter, (n + 8) * sizeof(unsigned int)));
      HISPMV_CUDA(cudaMemsetAsync(m->d_counter, 0, (n + 8) * sizeof(unsigned int), c->stream));
      st = pb_segments_device(&m->pb, m->d_desc, m->num_tiles, m->local_rows(), c->stream);
      if (st != HISPMV_OK) return st;
      st = pb_make_work(&m->pb, c->sm_count, m->pb_slab_cost, m->pb_piece_cost16, c->stream);
      if (st != HISPMV_OK) return st;
      HISPMV_CUDA(cudaMalloc((void**)&m->pb.d_part[0], ((size_t)m->pb.num_pieces + 64) * 4));
    }
  }
  if (m->kernel == HISPMV_KERNEL_ADAPTIVE || m->kernel == HISPMV_KERNEL_ROWSTAGE) {
    if (m->kernel == HISPMV_KERNEL_ADAPTIVE) {
      m->tile_items = kAdaptiveStreamItems;
      m->long_threshold = kAdaptiveLongThreshold;
      m->chunk_nnz = kAdaptiveChunkNnz;
      m->adaptive_threads = 256;
      if (const char* e = getenv("HISPMV_ADAPTIVE")) {  // "B,T,CH[,THREADS]" (development sweeps)
        int b = 0, t = 0, ch = 0, th = 256;
        if (sscanf(e, "%d,%d,%d,%d", &b, &t, &ch, &th) >= 3 && b >= 128 && t >= 16 && b + t <= 4096 && ch >= 512 &&
            ch <= 65536 && (th == 128 || th == 256)) {
          m->tile_items = b;
          m->long_threshold = t;
          m->chunk_nnz = ch;
          m->adaptive_threads = th;
        }
      }
    } else {
      rowstage_params(m->stats, m->lanes, &m->lanes, &m->tile_items, &m->long_threshold, &m->chunk_nnz);
      m->rowstage_threads = 128;
      if (const char* e = getenv("HISPMV_ROWSTAGE")) {  // "LANES,B,T,CH[,THREADS]" (development sweeps)
        int l = 0, b = 0, t = 0, ch = 0, th = 256;
        if (sscanf(e, "%d,%d,%d,%d,%d", &l, &b, &t, &ch, &th) >= 4 && (th == 128 || th == 256) && l >= 1 && l <= 32 && (l & (l - 1)) == 0 && b >= 128 &&
            t >= 16 && b + t <= kRowstageMaxCap && ch >= 1024) {
          m->lanes = l;
          m->tile_items = b;
          m->long_threshold = t;
          m->chunk_nnz = ch;
          m->rowstage_threads = th;
        }
      }
    }
    m->hot_cols = 0x7fffffff;
    m->persistent = false;
    m->pipeline = false;
    m->warptile = false;
    static const bool research = experimental_kernels_built();  // the research switches need `make EXPERIMENTAL=1`
    if (const char* e = research ? getenv("HISPMV_WARPTILE") : nullptr) {  // "B,T,CH"
      int b = 0, t = 0, ch = 0;
      if (m->kernel == HISPMV_KERNEL_ADAPTIVE && sscanf(e, "%d,%d,%d", &b, &t, &ch) == 3 && b >= 32 && t >= 16 &&
          b + t <= kWarpTileCap && ch >= 64) {
        m->warptile = true;
        m->tile_items = b;
        m->long_threshold = t;
        m->chunk_nnz = ch;
      }
    }
    if (const char* e = research ? getenv("HISPMV_PIPELINE") : nullptr) {
      if (atoi(e) > 0 && m->kernel == HISPMV_KERNEL_ADAPTIVE) {
        m->pipeline = true;
        m->tile_items = std::min(m->tile_items, kPipelineRows);
        m->long_threshold = std::min(m->long_threshold, kPipelineCap - m->tile_items);
        m->chunk_nnz = std::min(m->chunk_nnz, kPipelineCap);
      }
    }
    if (const char* e = research ? getenv("HISPMV_PERSIST") : nullptr) {  // research switch: the persistent x-window variant (never auto)
      const int h = atoi(e);
      if (m->kernel == HISPMV_KERNEL_ADAPTIVE) {
        m->persistent = h > 0;
        if (h > 0) m->hot_cols = std::min(std::min(h, m->cols), kPersistentMaxHot);
        else if (m->hot_cols <= kPersistentMaxHot) m->hot_cols = 0x7fffffff;
      }
    }
    if (const char* e = getenv("HISPMV_HOT")) {  // development sweeps: L1 split threshold of the non-persistent kernels
      const int h = atoi(e);
      if (h != 0 && !m->persistent) m->hot_cols = h;  // negative values: diagnostics (builds with -DHISPMV_DIAG)
    }
    st = adaptive_tiles_device(m->d_row_ptr, m->local_rows(), m->tile_items, m->long_threshold, m->chunk_nnz,
                               &m->num_tiles, &m->d_tile_row, &m->d_tile_chunk, &m->d_split_rows, &m->num_split,
                               c->stream);
    if (st != HISPMV_OK) return st;
    // L2 look-ahead (AdaptivePlan::ahead): thread THREADS-1 of every CTA asks L2 for the col/val range of the LONG
    // tile one CTA-per-SM ahead.  A LONG tile is four dependent load -> gather rounds per thread; with its stream
    // already in L2 each round is ~500 cycles shorter (C2: 0.384 -> 0.374 ms).  STREAM tiles are gather-bound and
    // lose 2 % when prefetched (C5/10: 0.432 -> 0.442 ms), so they are left alone, and matrices without long rows
    // never look ahead.
    m->ahead = 0;
    m->ahead_all = 0;
    if (m->kernel == HISPMV_KERNEL_ADAPTIVE && !m->warptile && !m->pipeline && !m->persistent) {
      if (m->stats.max_row_nnz >= m->long_threshold) m->ahead = c->sm_count;
      if (const char* e = getenv("HISPMV_AHEAD")) {  // development sweeps: "-1" off, "-2" own tile, "N[,all]"
        int a = 0;
        char all[8] = {0};
        const int got = sscanf(e, "%d,%7s", &a, all);
        if (got >= 1) {
          if (a == -1) m->ahead = 0;
          else if (a == -2) m->ahead = -1;
          else if (a > 0) m->ahead = a;
          m->ahead_all = got == 2 && all[0] == 'a';
        }
      }
    }
    st = tile_desc_device(m->d_row_ptr, m->d_tile_row, m->d_tile_chunk, m->num_tiles, m->chunk_nnz, &m->d_desc,
                          c->stream);
    if (st != HISPMV_OK) return st;
    m->h_tile_row.assign((size_t)m->num_tiles + 1, 0);
    m->h_tile_chunk.assign((size_t)std::max<int64_t>(m->num_tiles, 1), -1);
    HISPMV_CUDA(cudaMemcpyAsync(m->h_tile_row.data(), m->d_tile_row, (size_t)(m->num_tiles + 1) * 4,
                                cudaMemcpyDeviceToHost, c->stream));
    if (m->num_tiles)
      HISPMV_CUDA(cudaMemcpyAsync(m->h_tile_chunk.data(), m->d_tile_chunk, (size_t)m->num_tiles * 4,
                                  cudaMemcpyDeviceToHost, c->stream));
    const size_t n = (size_t)std::max<int64_t>(m->num_tiles, 1) * 2;
    HISPMV_CUDA(cudaMalloc((void**)&m->d_carry, n * sizeof(float)));
    HISPMV_CUDA(fill_u32_device(reinterpret_cast<uint32_t*>(m->d_carry), kCarryEmptyBits, n, c->stream));
    // two lanes of arrival counters, then two lanes of 4 scheduler words for the persistent kernel
    HISPMV_CUDA(cudaMalloc((void**)&m->d_counter, (n + 8) * sizeof(unsigned int)));
    HISPMV_CUDA(cudaMemsetAsync(m->d_counter, 0, (n + 8) * sizeof(unsigned int), c->stream));
  }
  HISPMV_CUDA(cudaStreamSynchronize(c->stream));
  return HISPMV_OK;
}

int check_capacity(hispmv_ctx* c, int64_t extra) {
  if (c->mem_limit > 0 && c->used_bytes() + extra > c->mem_limit) {
    set_error("device memory limit reached");
    return HISPMV_FULL;
  }
  return HISPMV_OK;
}

// Takes ownership of a full-matrix device CSR, shards it if requested, plans it, registers the handle.
int adopt_csr(hispmv_ctx* c, int32_t* d_row_ptr, int32_t* d_col, float* d_val, int64_t nnz, int32_t rows,
              int32_t cols) {
  Matrix* m = new Matrix();
  m->rows = rows;
  m->cols = cols;
  m->d_row_ptr = d_row_ptr;
  m->d_col = d_col;
  m->d_val = d_val;
  m->nnz = nnz;
  m->row_begin = 0;
  m->row_end = rows;
  int st = HISPMV_OK;
  if (c->shard_parts > 1) {
    std::vector<int32_t> bounds(c->shard_parts + 1);
    st = shard_bounds_device(d_row_ptr, rows, nnz, c->shard_parts, bounds.data(), c->stream);
    if (st == HISPMV_OK) {
      int32_t *rp = nullptr, *cl = nullptr;
      float* vl = nullptr;
      int64_t lnnz = 0;
      const int32_t rb = bounds[c->shard_part], re = bounds[c->shard_part + 1];
      st = csr_slice_device(d_row_ptr, d_col, d_val, rb, re, &rp, &cl, &vl, &lnnz, c->stream);
      if (st == HISPMV_OK) {
        cudaStreamSynchronize(c->stream);
        cudaFree(m->d_row_ptr);
        cudaFree(m->d_col);
        cudaFree(m->d_val);
        m->d_row_ptr = rp;
        m->d_col = cl;
        m->d_val = vl;
        m->nnz = lnnz;
        m->row_begin = rb;
        m->row_end = re;
      }
    }
  }
  if (st == HISPMV_OK) st = check_capacity(c, m->device_bytes());
  if (st == HISPMV_OK) st = plan_sparse(c, m);
  if (st != HISPMV_OK) {
    delete m;
    return st;
  }
  c->mats.push_back(m);
  return (int)c->mats.size() - 1;
}

int add_coo_common(hispmv_ctx* c, const int32_t* r, const int32_t* cc, const float* v, int64_t nnz, int32_t rows,
                   int32_t cols, bool on_device) {
  if (!c || rows < 0 || cols < 0 || nnz < 0 || (nnz > 0 && (!r || !cc || !v))) {
    set_error("add_sparse_coo: bad arguments");
    return HISPMV_ERR_ARG;
  }
  if (!c->kids.empty()) {
    if (on_device) return multi_refuse("add_sparse_coo_dev");
    return multi_add(c, [&](hispmv_ctx* kid) { return add_coo_common(kid, r, cc, v, nnz, rows, cols, false); });
  }
  DeviceGuard g(c->device);
  int st = check_capacity(c, nnz * 8 + ((int64_t)rows + 1) * 4);
  if (st != HISPMV_OK) return st;
  int32_t *d_r = nullptr, *d_c = nullptr;
  float* d_v = nullptr;
  const int32_t *ur = r, *uc = cc;
  const float* uv = v;
  if (!on_device && nnz > 0) {
    st = check_cuda(cudaMalloc((void**)&d_r, nnz * 4), "cudaMalloc(coo rows)", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)&d_c, nnz * 4), "cudaMalloc(coo cols)", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)&d_v, nnz * 4), "cudaMalloc(coo vals)", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMemcpyAsync(d_r, r, nnz * 4, cudaMemcpyHostToDevice, c->stream), "H2D rows", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMemcpyAsync(d_c, cc, nnz * 4, cudaMemcpyHostToDevice, c->stream), "H2D cols", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = check_cuda(cudaMemcpyAsync(d_v, v, nnz * 4, cudaMemcpyHostToDevice, c->stream), "H2D vals", __FILE__, __LINE__);
    ur = d_r;
    uc = d_c;
    uv = d_v;
  }
  int32_t *rp = nullptr, *cl = nullptr;
  float* vl = nullptr;
  if (st == HISPMV_OK) st = coo_to_csr_device(ur, uc, uv, nnz, rows, cols, &rp, &cl, &vl, c->stream);
  cudaFree(d_r);
  cudaFree(d_c);
  cudaFree(d_v);
  if (st != HISPMV_OK) return st;
  return adopt_csr(c, rp, cl, vl, nnz, rows, cols);
}

int add_csr_common(hispmv_ctx* c, const int32_t* row_ptr, const int32_t* col, const float* val, int32_t rows,
                   int32_t cols, bool on_device) {
  if (!c || rows < 0 || cols < 0 || !row_ptr) {
    set_error("add_sparse_csr: bad arguments");
    return HISPMV_ERR_ARG;
  }
  if (!c->kids.empty()) {
    if (on_device) return multi_refuse("add_sparse_csr_dev");
    return multi_add(c, [&](hispmv_ctx* kid) { return add_csr_common(kid, row_ptr, col, val, rows, cols, false); });
  }
  DeviceGuard g(c->device);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  int32_t last = 0, first = 0;
  if (on_device) {
    HISPMV_CUDA(cudaMemcpy(&last, row_ptr + rows, 4, cudaMemcpyDeviceToHost));
    HISPMV_CUDA(cudaMemcpy(&first, row_ptr, 4, cudaMemcpyDeviceToHost));
  } else {
    last = row_ptr[rows];
    first = row_ptr[0];
  }
  if (first != 0 || last < 0) {
    set_error("add_sparse_csr: row_ptr must start at 0 and be non-decreasing");
    return HISPMV_ERR_ARG;
  }
  const int64_t nnz = last;
  int st = check_capacity(c, nnz * 8 + ((int64_t)rows + 1) * 4);
  if (st != HISPMV_OK) return st;
  int32_t* rp = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)&rp, ((size_t)rows + 1 + 4) * 4));  // +4: 16-byte TMA windows may overrun the end
  st = check_cuda(cudaMemcpyAsync(rp, row_ptr, ((size_t)rows + 1) * 4, kind, c->stream), "copy row_ptr", __FILE__, __LINE__);
  int32_t* cl = nullptr;
  float* vl = nullptr;
  if (st == HISPMV_OK) st = alloc_padded_nnz_arrays(col, val, nnz, kind, &cl, &vl, c->stream);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamSynchronize(c->stream), "sync", __FILE__, __LINE__);
  // the contract of include/hispmv.h: row_ptr non-decreasing from 0 to nnz and every column inside [0, cols) -- refused
  // otherwise (the COO path refuses bad indices the same way); rows whose columns are out of order are re-sorted by
  // (column, value) exactly as the COO path orders them, because the column-slab and blocked plans cut rows by column
  int flags[3] = {0, 0, 0};
  if (st == HISPMV_OK) st = csr_validate_device(rp, cl, rows, cols, nnz, flags, c->stream);
  if (st == HISPMV_OK && (flags[0] || flags[1])) {
    set_error(flags[0] ? "add_sparse_csr: row_ptr must start at 0, be non-decreasing and end at nnz"
                       : "add_sparse_csr: a column index is outside [0,cols)");
    st = HISPMV_ERR_ARG;
  }
  if (st == HISPMV_OK && flags[2]) {
    int32_t* d_rows = nullptr;
    st = check_cuda(cudaMalloc((void**)&d_rows, (size_t)std::max<int64_t>(nnz, 1) * 4), "cudaMalloc(rows)", __FILE__, __LINE__);
    if (st == HISPMV_OK) st = csr_expand_rows_device(rp, rows, nnz, d_rows, c->stream);
    int32_t *rp2 = nullptr, *cl2 = nullptr;
    float* vl2 = nullptr;
    if (st == HISPMV_OK) st = coo_to_csr_device(d_rows, cl, vl, nnz, rows, cols, &rp2, &cl2, &vl2, c->stream);
    cudaFree(d_rows);
    if (st == HISPMV_OK) {
      cudaFree(rp);
      cudaFree(cl);
      cudaFree(vl);
      rp = rp2;
      cl = cl2;
      vl = vl2;
    }
  }
  if (st != HISPMV_OK) {
    cudaFree(rp);
    cudaFree(cl);
    cudaFree(vl);
    return st;
  }
  return adopt_csr(c, rp, cl, vl, nnz, rows, cols);
}

int add_dense_common(hispmv_ctx* c, const float* a, int32_t rows, int32_t cols, bool on_device) {
  if (!c || rows < 0 || cols < 0 || ((int64_t)rows * cols > 0 && !a)) {
    set_error("add_dense: bad arguments");
    return HISPMV_ERR_ARG;
  }
  if (!(c->flags & HISPMV_FLAG_DENSE_OVERLAY)) {
    // the reference asserts dense_overlay in prepareDenseMtxForFPGA (common/src/spmv-helper.cpp:718)
    set_error("create_dense_handle needs dense_overlay=True");
    return HISPMV_ERR_STATE;
  }
  if (!c->kids.empty()) {
    if (on_device) return multi_refuse("add_dense_dev");
    return multi_add(c, [&](hispmv_ctx* kid) { return add_dense_common(kid, a, rows, cols, false); });
  }
  DeviceGuard g(c->device);
  Matrix* m = new Matrix();
  m->dense = true;
  m->kernel = HISPMV_KERNEL_GEMV;
  m->rows = rows;
  m->cols = cols;
  m->row_begin = (int32_t)(((int64_t)rows * c->shard_part) / c->shard_parts);
  m->row_end = (int32_t)(((int64_t)rows * (c->shard_part + 1)) / c->shard_parts);
  m->ld = ((int64_t)cols + 3) & ~3LL;
  if (m->ld == 0) m->ld = 4;
  m->nnz = (int64_t)m->local_rows() * cols;
  int st = check_capacity(c, m->device_bytes());
  const size_t bytes = (size_t)std::max<int64_t>(1, (int64_t)m->local_rows() * m->ld) * 4;
  if (st == HISPMV_OK) st = check_cuda(cudaMalloc((void**)&m->d_a, bytes), "cudaMalloc(dense)", __FILE__, __LINE__);
  if (st == HISPMV_OK && m->ld != cols) st = check_cuda(cudaMemsetAsync(m->d_a, 0, bytes, c->stream), "memset", __FILE__, __LINE__);
  if (st == HISPMV_OK && m->local_rows() > 0 && cols > 0)
    st = check_cuda(cudaMemcpy2DAsync(m->d_a, (size_t)m->ld * 4, a + (int64_t)m->row_begin * cols, (size_t)cols * 4,
                                      (size_t)cols * 4, (size_t)m->local_rows(),
                                      on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream),
                    "copy dense", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamSynchronize(c->stream), "sync", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    delete m;
    return st;
  }
  c->mats.push_back(m);
  return (int)c->mats.size() - 1;
}

// `lane` selects one of the two carry / counter sets so that linear()'s two stream lanes never share them.
// Device-pointer callers get lane 0: one run in flight per matrix handle, as with the reference's xrt::run.
// `phases`: BLOCKED only -- bit 0 runs pass 1 (products of the whole matrix), bit 1 pass 2 over the given panels.
int run_matrix(hispmv_ctx* c, Matrix* m, const float* d_x, const float* d_bias, float* d_y, float alpha, float beta,
               int relu, cudaStream_t s, int lane, int64_t tile_begin, int64_t tile_count, int y_mc, int phases,
               int work_part) {
  Epilogue ep{alpha, beta, d_bias, relu};
  ep.y_mc = y_mc;
  if (y_mc && !m->dense && (m->kernel == HISPMV_KERNEL_MERGE || !m->slabs.empty() || m->pipeline)) {
    // these read y back (carry fix-up, slab accumulation) or keep their own store path: a multicast address is write-only
    set_error("run: this matrix's strategy cannot write y through a multicast address");
    return HISPMV_ERR_STATE;
  }
  if (beta != 0.0f && !d_bias) {
    set_error("run: bias is required when beta != 0");
    return HISPMV_ERR_ARG;
  }
  if (m->dense) {
    DenseDev D;
    D.rows = m->local_rows();
    D.cols = m->cols;
    D.ld = m->ld;
    D.a = m->d_a;
    return launch_gemv(D, d_x, d_y, ep, c->sm_count, s);
  }
  if (!m->slabs.empty()) {
    // y = alpha*A_0 x + beta*bias, then y += alpha*A_s x for the other slabs (the kernels read bias[r] and write y[r]
    // from the same thread, so y can be its own bias); ReLU only after the last slab
    // HISPMV_L2_PERSIST=1: mark the x slab a pass gathers from as persisting in L2 and everything else on the stream as
    // streaming (cudaStreamAttributeAccessPolicyWindow), so the 8-bytes-per-nonzero stream cannot push the slab out
    static const int want_persist = getenv("HISPMV_L2_PERSIST") ? atoi(getenv("HISPMV_L2_PERSIST")) : 0;
    const bool persist = want_persist > 0 && c->l2_persist_max > 0 && c->l2_window_max > 0;
    if (persist && !c->l2_persist_on) {
      HISPMV_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)c->l2_persist_max));
      c->l2_persist_on = true;
    }
    for (size_t k = 0; k < m->slabs.size(); ++k) {
      const bool first = k == 0, last = k + 1 == m->slabs.size();
      if (persist) {
        const int64_t lo = (int64_t)k * m->slab_cols;
        const int64_t bytes = std::min<int64_t>(std::min<int64_t>(m->slab_cols, m->cols - lo) * 4, c->l2_window_max);
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = const_cast<float*>(d_x + lo);
        av.accessPolicyWindow.num_bytes = (size_t)bytes;
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2_persist_max / (double)bytes);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        HISPMV_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av));
      }
      int st = run_matrix(c, m->slabs[k], d_x, first ? d_bias : d_y, d_y, alpha, first ? beta : 1.0f, last ? relu : 0, s,
                          lane);
      if (st != HISPMV_OK) return st;
    }
    if (persist) {
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.num_bytes = 0;
      HISPMV_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    return HISPMV_OK;
  }
  CsrDev A;
  A.rows = m->local_rows();
  A.cols = m->cols;
  A.nnz = m->nnz;
  A.row_ptr = m->d_row_ptr;
  A.col = m->d_col;
  A.val = m->d_val;
  switch (m->kernel) {
    case HISPMV_KERNEL_EMPTY: return launch_empty(A.rows, d_y, ep, s);
    case HISPMV_KERNEL_CSR_SCALAR: return launch_csr_scalar(A, d_x, d_y, ep, s);
    case HISPMV_KERNEL_CSR_VECTOR: return launch_csr_vector(A, m->lanes, d_x, d_y, ep, s);
    case HISPMV_KERNEL_MERGE: {
      MergePlan P;
      P.tile_items = m->tile_items;
      P.num_tiles = m->num_tiles;
      P.tile_row = m->d_tile_row;
      P.tile_nnz = m->d_tile_nnz;
      P.carry = m->d_carry + (size_t)lane * m->num_tiles;
      return launch_merge(A, P, d_x, d_y, ep, s);
    }
    case HISPMV_KERNEL_ADAPTIVE:
    case HISPMV_KERNEL_ROWSTAGE: {
      AdaptivePlan P;
      P.stream_items = m->tile_items;
      P.long_threshold = m->long_threshold;
      P.chunk_nnz = m->chunk_nnz;
      P.num_tiles = m->num_tiles;
      P.tile_row = m->d_tile_row;
      P.tile_chunk = m->d_tile_chunk;
      P.desc = m->d_desc;
      P.carry = m->d_carry + (size_t)lane * m->num_tiles;
      P.counter = m->d_counter + (size_t)lane * m->num_tiles;
      P.hot_cols = m->hot_cols;
      P.tile_begin = tile_begin;
      P.tile_count = tile_count;
      P.ahead = m->ahead;
      P.ahead_all = m->ahead_all;
      P.sched = m->d_counter + (size_t)std::max<int64_t>(m->num_tiles, 1) * 2 + 4 * (size_t)lane;
      if (m->kernel == HISPMV_KERNEL_ROWSTAGE)
        return launch_rowstage(A, P, m->lanes, m->rowstage_threads, d_x, d_y, ep, s);
      if (m->warptile) return launch_warptile(A, P, d_x, d_y, ep, s);
      if (m->pipeline) return launch_pipeline(A, P, d_x, d_y, ep, c->sm_count, s);
      if (m->persistent) return launch_adaptive_persistent(A, P, d_x, d_y, ep, c->sm_count, s);
      return launch_adaptive(A, P, m->adaptive_threads, d_x, d_y, ep, s);
    }
    case HISPMV_KERNEL_BLOCKED: {
      if (!m->pb.d_part[lane]) {  // the second stream lane of linear() gets its own partial-sum buffer on first use
        HISPMV_CUDA(cudaMalloc((void**)&m->pb.d_part[lane], ((size_t)m->pb.num_pieces + 64) * 4));
      }
      PbPlan P;
      P.slab_cols = m->pb.slab_cols;
      P.num_slabs = m->pb.num_slabs;
      P.padded_nnz = m->pb.padded_nnz;
      P.slab_ptr = m->pb.d_slab_ptr;
      P.val = m->pb.d_val;
      P.lcol = m->pb.d_lcol;
      P.flags = m->pb.d_flags;
      P.group_base = m->pb.d_group_base;
      P.perm = m->pb.d_perm;
      P.prow_ptr = m->pb.d_prow_ptr;
      P.part = m->pb.d_part[lane];
      P.num_pieces = m->pb.num_pieces;
      P.num_panels = m->num_tiles;
      P.desc = m->d_desc;
      P.panel_seg = m->pb.d_panel_seg;
      P.seg = m->pb.d_seg;
      P.max_panel_segs = m->pb.max_panel_segs;
      P.panel_chunk = m->pb.d_panel_chunk;
      P.chunk = m->pb.d_chunk;
      P.seg_copy = m->pb.d_seg_copy;
      P.perm2 = m->pb.d_perm2;
      P.chunk_src = m->pb.d_chunk_src;
      P.panel_aux = m->pb.d_panel_aux;
      P.end_bits = m->pb.d_end_bits;
      P.reduce_words = m->pb.reduce_words;
      P.work = m->pb.d_work;
      P.num_work = m->pb.num_work;
      P.work_begin = (work_part == 1 || work_part == 2) && m->pb.head_cols > 0 ? work_part * m->pb.num_work : 0;
      P.cap_words = m->tile_items + m->long_threshold;  // no STREAM panel holds more slots than this
      P.panel_begin = tile_begin;
      P.panel_count = tile_count;
      P.carry = m->d_carry + (size_t)lane * m->num_tiles;
      P.counter = m->d_counter + (size_t)lane * m->num_tiles;
      if (phases & 1) {
        int st = launch_pb_expand(P, m->cols, d_x, s);
        if (st != HISPMV_OK) return st;
      }
      if (phases & 2) return launch_pb_reduce(A, P, d_y, ep, s);
      return HISPMV_OK;
    }
    default: set_error("run: matrix has no plan"); return HISPMV_ERR_STATE;
  }
}

constexpr int kBatchMax = 8;              // vectors per pass (batch.cu)
constexpr int kBatchMaxRowNnz = 1 << 16;  // a sub-warp walks a whole row: keep the longest row bounded

int ensure_batch_xi(hispmv_ctx* c, int64_t n) {
  if (n > c->cap_xi) {
    c->cap_xi = 0;
    for (int l = 0; l < 2; ++l) {
      cudaFree(c->d_xi[l]);
      c->d_xi[l] = nullptr;
    }
    for (int l = 0; l < 2; ++l) {
      int st = check_cuda(cudaMalloc((void**)&c->d_xi[l], (size_t)n * 4), "cudaMalloc(batch x)", __FILE__, __LINE__);
      if (st != HISPMV_OK) return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st;
    }
    c->cap_xi = n;
  }
  return HISPMV_OK;
}

int ensure_batch_host_staging(hispmv_ctx* c, int64_t n_x, int64_t n_y) {
  if (n_x > c->cap_xb) {
    c->cap_xb = 0;
    cudaFree(c->d_xb);
    c->d_xb = nullptr;
    int st = check_cuda(cudaMalloc((void**)&c->d_xb, (size_t)n_x * 4), "cudaMalloc(batch x staging)", __FILE__, __LINE__);
    if (st != HISPMV_OK) return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st;
    c->cap_xb = n_x;
  }
  if (n_y > c->cap_yb) {
    c->cap_yb = 0;
    cudaFree(c->d_yb);
    c->d_yb = nullptr;
    int st = check_cuda(cudaMalloc((void**)&c->d_yb, (size_t)n_y * 4), "cudaMalloc(batch y staging)", __FILE__, __LINE__);
    if (st != HISPMV_OK) return st == HISPMV_FULL ? HISPMV_ERR_CUDA : st;
    c->cap_yb = n_y;
  }
  return HISPMV_OK;
}

// Every non-empty matrix takes several vectors per pass (rows a sub-warp should not walk alone -- more than
// kBatchMaxRowNnz nonzeros -- get one CTA each; column-slab and blocked matrices use their plain CSR here);
// HISPMV_BATCH=0 turns it off.
bool batch_eligible(const Matrix* m) {
  static const bool off = getenv("HISPMV_BATCH") && atoi(getenv("HISPMV_BATCH")) == 0;
  if (!off && m->dense) return m->local_rows() > 0 && m->cols > 0;
  return !off && !m->dense && m->kernel != HISPMV_KERNEL_EMPTY && m->nnz > 0 && m->local_rows() > 0;
}

// y [nv][rows] = alpha * A x_k + beta * bias for the nv vectors x [nv][cols] (both row-major in HBM), in passes of up to
// eight vectors where the matrix allows it, vector by vector otherwise.  `lane` picks the context's scratch set.
int run_matrix_batch(hispmv_ctx* c, Matrix* m, const float* d_x, const float* d_bias, float* d_y, int64_t nv, float alpha,
                     float beta, int relu, cudaStream_t s, int lane = 0) {
  const int64_t rows = m->local_rows();
  if (nv <= 0) return HISPMV_OK;
  HISPMV_CUDA(cudaPeekAtLastError());  // a stale error must not be blamed on the launches below
  if (nv == 1 || !batch_eligible(m)) {
    for (int64_t v = 0; v < nv; ++v) {
      int st = run_matrix(c, m, d_x + v * m->cols, d_bias, d_y + v * rows, alpha, beta, relu, s, lane);
      if (st != HISPMV_OK) return st;
    }
    return HISPMV_OK;
  }
  if (beta != 0.0f && !d_bias) {
    set_error("run: bias is required when beta != 0");
    return HISPMV_ERR_ARG;
  }
  int st = ensure_batch_xi(c, m->dense ? panel_floats(m->ld) : (int64_t)kBatchMax * m->cols);
  if (st != HISPMV_OK) return st;
  if (m->dense) {
    DenseDev D;
    D.rows = (int32_t)rows;
    D.cols = m->cols;
    D.ld = m->ld;
    D.a = m->d_a;
    Epilogue epd{alpha, beta, d_bias, relu};
    for (int64_t v0 = 0; v0 < nv; v0 += kBatchMax) {
      const int g = (int)std::min<int64_t>(kBatchMax, nv - v0);
      if (g == 1) {
        st = run_matrix(c, m, d_x + v0 * m->cols, d_bias, d_y + v0 * rows, alpha, beta, relu, s, lane);
      } else {
        st = launch_interleave_panels(d_x + v0 * m->cols, g, m->cols, m->ld, c->d_xi[lane], s);
        if (st == HISPMV_OK) st = launch_gemm_lite(D, c->d_xi[lane], d_y + v0 * rows, g, epd, s);
      }
      if (st != HISPMV_OK) return st;
    }
    return HISPMV_OK;
  }
  const double mean = (double)m->nnz / (double)rows;
  // lanes per row follow the mean row length; few long rows (one warp each would not fill the SMs) get a CTA each
  int lanes = mean >= 64 ? 32 : mean >= 32 ? 16 : mean >= 16 ? 8 : mean >= 8 ? 4 : 2;
  if (mean >= 512 && rows * 32 < (int64_t)c->sm_count * 1024) lanes = 256;
  CsrDev A;
  A.rows = (int32_t)rows;
  A.cols = m->cols;
  A.nnz = m->nnz;
  A.row_ptr = m->d_row_ptr;
  A.col = m->d_col;
  A.val = m->d_val;
  Epilogue ep{alpha, beta, d_bias, relu};
  if (m->num_batch_long_rows < 0 && lanes <= 32) {  // first batch on this matrix: list the rows a sub-warp must not walk
    m->num_batch_long_rows = 0;
    if (m->stats.max_row_nnz > kBatchMaxRowNnz) {
      st = batch_long_rows_device(m->d_row_ptr, (int32_t)rows, kBatchMaxRowNnz, m->nnz / kBatchMaxRowNnz + 1,
                                  &m->d_batch_long_rows, &m->num_batch_long_rows, s);
      if (st != HISPMV_OK) return st;
    }
  }
  const int64_t n_long = lanes <= 32 ? std::max<int64_t>(m->num_batch_long_rows, 0) : 0;
  for (int64_t v0 = 0; v0 < nv; v0 += kBatchMax) {
    const int g = (int)std::min<int64_t>(kBatchMax, nv - v0);
    if (g == 1) {
      st = run_matrix(c, m, d_x + v0 * m->cols, d_bias, d_y + v0 * rows, alpha, beta, relu, s, lane);
    } else {
      st = launch_interleave(d_x + v0 * m->cols, g, m->cols, m->cols, c->d_xi[lane], s);
      if (st == HISPMV_OK)
        st = launch_spmm_csr(A, lanes, m->d_batch_long_rows, n_long, kBatchMaxRowNnz, c->d_xi[lane], d_y + v0 * rows, g,
                             ep, s);
    }
    if (st != HISPMV_OK) return st;
  }
  return HISPMV_OK;
}

Matrix* get_matrix(hispmv_ctx* c, int64_t idx) {
  if (!c) {
    set_error("null context");
    return nullptr;
  }
  const std::vector<Matrix*>& mats = c->kids.empty() ? c->mats : c->kids[0]->mats;  // multi-GPU: child 0's block
  if (idx < 0 || idx >= (int64_t)mats.size()) {
    set_error("Matrix idx out of range");
    return nullptr;
  }
  return mats[(size_t)idx];
}

}  // namespace

extern "C" {

const char* hispmv_last_error(void) { return g_last_error.c_str(); }
int hispmv_version(void) { return 200 + (experimental_kernels_built() ? 1 : 0); }  // odd: research kernels built in

int hispmv_create(hispmv_ctx** out, int device_id, int flags) {
  if (!out) return HISPMV_ERR_ARG;
  *out = nullptr;
  int n = 0;
  HISPMV_CUDA(cudaGetDeviceCount(&n));
  if (device_id < 0 || device_id >= n) {
    set_error("hispmv_create: no such CUDA device");
    return HISPMV_ERR_CUDA;
  }
  cudaDeviceProp prop;
  HISPMV_CUDA(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) {
    set_error(std::string("hispmv_create: device '") + prop.name +
              "' is not sm_100 (B200); this library ships sm_100a code only and has no fallback");
    return HISPMV_ERR_CUDA;
  }
  DeviceGuard g(device_id);
  hispmv_ctx* c = new hispmv_ctx();
  c->device = device_id;
  c->flags = flags;
  c->sm_count = prop.multiProcessorCount;
  c->l2_persist_max = prop.persistingL2CacheMaxSize;
  c->l2_window_max = prop.accessPolicyMaxWindowSize;
  int st = check_cuda(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking), "stream", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaEventCreateWithFlags(&c->ev_bias, cudaEventDisableTiming), "event", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = check_cuda(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking), "stream", __FILE__, __LINE__);
  for (auto& e : c->ev_pipe)
    if (st == HISPMV_OK) st = check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event", __FILE__, __LINE__);
  if (st != HISPMV_OK) {
    delete c;
    return st;
  }
  *out = c;
  return HISPMV_OK;
}

int hispmv_create_multi(hispmv_ctx** out, int first_device, int n_gpus, int flags) {
  if (!out) return HISPMV_ERR_ARG;
  *out = nullptr;
  if (n_gpus < 1) {
    set_error("hispmv_create_multi: n_gpus must be at least 1");
    return HISPMV_ERR_ARG;
  }
  hispmv_ctx* parent = nullptr;
  int st = hispmv_create(&parent, first_device, flags);
  if (st != HISPMV_OK) return st;
  // HISPMV_MULTI_WRAP=1 (tests on a box with fewer GPUs): child k sits on device (first_device + k) mod device count
  int n_dev = 1;
  cudaGetDeviceCount(&n_dev);
  const bool wrap = getenv("HISPMV_MULTI_WRAP") && atoi(getenv("HISPMV_MULTI_WRAP")) != 0 && n_dev > 0;
  for (int k = 0; k < n_gpus && st == HISPMV_OK; ++k) {
    hispmv_ctx* kid = nullptr;
    st = hispmv_create(&kid, wrap ? (first_device + k) % n_dev : first_device + k, flags);
    if (st == HISPMV_OK) {
      parent->kids.push_back(kid);
      st = hispmv_set_shard(kid, k, n_gpus);
    }
  }
  if (st != HISPMV_OK) {
    const std::string keep = g_last_error;
    hispmv_destroy(parent);
    g_last_error = keep;
    return st;
  }
  *out = parent;
  return HISPMV_OK;
}

int hispmv_multi_gpus(hispmv_ctx* c) { return c ? (int)c->kids.size() : HISPMV_ERR_ARG; }
hispmv_ctx* hispmv_multi_child(hispmv_ctx* c, int k) {
  return (c && k >= 0 && k < (int)c->kids.size()) ? c->kids[(size_t)k] : nullptr;
}

void hispmv_destroy(hispmv_ctx* c) {
  if (!c) return;
  for (auto* kid : c->kids) hispmv_destroy(kid);
  c->kids.clear();
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream2);
  for (auto* m : c->mats) delete m;
  for (int l = 0; l < 2; ++l) {
    cudaFree(c->d_x[l]);
    cudaFree(c->d_y[l]);
  }
  cudaFree(c->d_bias);
  cudaFree(c->d_xi[0]);
  cudaFree(c->d_xi[1]);
  cudaFree(c->d_xb);
  cudaFree(c->d_yb);
  c->drop_small_graphs();
  if (c->h_small) cudaFreeHost(c->h_small);
  cudaFree(c->d_small);
  delete c->copier;
  for (int i = 0; i < kPinSlots; ++i) {
    if (c->h_pin[i]) cudaFreeHost(c->h_pin[i]);
    if (c->ev_pin[i]) cudaEventDestroy(c->ev_pin[i]);
  }
  cudaEventDestroy(c->ev_bias);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  for (auto& e : c->ev_pipe)
    if (e) cudaEventDestroy(e);
  delete c;
}

int hispmv_set_shard(hispmv_ctx* c, int part, int n_parts) {
  if (!c || n_parts < 1 || part < 0 || part >= n_parts) {
    set_error("set_shard: need 0 <= part < n_parts");
    return HISPMV_ERR_ARG;
  }
  if (!c->kids.empty()) return multi_refuse("set_shard");  // the children are the shards
  c->shard_part = part;
  c->shard_parts = n_parts;
  return HISPMV_OK;
}

int hispmv_shard_bounds(const int32_t* row_ptr, int32_t rows, int n_parts, int32_t* bounds) {
  if (!row_ptr || !bounds || rows < 0 || n_parts < 1) {
    set_error("shard_bounds: bad arguments");
    return HISPMV_ERR_ARG;
  }
  // host row_ptr in, host bounds out; the search itself runs on the current device like every other
  // partition artefact (one tiny kernel), so there is a single implementation to keep bit-exact.
  int32_t* d = nullptr;
  HISPMV_CUDA(cudaMalloc((void**)&d, ((size_t)rows + 1) * 4));
  int st = check_cuda(cudaMemcpy(d, row_ptr, ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice), "H2D", __FILE__, __LINE__);
  if (st == HISPMV_OK) st = shard_bounds_device(d, rows, row_ptr[rows], n_parts, bounds, 0);
  cudaFree(d);
  return st;
}

int hispmv_set_memory_limit(hispmv_ctx* c, int64_t bytes) {
  if (!c || bytes < 0) return HISPMV_ERR_ARG;
  c->mem_limit = bytes;
  for (auto* kid : c->kids) kid->mem_limit = bytes;  // per GPU
  return HISPMV_OK;
}

int hispmv_add_sparse_coo(hispmv_ctx* c, const int32_t* r, const int32_t* cc, const float* v, int64_t nnz,
                          int32_t rows, int32_t cols) {
  return add_coo_common(c, r, cc, v, nnz, rows, cols, false);
}
int hispmv_add_sparse_coo_dev(hispmv_ctx* c, const int32_t* r, const int32_t* cc, const float* v, int64_t nnz,
                              int32_t rows, int32_t cols) {
  return add_coo_common(c, r, cc, v, nnz, rows, cols, true);
}
int hispmv_add_sparse_csr(hispmv_ctx* c, const int32_t* rp, const int32_t* ci, const float* v, int32_t rows,
                          int32_t cols) {
  return add_csr_common(c, rp, ci, v, rows, cols, false);
}
int hispmv_add_sparse_csr_dev(hispmv_ctx* c, const int32_t* rp, const int32_t* ci, const float* v, int32_t rows,
                              int32_t cols) {
  return add_csr_common(c, rp, ci, v, rows, cols, true);
}
int hispmv_add_dense(hispmv_ctx* c, const float* a, int32_t rows, int32_t cols) {
  return add_dense_common(c, a, rows, cols, false);
}
int hispmv_add_dense_dev(hispmv_ctx* c, const float* a, int32_t rows, int32_t cols) {
  return add_dense_common(c, a, rows, cols, true);
}

int hispmv_commit(hispmv_ctx* c) {
  if (!c) return HISPMV_ERR_ARG;
  // Matrices are uploaded and planned when they are added (the GPU has no separate "sync BO" step), so
  // commit only fences outstanding work.  Idempotent; handles may still be added afterwards.
  if (!c->kids.empty()) {
    for (auto* kid : c->kids) {
      const int st = hispmv_commit(kid);
      if (st != HISPMV_OK) return st;
    }
    c->committed = true;
    return HISPMV_OK;
  }
  DeviceGuard g(c->device);
  HISPMV_CUDA(cudaStreamSynchronize(c->stream));
  c->committed = true;
  return HISPMV_OK;
}

int hispmv_num_matrices(hispmv_ctx* c) {
  if (!c) return HISPMV_ERR_ARG;
  return (int)(c->kids.empty() ? c->mats.size() : c->kids[0]->mats.size());
}

int hispmv_select(hispmv_ctx* c, uint32_t idx) {
  if (!get_matrix(c, idx)) return c ? HISPMV_ERR_INDEX : HISPMV_ERR_ARG;
  c->selected = (int)idx;
  for (auto* kid : c->kids) kid->selected = (int)idx;
  return HISPMV_OK;
}

int hispmv_force_kernel(hispmv_ctx* c, int idx, int kernel, int lanes) {
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!c->kids.empty())
    return for_each_kid(c, [&](hispmv_ctx* kid, int) { return hispmv_force_kernel(kid, idx, kernel, lanes); });
  if (m->dense) {
    set_error("force_kernel: dense handles always use the GeMV kernel");
    return HISPMV_ERR_ARG;
  }
  if (kernel != HISPMV_KERNEL_AUTO && kernel != HISPMV_KERNEL_CSR_SCALAR && kernel != HISPMV_KERNEL_CSR_VECTOR &&
      kernel != HISPMV_KERNEL_MERGE && kernel != HISPMV_KERNEL_ADAPTIVE && kernel != HISPMV_KERNEL_ROWSTAGE &&
      kernel != HISPMV_KERNEL_BLOCKED) {
    set_error("force_kernel: unknown kernel");
    return HISPMV_ERR_ARG;
  }
  if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) {
    set_error("force_kernel: lanes must be 0,1,2,4,8,16,32");
    return HISPMV_ERR_ARG;
  }
  DeviceGuard g(c->device);
  c->drop_small_graphs();  // captured launches of the old plan
  c->small_seen.erase(m);
  m->forced = kernel != HISPMV_KERNEL_AUTO;
  m->kernel = kernel;
  m->lanes = lanes;
  return plan_sparse(c, m);
}

int hispmv_run_dev(hispmv_ctx* c, int idx, const float* d_x, const float* d_bias, float* d_y, float alpha, float beta,
                   void* stream) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_run_dev");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  DeviceGuard g(c->device);
  return run_matrix(c, m, d_x, d_bias, d_y, alpha, beta, 0, (cudaStream_t)stream);
}

int hispmv_run_dev_phase(hispmv_ctx* c, int idx, const float* d_x, const float* d_bias, float* d_y, float alpha,
                         float beta, int phases, void* stream) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_run_dev_phase");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (phases < 1 || phases > 3) {
    set_error("run_dev_phase: phases must be 1 (products), 2 (row sums) or 3 (both)");
    return HISPMV_ERR_ARG;
  }
  DeviceGuard g(c->device);
  const bool two_pass = !m->dense && m->kernel == HISPMV_KERNEL_BLOCKED;
  if (!two_pass && !(phases & 2)) return HISPMV_OK;  // one-pass strategies do all their work in the second phase
  return run_matrix(c, m, d_x, d_bias, d_y, alpha, beta, 0, (cudaStream_t)stream, 0, 0, -1, 0, two_pass ? phases : 3);
}

int hispmv_linear_dev(hispmv_ctx* c, int idx, const float* d_x, const float* d_bias, float* d_y, int relu,
                      void* stream) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_linear_dev");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  DeviceGuard g(c->device);
  return run_matrix(c, m, d_x, d_bias, d_y, 1.0f, d_bias ? 1.0f : 0.0f, relu, (cudaStream_t)stream);
}

int hispmv_run_dev_mc(hispmv_ctx* c, int idx, const float* d_x, const float* d_bias, float* mc_y, float alpha,
                      float beta, int relu, void* stream) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_run_dev_mc");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!mc_y) {
    set_error("run_dev_mc: null multicast address");
    return HISPMV_ERR_ARG;
  }
  DeviceGuard g(c->device);
  return run_matrix(c, m, d_x, d_bias, mc_y, alpha, beta, relu, (cudaStream_t)stream, 0, 0, -1, 1);
}

int hispmv_run_dev_batch(hispmv_ctx* c, int idx, const float* d_x, const float* d_bias, float* d_y, int64_t num_vecs,
                         float alpha, float beta, int relu, void* stream) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_run_dev_batch");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (num_vecs < 0 || (num_vecs > 0 && (!d_x || !d_y))) {
    set_error("run_dev_batch: bad arguments");
    return HISPMV_ERR_ARG;
  }
  DeviceGuard g(c->device);
  return run_matrix_batch(c, m, d_x, d_bias, d_y, num_vecs, alpha, beta, relu, (cudaStream_t)stream);
}

void* hispmv_stream(hispmv_ctx* c) { return c ? (void*)c->stream : nullptr; }

int hispmv_sync(hispmv_ctx* c) {
  if (!c) return HISPMV_ERR_ARG;
  for (auto* kid : c->kids) {
    const int st = hispmv_sync(kid);
    if (st != HISPMV_OK) return st;
  }
  DeviceGuard g(c->device);
  HISPMV_CUDA(cudaStreamSynchronize(c->stream));
  HISPMV_CUDA(cudaStreamSynchronize(c->stream2));
  HISPMV_CUDA(cudaStreamSynchronize(c->stream3));
  return HISPMV_OK;
}

int hispmv_launches_per_run(hispmv_ctx* c, int idx) {
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!c->kids.empty()) {  // per step: every GPU's launches
    int total = 0;
    for (auto* kid : c->kids) total += std::max(0, hispmv_launches_per_run(kid, idx));
    return total;
  }
  if (m->local_rows() <= 0) return 0;
  if (!m->dense && m->kernel == HISPMV_KERNEL_MERGE) return m->num_tiles > 1 ? 2 : 1;
  if (!m->slabs.empty()) return (int)m->slabs.size();
  if (!m->dense && m->kernel == HISPMV_KERNEL_BLOCKED) return 2;
  return 1;
}

}  // extern "C"

// The host-buffer step behind hispmv_run / hispmv_run_xdev.  x comes either from host memory (x_host) or is already
// in HBM (d_x_ext, complete once the work queued on x_stream so far has run); bias and y are host buffers.
static int run_host_step(hispmv_ctx* c, const float* x_host, const float* d_x_ext, cudaStream_t x_stream,
                         const float* bias, float* y, float alpha, float beta) {
  if (!c) return HISPMV_ERR_ARG;
  if (c->selected < 0) {
    set_error("Run Kernel called before selecting a matrix");  // reference: assert, fpga_handle.cpp:292
    return HISPMV_ERR_STATE;
  }
  if ((!x_host && !d_x_ext) || !y || (beta != 0.0f && !bias)) {
    set_error("run: null vector");
    return HISPMV_ERR_ARG;
  }
  if (!c->kids.empty()) {
    // every GPU takes the whole x and its own block of bias and y, concurrently; rows are independent, so the blocks
    // of y are simply written side by side (the reference's handle is single-device, fpga_handle.cpp:55)
    if (d_x_ext) return multi_refuse("run_xdev");
    return for_each_kid(c, [&](hispmv_ctx* kid, int) {
      const Matrix* km = kid->mats[(size_t)kid->selected];
      return run_host_step(kid, x_host, nullptr, nullptr, bias ? bias + km->row_begin : nullptr, y + km->row_begin, alpha,
                           beta);
    });
  }
  Matrix* m = c->mats[(size_t)c->selected];
  DeviceGuard g(c->device);
  const int64_t n_y = m->local_rows();
  int st = ensure_staging(c, m->cols, n_y);
  if (st != HISPMV_OK) return st;
  cudaStream_t s = c->stream;
  const float* d_x = d_x_ext ? d_x_ext : c->d_x[0];
  if (d_x_ext) {  // order the compute stream after whatever produces x
    HISPMV_CUDA(cudaEventRecord(c->ev_pipe[0], x_stream));
    HISPMV_CUDA(cudaStreamWaitEvent(s, c->ev_pipe[0], 0));
  }
  // Large tiled matrices: the rows are cut into up to 8 ranges at tile boundaries and the ranges are pipelined over
  // three streams -- bias range i+1 goes up and y range i-1 comes down (PCIe is full duplex) while range i computes.
  // The reference overlaps its host-side fill with the running kernel the same way (fpga_handle.cpp:366-379).
  const bool tiled = !m->dense && (m->kernel == HISPMV_KERNEL_ADAPTIVE || m->kernel == HISPMV_KERNEL_ROWSTAGE ||
                                   m->kernel == HISPMV_KERNEL_BLOCKED) &&
                     !m->pipeline && !m->persistent && !m->warptile && m->slabs.empty();
  int chunks = 1;
  if (tiled && n_y >= (1 << 20) && m->num_tiles >= 64) chunks = n_y >= (1 << 22) ? 8 : 4;
  // Callers with PAGEABLE memory (plain numpy arrays through pyhispmv.FpgaHandle.run_kernel): cudaMemcpyAsync on such
  // pointers is staged by the driver one copy after the other (C2: 8.9 ms per call against 1.9 ms from pinned memory),
  // so large vectors go through the context's own pinned ring, filled and drained by a few copier threads while the
  // neighbouring slice crosses PCIe.
  const bool big = (int64_t)m->cols * 4 >= (4 << 20) || n_y * 4 >= (4 << 20);
  const bool pageable = big && ((x_host && is_pageable(x_host)) || (bias && is_pageable(bias)) || is_pageable(y));
  if (pageable) {
    cudaStream_t s_up = c->stream2;
    if (x_host && m->cols > 0) {
      st = staged_h2d(c, c->d_x[0], x_host, (size_t)m->cols * 4, s_up);
      if (st != HISPMV_OK) return st;
    }
    if (bias && n_y > 0) {
      st = staged_h2d(c, c->d_bias, bias, (size_t)n_y * 4, s_up);
      if (st != HISPMV_OK) return st;
    }
    HISPMV_CUDA(cudaEventRecord(c->ev_pipe[1], s_up));
    HISPMV_CUDA(cudaStreamWaitEvent(s, c->ev_pipe[1], 0));
    st = run_matrix(c, m, d_x, bias ? c->d_bias : nullptr, c->d_y[0], alpha, beta, 0, s);
    if (st != HISPMV_OK) return st;
    if (n_y > 0) {
      st = staged_d2h(c, y, c->d_y[0], (size_t)n_y * 4, s);
      if (st != HISPMV_OK) return st;
    }
    HISPMV_CUDA(cudaStreamSynchronize(s));
    return HISPMV_OK;
  }
  if (chunks == 1) {
    if (x_host && ((int64_t)m->cols + 2 * n_y) * 4 <= kSmallCallBytes && n_y > 0)
      return small_call(c, m, x_host, bias, y, alpha, beta);
    if (x_host && m->cols > 0)
      HISPMV_CUDA(cudaMemcpyAsync(c->d_x[0], x_host, (size_t)m->cols * 4, cudaMemcpyHostToDevice, s));
    if (bias && n_y > 0) HISPMV_CUDA(cudaMemcpyAsync(c->d_bias, bias, (size_t)n_y * 4, cudaMemcpyHostToDevice, s));
    st = run_matrix(c, m, d_x, bias ? c->d_bias : nullptr, c->d_y[0], alpha, beta, 0, s);
    if (st != HISPMV_OK) return st;
    if (n_y > 0) HISPMV_CUDA(cudaMemcpyAsync(y, c->d_y[0], (size_t)n_y * 4, cudaMemcpyDeviceToHost, s));
    HISPMV_CUDA(cudaStreamSynchronize(s));
    return HISPMV_OK;
  }
  // range boundaries: tile indices where a new row starts (never between two chunks of a split row)
  // Ranges are sized by ROWS (what the copies move), equal shares.  Measured on C2 (profiles/r1_e2e_shares.txt):
  // 8 to 16 equal ranges all take 1.92-1.95 ms; shares that shrink towards the end, or fewer and larger ranges, are
  // 2-8 % slower because the first y copy starts later.  With both directions busy the link moves ~37 GB/s each way
  // (HISPMV_RUN_TRACE=1 prints the timeline), against 55 GB/s one way: that, not the kernel, is the floor here.
  double share[kMaxRunChunks];
  for (int i = 0; i < chunks; ++i) share[i] = 1.0 / chunks;
  if (const char* e = getenv("HISPMV_RUN_SHARES")) {  // development sweeps: "f0,f1,...": 2..8 positive shares
    double f[kMaxRunChunks];
    int n = 0;
    const char* q = e;
    while (n < kMaxRunChunks) {
      char* end = nullptr;
      const double v = strtod(q, &end);
      if (end == q || !(v > 0)) break;
      f[n++] = v;
      if (*end != ',') break;
      q = end + 1;
    }
    if (n >= 2) {
      double sum = 0;
      for (int i = 0; i < n; ++i) sum += f[i];
      chunks = n;
      for (int i = 0; i < n; ++i) share[i] = f[i] / sum;
    }
  }
  int64_t tb[kMaxRunChunks + 1];
  tb[0] = 0;
  tb[chunks] = m->num_tiles;
  double cum = 0;
  for (int i = 1; i < chunks; ++i) {
    cum += share[i - 1];
    const int32_t want = (int32_t)std::min<double>((double)n_y, cum * (double)n_y);
    int64_t t = std::lower_bound(m->h_tile_row.begin(), m->h_tile_row.begin() + m->num_tiles, want) -
                m->h_tile_row.begin();
    t = std::max(t, tb[i - 1]);
    while (t < m->num_tiles && m->h_tile_chunk[(size_t)t] > 0) ++t;
    tb[i] = t;
  }
  cudaStream_t s_up = c->stream2, s_down = c->stream3;
  // HISPMV_RUN_TRACE=1: device-side timeline of one call (development; events are created per call)
  static const bool trace = getenv("HISPMV_RUN_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  std::vector<std::string> tname;
  auto mark = [&](cudaStream_t on, const std::string& what) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, on);
    tev.push_back(e);
    tname.push_back(what);
  };
  mark(s_up, "start");
  const bool two_pass = m->kernel == HISPMV_KERNEL_BLOCKED;
  // BLOCKED with x in host memory: the HEAD slabs hold most of the entries in a small part of x (C2: two thirds of the
  // entries in the first few megabytes), so x goes up in two pieces and pass 1 over the head runs while the rest of x
  // is still crossing PCIe; pass 1 over the remaining slabs, cut into one range per SM of its own, follows when x is
  // complete.  Either way pass 1 does not wait for the first bias range.
  static const bool split_ok = !(getenv("HISPMV_RUN_SPLIT_X") && atoi(getenv("HISPMV_RUN_SPLIT_X")) == 0);
  const int64_t head = two_pass && x_host && split_ok ? std::min<int64_t>(m->pb.head_cols, m->cols) : 0;
  if (x_host) {
    cudaEvent_t ev_x = c->ev_pipe[0];
    if (head > 0 && head < m->cols) {
      cudaEvent_t ev_h = c->ev_pipe[1 + 2 * kMaxRunChunks];
      HISPMV_CUDA(cudaMemcpyAsync(c->d_x[0], x_host, (size_t)head * 4, cudaMemcpyHostToDevice, s_up));
      HISPMV_CUDA(cudaEventRecord(ev_h, s_up));
      HISPMV_CUDA(cudaStreamWaitEvent(s, ev_h, 0));
      mark(s_up, "x head up");
      st = run_matrix(c, m, d_x, bias ? c->d_bias : nullptr, c->d_y[0], alpha, beta, 0, s, 0, 0, -1, 0, 1, 1);
      if (st != HISPMV_OK) return st;
      mark(s, "pass 1 head");
      HISPMV_CUDA(cudaMemcpyAsync(c->d_x[0] + head, x_host + head, (size_t)(m->cols - head) * 4, cudaMemcpyHostToDevice, s_up));
    } else {
      HISPMV_CUDA(cudaMemcpyAsync(c->d_x[0], x_host, (size_t)m->cols * 4, cudaMemcpyHostToDevice, s_up));
    }
    HISPMV_CUDA(cudaEventRecord(ev_x, s_up));
    HISPMV_CUDA(cudaStreamWaitEvent(s, ev_x, 0));
    mark(s_up, "x up");
  }
  if (two_pass) {  // pass 1 needs only x: it starts as soon as x is up, under the upload of the first bias range
    st = run_matrix(c, m, d_x, bias ? c->d_bias : nullptr, c->d_y[0], alpha, beta, 0, s, 0, 0, -1, 0, 1,
                    head > 0 && head < m->cols ? 2 : 0);  // (pass 1 never reads bias: only the pointer check sees it)
    if (st != HISPMV_OK) return st;
    mark(s, "pass 1");
  }
  for (int i = 0; i < chunks; ++i) {
    const int64_t r0 = m->h_tile_row[(size_t)tb[i]], r1 = m->h_tile_row[(size_t)tb[i + 1]];
    cudaEvent_t ev_b = c->ev_pipe[1 + i], ev_k = c->ev_pipe[1 + kMaxRunChunks + i];
    if (bias && r1 > r0) {
      HISPMV_CUDA(cudaMemcpyAsync(c->d_bias + r0, bias + r0, (size_t)(r1 - r0) * 4, cudaMemcpyHostToDevice, s_up));
      HISPMV_CUDA(cudaEventRecord(ev_b, s_up));
      HISPMV_CUDA(cudaStreamWaitEvent(s, ev_b, 0));
      mark(s_up, "bias up " + std::to_string(i));
    }
    // BLOCKED: pass 1 ran once, ahead of the first range; pass 2 follows range by range
    st = run_matrix(c, m, d_x, bias ? c->d_bias : nullptr, c->d_y[0], alpha, beta, 0, s, 0, tb[i],
                    tb[i + 1] - tb[i], 0, two_pass ? 2 : 3);
    if (st != HISPMV_OK) return st;
    HISPMV_CUDA(cudaEventRecord(ev_k, s));
    HISPMV_CUDA(cudaStreamWaitEvent(s_down, ev_k, 0));
    mark(s, "kernel " + std::to_string(i));
    if (r1 > r0)
      HISPMV_CUDA(cudaMemcpyAsync(y + r0, c->d_y[0] + r0, (size_t)(r1 - r0) * 4, cudaMemcpyDeviceToHost, s_down));
    mark(s_down, "y down " + std::to_string(i) + " (" + std::to_string((r1 - r0) * 4 / 1000) + " KB)");
  }
  HISPMV_CUDA(cudaStreamSynchronize(s_down));
  HISPMV_CUDA(cudaStreamSynchronize(s));
  HISPMV_CUDA(cudaStreamSynchronize(s_up));
  if (trace) {
    for (size_t i = 1; i < tev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[0], tev[i]);
      fprintf(stderr, "[hispmv_run] %-24s done at %.3f ms\n", tname[i].c_str(), ms);
    }
    for (auto e : tev) cudaEventDestroy(e);
  }
  return HISPMV_OK;
}

extern "C" {

int hispmv_run(hispmv_ctx* c, const float* x, const float* bias, float* y, float alpha, float beta) {
  return run_host_step(c, x, nullptr, nullptr, bias, y, alpha, beta);
}

int hispmv_run_xdev(hispmv_ctx* c, const float* d_x, void* x_stream, const float* bias, float* y, float alpha,
                    float beta) {
  if (!d_x) {
    set_error("run_xdev: null x");
    return HISPMV_ERR_ARG;
  }
  return run_host_step(c, nullptr, d_x, (cudaStream_t)x_stream, bias, y, alpha, beta);
}

int hispmv_linear(hispmv_ctx* c, int idx, const float* x, int64_t x_len, const float* bias, float* y_out) {
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!x || !y_out || !bias) {
    set_error("linear: null vector");
    return HISPMV_ERR_ARG;
  }
  if (m->cols <= 0) {
    set_error("linear: matrix has no columns");
    return HISPMV_ERR_ARG;
  }
  if (!c->kids.empty()) {
    const int64_t nv = x_len / m->cols;
    const int64_t rows = m->rows;
    if (nv == 1)  // one vector: the blocks of y sit side by side
      return for_each_kid(c, [&](hispmv_ctx* kid, int) {
        const Matrix* km = kid->mats[(size_t)idx];
        return hispmv_linear(kid, idx, x, x_len, bias + km->row_begin, y_out + km->row_begin);
      });
    return for_each_kid(c, [&](hispmv_ctx* kid, int) {  // several: [vector][block] per GPU, interleaved into [vector][row]
      const Matrix* km = kid->mats[(size_t)idx];
      const int64_t nl = km->local_rows();
      std::vector<float> tmp((size_t)std::max<int64_t>(1, nv * nl));
      const int st = hispmv_linear(kid, idx, x, x_len, bias + km->row_begin, tmp.data());
      if (st != HISPMV_OK) return st;
      for (int64_t v = 0; v < nv; ++v)
        if (nl > 0) memcpy(y_out + v * rows + km->row_begin, tmp.data() + v * nl, (size_t)nl * 4);
      return (int)HISPMV_OK;
    });
  }
  DeviceGuard g(c->device);
  const int64_t num_vecs = x_len / m->cols;  // reference: integer division, remainder ignored (fpga_handle.cpp:336)
  const int64_t n_y = m->local_rows();
  if (num_vecs == 1 && n_y > 0 && ((int64_t)m->cols + 2 * n_y) * 4 <= kSmallCallBytes)
    return small_call(c, m, x, bias, y_out, 1.0f, 1.0f);  // batch 1, a layer of model_test: one graph launch
  int st = ensure_staging(c, m->cols, n_y);
  if (st != HISPMV_OK) return st;
  cudaStream_t lanes[2] = {c->stream, c->stream2};
  if (n_y > 0) HISPMV_CUDA(cudaMemcpyAsync(c->d_bias, bias, (size_t)n_y * 4, cudaMemcpyHostToDevice, lanes[0]));
  HISPMV_CUDA(cudaEventRecord(c->ev_bias, lanes[0]));
  HISPMV_CUDA(cudaStreamWaitEvent(lanes[1], c->ev_bias, 0));
  HISPMV_CUDA(cudaPeekAtLastError());
  if (num_vecs >= 2 && batch_eligible(m)) {
    // Several vectors: groups of up to eight share one pass over the matrix (batch.cu).  Groups alternate between the two
    // stream lanes, each with its own half of the staging, so the copies of one group overlap the pass of the other.
    st = ensure_batch_host_staging(c, 2 * (int64_t)kBatchMax * m->cols, 2 * (int64_t)kBatchMax * n_y);
    if (st != HISPMV_OK) return st;
    int l = 0;
    for (int64_t v0 = 0; v0 < num_vecs; v0 += kBatchMax, l ^= 1) {
      const int64_t g = std::min<int64_t>(kBatchMax, num_vecs - v0);
      float* xb = c->d_xb + (int64_t)l * kBatchMax * m->cols;
      float* yb = c->d_yb + (int64_t)l * kBatchMax * n_y;
      HISPMV_CUDA(cudaMemcpyAsync(xb, x + v0 * m->cols, (size_t)(g * m->cols) * 4, cudaMemcpyHostToDevice, lanes[l]));
      st = run_matrix_batch(c, m, xb, c->d_bias, yb, g, 1.0f, 1.0f, 0, lanes[l], l);
      if (st != HISPMV_OK) return st;
      if (n_y > 0)
        HISPMV_CUDA(cudaMemcpyAsync(y_out + v0 * n_y, yb, (size_t)(g * n_y) * 4, cudaMemcpyDeviceToHost, lanes[l]));
    }
    HISPMV_CUDA(cudaStreamSynchronize(lanes[0]));
    HISPMV_CUDA(cudaStreamSynchronize(lanes[1]));
    return HISPMV_OK;
  }
  // Vectors alternate between two stream lanes so the copy of vector v+1 overlaps the kernel of vector v
  // (the reference overlaps host fill with the FPGA run the same way, fpga_handle.cpp:366-379).
  for (int64_t v = 0; v < num_vecs; ++v) {
    const int l = (int)(v & 1);
    HISPMV_CUDA(cudaMemcpyAsync(c->d_x[l], x + v * m->cols, (size_t)m->cols * 4, cudaMemcpyHostToDevice, lanes[l]));
    st = run_matrix(c, m, c->d_x[l], c->d_bias, c->d_y[l], 1.0f, 1.0f, 0, lanes[l], l);
    if (st != HISPMV_OK) return st;
    if (n_y > 0)
      HISPMV_CUDA(cudaMemcpyAsync(y_out + v * n_y, c->d_y[l], (size_t)n_y * 4, cudaMemcpyDeviceToHost, lanes[l]));
  }
  HISPMV_CUDA(cudaStreamSynchronize(lanes[0]));
  HISPMV_CUDA(cudaStreamSynchronize(lanes[1]));
  return HISPMV_OK;
}

int hispmv_matrix_info_get(hispmv_ctx* c, int idx, hispmv_matrix_info* out) {
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!out) return HISPMV_ERR_ARG;
  if (!c->kids.empty()) {  // child 0's plan facts, the whole matrix's extent and totals
    int st = hispmv_matrix_info_get(c->kids[0], idx, out);
    if (st != HISPMV_OK) return st;
    out->row_begin = 0;
    out->row_end = out->rows;
    for (size_t k = 1; k < c->kids.size(); ++k) {
      hispmv_matrix_info o;
      st = hispmv_matrix_info_get(c->kids[k], idx, &o);
      if (st != HISPMV_OK) return st;
      out->nnz += o.nnz;
      out->num_tiles += o.num_tiles;
      out->num_split_rows += o.num_split_rows;
      out->device_bytes += o.device_bytes;
      out->empty_rows += o.empty_rows;
      out->max_row_nnz = std::max(out->max_row_nnz, o.max_row_nnz);
      for (int i = 0; i < HISPMV_HIST_BINS; ++i) out->hist[i] += o.hist[i];
    }
    return HISPMV_OK;
  }
  memset(out, 0, sizeof(*out));
  out->rows = m->rows;
  out->cols = m->cols;
  out->row_begin = m->row_begin;
  out->row_end = m->row_end;
  out->nnz = m->nnz;
  out->is_dense = m->dense;
  out->kernel = m->kernel;
  out->vector_lanes = m->lanes;
  out->tile_items = m->tile_items;
  out->num_tiles = m->num_tiles;
  out->num_split_rows = m->num_split;
  out->max_row_nnz = m->dense ? m->cols : m->stats.max_row_nnz;
  out->empty_rows = m->dense ? 0 : m->stats.empty_rows;
  if (!m->dense)
    for (int i = 0; i < HISPMV_HIST_BINS; ++i) out->hist[i] = m->stats.hist[i];
  out->device_bytes = m->device_bytes();
  out->x_window_cols = (!m->dense && m->persistent) ? m->hot_cols : 0;
  out->num_slabs = (int32_t)m->slabs.size();
  out->slab_cols = m->slab_cols;
  if (!m->slabs.empty()) {  // the slabs carry the plans: report their totals
    out->num_tiles = 0;
    out->num_split_rows = 0;
    for (auto* sm : m->slabs) {
      out->num_tiles += sm->num_tiles;
      out->num_split_rows += sm->num_split;
    }
    out->tile_items = m->slabs[0]->tile_items;
    out->long_threshold = m->slabs[0]->long_threshold;
    out->chunk_nnz = m->slabs[0]->chunk_nnz;
  }
  out->slab_runs = m->dense ? 0 : m->slab_runs;
  out->probe_near = m->dense ? 0 : m->probe.near;
  out->probe_cmp = m->dense ? 0 : m->probe.cmp;
  out->long_threshold = m->dense ? 0 : m->long_threshold;
  out->chunk_nnz = m->dense ? 0 : m->chunk_nnz;
  return HISPMV_OK;
}

int hispmv_plan_csr(hispmv_ctx* c, int idx, int32_t* row_ptr, int32_t* col_idx, float* vals) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_csr");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (m->dense) {
    set_error("plan_csr: dense handle");
    return HISPMV_ERR_ARG;
  }
  DeviceGuard g(c->device);
  if (row_ptr) HISPMV_CUDA(cudaMemcpy(row_ptr, m->d_row_ptr, ((size_t)m->local_rows() + 1) * 4, cudaMemcpyDeviceToHost));
  if (col_idx && m->nnz) HISPMV_CUDA(cudaMemcpy(col_idx, m->d_col, (size_t)m->nnz * 4, cudaMemcpyDeviceToHost));
  if (vals && m->nnz) HISPMV_CUDA(cudaMemcpy(vals, m->d_val, (size_t)m->nnz * 4, cudaMemcpyDeviceToHost));
  return HISPMV_OK;
}

int hispmv_plan_tiles(hispmv_ctx* c, int idx, int32_t* tile_row, int64_t* tile_nnz) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_tiles");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!m->d_tile_row) {
    set_error("plan_tiles: matrix is not planned for a tiled kernel");
    return HISPMV_ERR_STATE;
  }
  DeviceGuard g(c->device);
  if (tile_row) HISPMV_CUDA(cudaMemcpy(tile_row, m->d_tile_row, (size_t)(m->num_tiles + 1) * 4, cudaMemcpyDeviceToHost));
  if (tile_nnz && m->kernel == HISPMV_KERNEL_MERGE)
    HISPMV_CUDA(cudaMemcpy(tile_nnz, m->d_tile_nnz, (size_t)(m->num_tiles + 1) * 8, cudaMemcpyDeviceToHost));
  if (tile_nnz && (m->kernel == HISPMV_KERNEL_ADAPTIVE || m->kernel == HISPMV_KERNEL_ROWSTAGE ||
                   m->kernel == HISPMV_KERNEL_BLOCKED)) {
    // offset of each tile's first nonzero: row_ptr[tile_row] (+ chunk * chunk_nnz for LONG tiles)
    std::vector<int32_t> tr((size_t)m->num_tiles + 1), tc((size_t)std::max<int64_t>(m->num_tiles, 1));
    HISPMV_CUDA(cudaMemcpy(tr.data(), m->d_tile_row, (size_t)(m->num_tiles + 1) * 4, cudaMemcpyDeviceToHost));
    if (m->num_tiles)
      HISPMV_CUDA(cudaMemcpy(tc.data(), m->d_tile_chunk, (size_t)m->num_tiles * 4, cudaMemcpyDeviceToHost));
    const int32_t* offsets = m->kernel == HISPMV_KERNEL_BLOCKED ? m->pb.d_prow_ptr : m->d_row_ptr;  // pieces / nonzeros
    for (int64_t t = 0; t <= m->num_tiles; ++t) {
      int32_t rp = 0;
      HISPMV_CUDA(cudaMemcpy(&rp, offsets + tr[(size_t)t], 4, cudaMemcpyDeviceToHost));
      tile_nnz[t] = (int64_t)rp + (t < m->num_tiles && tc[(size_t)t] > 0 ? (int64_t)tc[(size_t)t] * m->chunk_nnz : 0);
    }
  }
  return HISPMV_OK;
}

int64_t hispmv_plan_slab_nnz(hispmv_ctx* c, int idx, int slab) {
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (slab < 0 || slab >= (int)m->slabs.size()) {
    set_error("plan_slab: no such slab");
    return HISPMV_ERR_ARG;
  }
  return m->slabs[(size_t)slab]->nnz;
}

int hispmv_plan_slab_csr(hispmv_ctx* c, int idx, int slab, int32_t* row_ptr, int32_t* col_idx, float* vals) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_slab_csr");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (slab < 0 || slab >= (int)m->slabs.size()) {
    set_error("plan_slab: no such slab");
    return HISPMV_ERR_ARG;
  }
  Matrix* sm = m->slabs[(size_t)slab];
  DeviceGuard g(c->device);
  if (row_ptr) HISPMV_CUDA(cudaMemcpy(row_ptr, sm->d_row_ptr, ((size_t)sm->local_rows() + 1) * 4, cudaMemcpyDeviceToHost));
  if (col_idx && sm->nnz) HISPMV_CUDA(cudaMemcpy(col_idx, sm->d_col, (size_t)sm->nnz * 4, cudaMemcpyDeviceToHost));
  if (vals && sm->nnz) HISPMV_CUDA(cudaMemcpy(vals, sm->d_val, (size_t)sm->nnz * 4, cudaMemcpyDeviceToHost));
  return HISPMV_OK;
}

int hispmv_plan_tile_chunks(hispmv_ctx* c, int idx, int32_t* chunk_out) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_tile_chunks");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if ((m->kernel != HISPMV_KERNEL_ADAPTIVE && m->kernel != HISPMV_KERNEL_ROWSTAGE &&
       m->kernel != HISPMV_KERNEL_BLOCKED) || !m->d_tile_chunk) {
    set_error("plan_tile_chunks: matrix is not planned for the adaptive kernel");
    return HISPMV_ERR_STATE;
  }
  DeviceGuard g(c->device);
  if (chunk_out && m->num_tiles)
    HISPMV_CUDA(cudaMemcpy(chunk_out, m->d_tile_chunk, (size_t)m->num_tiles * 4, cudaMemcpyDeviceToHost));
  return HISPMV_OK;
}

int hispmv_plan_blocked_info(hispmv_ctx* c, int idx, int64_t* out8) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_blocked_info");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (!out8) return HISPMV_ERR_ARG;
  if (m->dense || m->kernel != HISPMV_KERNEL_BLOCKED) {
    set_error("plan_blocked: matrix is not planned for the blocked strategy");
    return HISPMV_ERR_STATE;
  }
  out8[0] = m->pb.slab_cols;
  out8[1] = m->pb.num_slabs;
  out8[2] = m->pb.padded_nnz;
  out8[3] = m->pb.num_seg;
  out8[4] = m->pb.num_chunks;
  out8[5] = m->pb.num_work;
  out8[6] = m->num_tiles;
  out8[7] = m->pb.num_pieces;
  return HISPMV_OK;
}

int hispmv_plan_blocked(hispmv_ctx* c, int idx, int32_t* slab_ptr, float* vals, uint16_t* lcol, uint16_t* flags,
                        int32_t* group_base, int32_t* prow_ptr, uint16_t* perm, int32_t* panel_seg,
                        int32_t* seg_start_off, int32_t* panel_chunk, int32_t* chunk_start_count, int32_t* work) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_blocked");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (m->dense || m->kernel != HISPMV_KERNEL_BLOCKED) {
    set_error("plan_blocked: matrix is not planned for the blocked strategy");
    return HISPMV_ERR_STATE;
  }
  DeviceGuard g(c->device);
  const PbArrays& a = m->pb;
  const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
  if (slab_ptr) HISPMV_CUDA(cudaMemcpy(slab_ptr, a.d_slab_ptr, ((size_t)a.num_slabs + 1) * 4, k));
  if (vals) HISPMV_CUDA(cudaMemcpy(vals, a.d_val, (size_t)a.padded_nnz * 4, k));
  if (lcol) HISPMV_CUDA(cudaMemcpy(lcol, a.d_lcol, (size_t)a.padded_nnz * 2, k));
  if (flags) HISPMV_CUDA(cudaMemcpy(flags, a.d_flags, (size_t)a.padded_nnz / 8, k));
  if (group_base) HISPMV_CUDA(cudaMemcpy(group_base, a.d_group_base, ((size_t)a.padded_nnz / kPbGroup + 1) * 4, k));
  if (prow_ptr) HISPMV_CUDA(cudaMemcpy(prow_ptr, a.d_prow_ptr, ((size_t)m->local_rows() + 1) * 4, k));
  if (perm) HISPMV_CUDA(cudaMemcpy(perm, a.d_perm, (size_t)a.num_pieces * 2, k));
  if (panel_seg) HISPMV_CUDA(cudaMemcpy(panel_seg, a.d_panel_seg, ((size_t)m->num_tiles + 1) * 4, k));
  if (seg_start_off) HISPMV_CUDA(cudaMemcpy(seg_start_off, a.d_seg, (size_t)a.num_seg * 8, k));
  if (panel_chunk) HISPMV_CUDA(cudaMemcpy(panel_chunk, a.d_panel_chunk, ((size_t)m->num_tiles + 1) * 4, k));
  if (chunk_start_count) HISPMV_CUDA(cudaMemcpy(chunk_start_count, a.d_chunk, (size_t)a.num_chunks * 8, k));
  if (work) HISPMV_CUDA(cudaMemcpy(work, a.d_work, (size_t)a.num_work * 8, k));
  return HISPMV_OK;
}

int hispmv_plan_blocked_stage(hispmv_ctx* c, int idx, int64_t* out4, int32_t* seg_copy, uint16_t* perm2,
                              int32_t* panel_aux, uint32_t* end_bits, int32_t* chunk_src) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_blocked_stage");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  if (m->dense || m->kernel != HISPMV_KERNEL_BLOCKED) {
    set_error("plan_blocked: matrix is not planned for the blocked strategy");
    return HISPMV_ERR_STATE;
  }
  DeviceGuard g(c->device);
  const PbArrays& a = m->pb;
  const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
  if (out4) {
    out4[0] = a.stage_total;
    out4[1] = a.bit_words;
    out4[2] = a.reduce_words;
    out4[3] = a.num_seg;
  }
  if (seg_copy) HISPMV_CUDA(cudaMemcpy(seg_copy, a.d_seg_copy, (size_t)a.num_seg * 8, k));
  if (perm2) HISPMV_CUDA(cudaMemcpy(perm2, a.d_perm2, (size_t)a.stage_total * 2, k));
  if (chunk_src) HISPMV_CUDA(cudaMemcpy(chunk_src, a.d_chunk_src, (size_t)a.stage_total, k));
  if (panel_aux) HISPMV_CUDA(cudaMemcpy(panel_aux, a.d_panel_aux, ((size_t)m->num_tiles + 1) * 8, k));
  if (end_bits) HISPMV_CUDA(cudaMemcpy(end_bits, a.d_end_bits, (size_t)a.bit_words * 4, k));
  return HISPMV_OK;
}

int hispmv_plan_split_rows(hispmv_ctx* c, int idx, int32_t* rows_out) {
  if (c && !c->kids.empty()) return multi_refuse("hispmv_plan_split_rows");
  Matrix* m = get_matrix(c, idx);
  if (!m) return HISPMV_ERR_INDEX;
  DeviceGuard g(c->device);
  if (rows_out && m->num_split > 0)
    HISPMV_CUDA(cudaMemcpy(rows_out, m->d_split_rows, (size_t)m->num_split * 4, cudaMemcpyDeviceToHost));
  return HISPMV_OK;
}

// Matrix Market reader with the semantics of HiSpmvHandle::loadMtx (common/src/spmv-helper.cpp:34-136):
// banner "%%MatrixMarket matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}",
// comment lines skipped, 1-based indices, pattern entries get 1.0, explicit zeros are dropped, symmetric
// and skew-symmetric files are expanded with (c, r, +-v) for off-diagonal entries.
int hispmv_parse_mtx(const char* path, int32_t* rows_out, int32_t* cols_out, int64_t* nnz_out, int32_t** r_out,
                     int32_t** c_out, float** v_out) {
  if (!path || !rows_out || !cols_out || !nnz_out || !r_out || !c_out || !v_out) return HISPMV_ERR_ARG;
  *r_out = nullptr;
  *c_out = nullptr;
  *v_out = nullptr;
  *nnz_out = 0;
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error(std::string("Error: Unable to open file ") + path);
    return HISPMV_ERR_IO;
  }
  std::string data;
  {
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(sz > 0 ? (size_t)sz : 0);
    if (sz > 0 && fread(&data[0], 1, (size_t)sz, f) != (size_t)sz) {
      fclose(f);
      set_error("Error: short read");
      return HISPMV_ERR_IO;
    }
    fclose(f);
  }
  const char* p = data.c_str();
  const char* end = p + data.size();
  auto next_line = [&](std::string& out) -> bool {
    if (p >= end) return false;
    const char* q = (const char*)memchr(p, '\n', end - p);
    if (!q) q = end;
    out.assign(p, q - p);
    p = q < end ? q + 1 : end;
    return true;
  };
  std::string line;
  if (!next_line(line)) {
    set_error("Error: Not a valid Matrix Market file.");
    return HISPMV_ERR_IO;
  }
  std::string h[5];
  {
    std::istringstream hs(line);
    for (auto& s : h) hs >> s;
  }
  if (h[0] != "%%MatrixMarket" || h[1] != "matrix") {
    set_error("Error: Not a valid Matrix Market file.");
    return HISPMV_ERR_IO;
  }
  if (h[2] != "coordinate") {
    set_error("Error: Only sparse matrices in 'coordinate' format are supported.");
    return HISPMV_ERR_IO;
  }
  const bool pattern = h[3] == "pattern";
  if (h[3] != "real" && h[3] != "integer" && !pattern) {
    set_error("Error: Unsupported data type.");
    return HISPMV_ERR_IO;
  }
  const bool symm = h[4] == "symmetric", skew = h[4] == "skew-symmetric";
  if (h[4] != "general" && !symm && !skew) {
    set_error("Error: Unsupported symmetry type. Only 'general', 'symmetric', and 'skew-symmetric' are supported.");
    return HISPMV_ERR_IO;
  }
  do {
    if (!next_line(line)) {
      set_error("Error: missing size line");
      return HISPMV_ERR_IO;
    }
  } while (!line.empty() && line[0] == '%');
  long long rows = 0, cols = 0, nnz_decl = 0;
  if (sscanf(line.c_str(), "%lld %lld %lld", &rows, &cols, &nnz_decl) != 3 || rows < 0 || cols < 0 ||
      rows > INT32_MAX || cols > INT32_MAX) {
    set_error("Error: bad size line");
    return HISPMV_ERR_IO;
  }
  // The entry lines are cut into one chunk per thread at line boundaries; every thread parses its chunk into its own
  // arrays and the chunks are concatenated in file order, so the COO comes out exactly as a serial reader produces it.
  struct Chunk {
    std::vector<int32_t> r, c;
    std::vector<float> v;
    bool stopped = false;  // a line that does not parse ends the file (as in the reference readers)
  };
  const size_t body = (size_t)(end - p);
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(nt ? nt : 1u, 16u));
  nt = (unsigned)std::max<size_t>(1, std::min<size_t>(nt, body / (1u << 20) + 1));
  if (const char* e = getenv("HISPMV_MTX_THREADS")) nt = (unsigned)std::max(1, std::min(64, atoi(e)));
  std::vector<const char*> cut(nt + 1);
  cut[0] = p;
  cut[nt] = end;
  for (unsigned i = 1; i < nt; ++i) {
    const char* q = p + body * i / nt;
    q = std::max(q, cut[i - 1]);
    const char* nl = q < end ? (const char*)memchr(q, '\n', end - q) : nullptr;
    cut[i] = nl ? nl + 1 : end;
  }
  std::vector<Chunk> chunks(nt);
  auto parse = [&](unsigned i) {
    Chunk& ck = chunks[i];
    const char* s = cut[i];
    const char* e = cut[i + 1];
    const size_t guess = (size_t)(e - s) / 12 + 16;
    ck.r.reserve(guess);
    ck.c.reserve(guess);
    ck.v.reserve(guess);
    // One entry per LINE, as the reference's getline loop reads them (common/src/spmv-helper.cpp:92-97): a number is
    // only looked for inside its own line (strtol / strtof would otherwise skip the newline and borrow tokens from the
    // next line), and a line that does not hold "row col [value]" is skipped, as oracle_load_mtx skips it.
    auto blanks = [](const char* a, const char* b) {
      while (a < b && (*a == ' ' || *a == '\t' || *a == '\r')) ++a;
      return a;
    };
    while (s < e) {
      const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
      const char* le = nl ? nl : e;          // the line is [s, le); the last one may lack its newline
      const char* next = nl ? nl + 1 : e;
      char* q;
      const char* a = blanks(s, le);
      s = next;
      if (a >= le) continue;
      const long r = strtol(a, &q, 10);
      if (q == a) continue;
      a = blanks(q, le);
      if (a >= le) continue;
      const long cc = strtol(a, &q, 10);
      if (q == a) continue;
      float v = 1.0f;
      if (!pattern) {
        a = blanks(q, le);
        if (a >= le) continue;
        v = strtof(a, &q);
        if (q == a) continue;
      }
      if (v == 0) continue;
      ck.r.push_back((int32_t)(r - 1));
      ck.c.push_back((int32_t)(cc - 1));
      ck.v.push_back(v);
      if ((symm || skew) && r != cc) {
        ck.r.push_back((int32_t)(cc - 1));
        ck.c.push_back((int32_t)(r - 1));
        ck.v.push_back(skew ? -v : v);
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < nt; ++i) pool.emplace_back(parse, i);
    parse(0);
    for (auto& t : pool) t.join();
  }
  size_t total = 0;
  unsigned used = 0;
  for (; used < nt; ++used) {
    total += chunks[used].r.size();
    if (chunks[used].stopped) { ++used; break; }
  }
  int32_t* R = (int32_t*)malloc(std::max<size_t>(total, 1) * sizeof(int32_t));
  int32_t* Cc = (int32_t*)malloc(std::max<size_t>(total, 1) * sizeof(int32_t));
  float* V = (float*)malloc(std::max<size_t>(total, 1) * sizeof(float));
  if (!R || !Cc || !V) {
    free(R);
    free(Cc);
    free(V);
    set_error("Error: out of host memory");
    return HISPMV_ERR_IO;
  }
  size_t off = 0;
  for (unsigned i = 0; i < used; ++i) {
    const size_t n = chunks[i].r.size();
    if (n) {
      memcpy(R + off, chunks[i].r.data(), n * sizeof(int32_t));
      memcpy(Cc + off, chunks[i].c.data(), n * sizeof(int32_t));
      memcpy(V + off, chunks[i].v.data(), n * sizeof(float));
    }
    off += n;
  }
  *rows_out = (int32_t)rows;
  *cols_out = (int32_t)cols;
  *nnz_out = (int64_t)total;
  *r_out = R;
  *c_out = Cc;
  *v_out = V;
  return HISPMV_OK;
}

void hispmv_parse_mtx_free(int32_t* r, int32_t* c, float* v) {
  free(r);
  free(c);
  free(v);
}

int hispmv_load_mtx(hispmv_ctx* c, const char* path) {
  if (!c || !path) return HISPMV_ERR_ARG;
  int32_t rows = 0, cols = 0;
  int64_t nnz = 0;
  int32_t *R = nullptr, *Cc = nullptr;
  float* V = nullptr;
  int st = hispmv_parse_mtx(path, &rows, &cols, &nnz, &R, &Cc, &V);
  if (st != HISPMV_OK) return st;
  st = hispmv_add_sparse_coo(c, R, Cc, V, nnz, rows, cols);
  hispmv_parse_mtx_free(R, Cc, V);
  return st;
}

}  // extern "C"
