// pyhispmv -- drop-in for the reference's pybind11 module (pyhispmv/src/pyhispmv_bindings.cpp:3-40,
// pyhispmv/include/fpga_handle.h:9-74): same module name, same class FpgaHandle, same method and
// keyword names, same dtypes.  The body is a thin shim over the C-ABI in include/hispmv.h; all compute
// is CUDA.  There is no CPU path: constructing a handle without a B200 raises RuntimeError.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "hispmv.h"

namespace py = pybind11;

namespace {

[[noreturn]] void fail(const char* what) {
  throw std::runtime_error(std::string(what) + ": " + hispmv_last_error());
}

using f32_array = py::array_t<float, py::array::c_style | py::array::forcecast>;
using i32_array = py::array_t<int, py::array::c_style | py::array::forcecast>;

class FpgaHandle {
 public:
  // Reference ctor: fpga_handle.cpp:40-154.  xclbin_path and the HBM channel counts describe FPGA fabric
  // and are accepted and ignored; device_id is the CUDA ordinal; dense_overlay gates create_dense_handle
  // (spmv-helper.cpp:718); row_dist_net allows heavy rows to be split across CTAs; pre_accumulator has no
  // GPU meaning (partial sums are always pre-reduced in registers).
  FpgaHandle(const std::string& xclbin_path, int device_id, int num_ch_A, int num_ch_B, int num_ch_C,
             int urams_per_pe, int fp_acc_latency, bool dense_overlay, bool pre_accumulator, bool row_dist_net) {
    (void)xclbin_path; (void)num_ch_A; (void)num_ch_B; (void)num_ch_C;
    (void)urams_per_pe; (void)fp_acc_latency; (void)pre_accumulator;
    int flags = 0;
    if (dense_overlay) flags |= HISPMV_FLAG_DENSE_OVERLAY;
    if (row_dist_net) flags |= HISPMV_FLAG_ROW_DIST_NET;
    // HISPMV_GPUS=G in the environment: the same single handle over the GPUs device_id .. device_id + G - 1 (every
    // matrix row-sharded, run_kernel / linear fan out); nothing else about the API changes
    int gpus = 1;
    if (const char* e = std::getenv("HISPMV_GPUS")) gpus = std::atoi(e);
    const int st = gpus > 1 ? hispmv_create_multi(&ctx_, device_id, gpus, flags) : hispmv_create(&ctx_, device_id, flags);
    if (st != HISPMV_OK) fail("FpgaHandle");
  }
  ~FpgaHandle() { hispmv_destroy(ctx_); }
  FpgaHandle(const FpgaHandle&) = delete;
  FpgaHandle& operator=(const FpgaHandle&) = delete;

  int createDenseMtxHandle(const f32_array& flattened_dense_values, int rows, int cols) {
    if ((int64_t)flattened_dense_values.size() < (int64_t)rows * cols)
      throw std::invalid_argument("create_dense_handle: flattened_dense_values is shorter than rows*cols");
    int idx;
    {
      const float* p = flattened_dense_values.data();
      py::gil_scoped_release nogil;
      idx = hispmv_add_dense(ctx_, p, rows, cols);
    }
    if (idx < HISPMV_FULL) fail("create_dense_handle");
    return idx;  // >= 0, or -1 when device memory is full (fpga_handle.cpp:235-238)
  }

  int createSparseMtxHandle(const i32_array& coo_rows, const i32_array& coo_cols, const f32_array& coo_values,
                            int rows, int cols) {
    const int64_t nnz = coo_rows.size();
    if (coo_cols.size() != nnz || coo_values.size() != nnz)
      throw std::invalid_argument("create_sparse_handle: coo_rows, coo_cols and coo_values differ in length");
    int idx;
    {
      const int *r = coo_rows.data(), *c = coo_cols.data();
      const float* v = coo_values.data();
      py::gil_scoped_release nogil;
      idx = hispmv_add_sparse_coo(ctx_, r, c, v, nnz, rows, cols);
    }
    if (idx < HISPMV_FULL) fail("create_sparse_handle");
    return idx;
  }

  void loadMatrices() {
    if (hispmv_commit(ctx_) != HISPMV_OK) fail("load_matrices");
  }

  void selectMatrix(uint32_t matrix_idx) {
    const int st = hispmv_select(ctx_, matrix_idx);
    if (st == HISPMV_ERR_INDEX) throw py::index_error("Matrix idx out of range");
    if (st != HISPMV_OK) fail("select_matrix");
    selected_ = (int)matrix_idx;
  }

  // y is written in place when the caller passes a C-contiguous float32 array (as the reference's
  // py::array_t<float>& does, fpga_handle.cpp:286-321).
  void runKernel(const f32_array& x, const f32_array& bias, py::array_t<float, py::array::c_style>& y, float alpha,
                 float beta) {
    hispmv_matrix_info info;
    const int sel = selected_;
    if (sel < 0) throw std::runtime_error("Run Kernel called before selecting a matrix");
    if (hispmv_matrix_info_get(ctx_, sel, &info) != HISPMV_OK) fail("run_kernel");
    const int64_t n_y = (int64_t)info.row_end - info.row_begin;
    if (x.size() < info.cols || y.size() < n_y || bias.size() < n_y)
      throw std::invalid_argument("run_kernel: x, bias or y is shorter than the selected matrix needs");
    const float *xp = x.data(), *bp = bias.data();
    float* yp = y.mutable_data();
    int st;
    {
      py::gil_scoped_release nogil;
      st = hispmv_run(ctx_, xp, bp, yp, alpha, beta);
    }
    if (st != HISPMV_OK) fail("run_kernel");
  }

  f32_array runLinear(int matrix_idx, const f32_array& x_arr, const f32_array& bias_arr) {
    hispmv_matrix_info info;
    const int st0 = hispmv_matrix_info_get(ctx_, matrix_idx, &info);
    if (st0 == HISPMV_ERR_INDEX) throw py::index_error("Matrix idx out of range");
    if (st0 != HISPMV_OK) fail("linear");
    const int64_t n_y = (int64_t)info.row_end - info.row_begin;
    if (info.cols <= 0) throw std::invalid_argument("linear: matrix has no columns");
    if (bias_arr.size() < n_y) throw std::invalid_argument("linear: bias is shorter than the matrix has rows");
    const int64_t num_vecs = x_arr.size() / info.cols;
    f32_array y(num_vecs * n_y);
    const float *xp = x_arr.data(), *bp = bias_arr.data();
    float* yp = y.mutable_data();
    int st;
    {
      py::gil_scoped_release nogil;
      st = hispmv_linear(ctx_, matrix_idx, xp, x_arr.size(), bp, yp);
    }
    if (st != HISPMV_OK) fail("linear");
    return y;
  }

  // ---- extras beyond the reference surface (plan introspection for tests / tools) ----
  py::dict matrixInfo(int idx) {
    hispmv_matrix_info i;
    const int st = hispmv_matrix_info_get(ctx_, idx, &i);
    if (st == HISPMV_ERR_INDEX) throw py::index_error("Matrix idx out of range");
    if (st != HISPMV_OK) fail("matrix_info");
    py::dict d;
    d["rows"] = i.rows; d["cols"] = i.cols; d["row_begin"] = i.row_begin; d["row_end"] = i.row_end;
    d["nnz"] = i.nnz; d["is_dense"] = (bool)i.is_dense; d["kernel"] = i.kernel; d["vector_lanes"] = i.vector_lanes;
    d["tile_items"] = i.tile_items; d["num_tiles"] = i.num_tiles; d["num_split_rows"] = i.num_split_rows;
    d["max_row_nnz"] = i.max_row_nnz; d["empty_rows"] = i.empty_rows; d["device_bytes"] = i.device_bytes;
    py::list h;
    for (int k = 0; k < HISPMV_HIST_BINS; ++k) h.append(i.hist[k]);
    d["hist"] = h;
    return d;
  }
  void setShard(int part, int n_parts) {
    if (hispmv_set_shard(ctx_, part, n_parts) != HISPMV_OK) fail("set_shard");
  }
  void setMemoryLimit(int64_t bytes) {
    if (hispmv_set_memory_limit(ctx_, bytes) != HISPMV_OK) fail("set_memory_limit");
  }
  int loadMtx(const std::string& path) {
    const int idx = hispmv_load_mtx(ctx_, path.c_str());
    if (idx < HISPMV_FULL) fail("load_mtx");
    return idx;
  }
  uintptr_t raw() const { return reinterpret_cast<uintptr_t>(ctx_); }

 private:
  hispmv_ctx* ctx_ = nullptr;
  int selected_ = -1;
};

}  // namespace

PYBIND11_MODULE(pyhispmv, m) {
  m.doc() = "Python binding for the B200 SpMV/GeMV engine (drop-in for HiSpMV's FPGA-based pyhispmv)";

  py::class_<FpgaHandle>(m, "FpgaHandle")
      .def(py::init<const std::string&, int, int, int, int, int, int, bool, bool, bool>(), py::arg("xclbin_path"),
           py::arg("device_id"), py::arg("num_ch_A"), py::arg("num_ch_B"), py::arg("num_ch_C"),
           py::arg("urams_per_pe"), py::arg("fp_acc_latency"), py::arg("dense_overlay"),
           py::arg("pre_accumulator"), py::arg("row_dist_net"))
      .def("create_dense_handle", &FpgaHandle::createDenseMtxHandle, py::arg("flattened_dense_values"),
           py::arg("rows"), py::arg("cols"), "Creates a matrix handle for a dense matrix")
      .def("create_sparse_handle", &FpgaHandle::createSparseMtxHandle, py::arg("coo_rows"), py::arg("coo_cols"),
           py::arg("coo_values"), py::arg("rows"), py::arg("cols"), "Creates a matrix handle for a sparse matrix")
      .def("load_matrices", &FpgaHandle::loadMatrices, "Loads matrices onto the device")
      .def("select_matrix", &FpgaHandle::selectMatrix, py::arg("matrix_idx"), "Select a matrix by its index")
      .def("run_kernel", &FpgaHandle::runKernel, py::arg("x"), py::arg("bias"), py::arg("y").noconvert(),
           py::arg("alpha"), py::arg("beta"),
           "Runs the SpMV kernel with the provided input/output vectors and scalars")
      .def("linear", &FpgaHandle::runLinear, py::arg("matrix_idx"), py::arg("x"), py::arg("bias"),
           "Run SpMV for given input tensors in a flattened np arrays")
      // extras
      .def("matrix_info", &FpgaHandle::matrixInfo, py::arg("matrix_idx"))
      .def("set_shard", &FpgaHandle::setShard, py::arg("part"), py::arg("n_parts"))
      .def("set_memory_limit", &FpgaHandle::setMemoryLimit, py::arg("bytes"))
      .def("load_mtx", &FpgaHandle::loadMtx, py::arg("path"))
      .def("_ctx", &FpgaHandle::raw);
}
