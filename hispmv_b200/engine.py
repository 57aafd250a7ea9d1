"""Host-side mirror of the reference's accelerator handle, over the C-ABI.

`Engine` has the reference's method names and argument meaning (pyhispmv/src/pyhispmv_bindings.cpp:6-39:
create_dense_handle / create_sparse_handle / load_matrices / select_matrix / run_kernel / linear) and adds
device-resident variants that take torch CUDA tensors (`run_dev`, `linear_dev`) so benchmarks and chained
layers never cross PCIe.  The compiled `pyhispmv` module is the drop-in users import; this class is what
bench.py, the sharded runner and the tests drive.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import lib, check, check_handle


def _np(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _dptr(t, n: int = -1, what: str = "tensor", f32: bool = False, device: int = -1) -> C.c_void_p:
    """Device pointer of a torch CUDA tensor (None -> NULL); n >= 0 also checks the element count, f32 the dtype and
    device >= 0 the GPU it lives on -- a short vector would otherwise be an out-of-bounds access inside the kernel."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{what}: expected a contiguous CUDA tensor")
    if f32 and str(t.dtype) != "torch.float32":
        raise TypeError(f"{what}: expected float32, got {t.dtype}")
    if device >= 0 and t.device.index != device:
        raise ValueError(f"{what}: lives on cuda:{t.device.index}, the engine on cuda:{device}")
    if n >= 0 and t.numel() < n:
        raise ValueError(f"{what}: {t.numel()} elements, the matrix needs {n}")
    return C.c_void_p(t.data_ptr())


class Engine:
    """One GPU, many matrices.  Mirrors FpgaHandle (pyhispmv/include/fpga_handle.h:9-74)."""

    def __init__(self, device_id: int = 0, dense_overlay: bool = True, row_dist_net: bool = True,
                 shard: Optional[Sequence[int]] = None, memory_limit: int = 0, n_gpus: int = 1):
        """n_gpus > 1: one handle over the GPUs device_id .. device_id + n_gpus - 1 in this process
        (hispmv_create_multi): the host-buffer calls shard every matrix by rows and run the blocks concurrently."""
        flags = (capi.FLAG_DENSE_OVERLAY if dense_overlay else 0) | (capi.FLAG_ROW_DIST_NET if row_dist_net else 0)
        ctx = C.c_void_p()
        if n_gpus > 1:
            check(lib.hispmv_create_multi(C.byref(ctx), device_id, n_gpus, flags), "hispmv_create_multi")
        else:
            check(lib.hispmv_create(C.byref(ctx), device_id, flags), "hispmv_create")
        self.n_gpus = max(1, n_gpus)
        self._ctx = ctx
        self.device_id = device_id
        if shard is not None:
            check(lib.hispmv_set_shard(self._ctx, int(shard[0]), int(shard[1])), "hispmv_set_shard")
        if memory_limit:
            check(lib.hispmv_set_memory_limit(self._ctx, int(memory_limit)), "hispmv_set_memory_limit")

    def close(self):
        if getattr(self, "_ctx", None):
            lib.hispmv_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference surface -------------------------------------------------------------------
    def create_dense_handle(self, flattened_dense_values, rows: int, cols: int) -> int:
        a = _np(flattened_dense_values, np.float32).reshape(-1)
        if a.size < rows * cols:
            raise ValueError("flattened_dense_values is shorter than rows*cols")
        return check_handle(lib.hispmv_add_dense(self._ctx, _ptr(a), rows, cols), "create_dense_handle")

    def create_sparse_handle(self, coo_rows, coo_cols, coo_values, rows: int, cols: int) -> int:
        r, c, v = _np(coo_rows, np.int32), _np(coo_cols, np.int32), _np(coo_values, np.float32)
        if not (r.size == c.size == v.size):
            raise ValueError("coo_rows, coo_cols and coo_values differ in length")
        return check_handle(lib.hispmv_add_sparse_coo(self._ctx, _ptr(r), _ptr(c), _ptr(v), r.size, rows, cols),
                     "create_sparse_handle")

    def _shape(self, matrix_idx: int):
        """(cols, local rows) of a matrix, cached (shapes never change after creation)."""
        cache = self.__dict__.setdefault("_shapes", {})
        if matrix_idx not in cache:
            info = self.matrix_info(matrix_idx)
            cache[matrix_idx] = (info["cols"], info["row_end"] - info["row_begin"])
        return cache[matrix_idx]

    def _vecs(self, matrix_idx: int, x, bias, y, beta=1.0):
        """Device pointers of x / bias / y after checking them against the matrix's shape."""
        cols, n_y = self._shape(matrix_idx)
        dev = self.device_id
        if bias is None and beta != 0.0:
            raise ValueError("bias is required when beta != 0")
        return (_dptr(x, cols, "x", True, dev), _dptr(bias, n_y, "bias", True, dev), _dptr(y, n_y, "y", True, dev))

    def load_matrices(self) -> None:
        check(lib.hispmv_commit(self._ctx), "load_matrices")

    def select_matrix(self, matrix_idx: int) -> None:
        if matrix_idx < 0:
            raise IndexError("Matrix idx out of range")
        check(lib.hispmv_select(self._ctx, matrix_idx), "select_matrix")
        self._selected = matrix_idx

    def run_kernel(self, x, bias, y: np.ndarray, alpha: float, beta: float) -> None:
        if not (isinstance(y, np.ndarray) and y.dtype == np.float32 and y.flags.c_contiguous):
            raise TypeError("y must be a C-contiguous float32 numpy array (it is written in place)")
        xs, bs = _np(x, np.float32), _np(bias, np.float32)
        sel = getattr(self, "_selected", None)
        if sel is not None:
            cols, n_y = self._shape(sel)
            if xs.size < cols or bs.size < n_y or y.size < n_y:
                raise ValueError(f"run_kernel: the selected matrix needs x[{cols}], bias[{n_y}], y[{n_y}]; got "
                                 f"{xs.size}, {bs.size}, {y.size}")
        check(lib.hispmv_run(self._ctx, _ptr(xs), _ptr(bs), _ptr(y), alpha, beta), "run_kernel")

    def linear(self, matrix_idx: int, x, bias) -> np.ndarray:
        cols, n_y = self._shape(matrix_idx)
        xs, bs = _np(x, np.float32).reshape(-1), _np(bias, np.float32)
        if cols <= 0:
            raise ValueError("linear: the matrix has no columns")
        if bs.size < n_y:
            raise ValueError(f"linear: bias has {bs.size} entries, the matrix {n_y} rows")
        num_vecs = xs.size // cols
        y = np.empty(num_vecs * n_y, dtype=np.float32)
        check(lib.hispmv_linear(self._ctx, matrix_idx, _ptr(xs), xs.size, _ptr(bs), _ptr(y)), "linear")
        return y

    # ---- beyond the reference: CSR / device inputs, plans ---------------------------------------------
    def create_sparse_handle_csr(self, row_ptr, col_idx, values, rows: int, cols: int) -> int:
        rp, ci, v = _np(row_ptr, np.int32), _np(col_idx, np.int32), _np(values, np.float32)
        return check_handle(lib.hispmv_add_sparse_csr(self._ctx, _ptr(rp), _ptr(ci), _ptr(v), rows, cols),
                     "create_sparse_handle_csr")

    def create_sparse_handle_csr_dev(self, d_row_ptr: int, d_col: int, d_val: int, rows: int, cols: int) -> int:
        return check_handle(lib.hispmv_add_sparse_csr_dev(self._ctx, C.c_void_p(d_row_ptr), C.c_void_p(d_col),
                                                   C.c_void_p(d_val), rows, cols), "create_sparse_handle_csr_dev")

    def create_sparse_handle_coo_dev(self, rows_t, cols_t, vals_t, rows: int, cols: int) -> int:
        return check_handle(lib.hispmv_add_sparse_coo_dev(self._ctx, _dptr(rows_t), _dptr(cols_t), _dptr(vals_t),
                                                   rows_t.numel(), rows, cols), "create_sparse_handle_coo_dev")

    def create_dense_handle_dev(self, a_t, rows: int, cols: int) -> int:
        return check_handle(lib.hispmv_add_dense_dev(self._ctx, _dptr(a_t), rows, cols), "create_dense_handle_dev")

    def load_mtx(self, path: str) -> int:
        return check_handle(lib.hispmv_load_mtx(self._ctx, path.encode()), "load_mtx")

    def force_kernel(self, matrix_idx: int, kernel: int, lanes: int = 0) -> None:
        check(lib.hispmv_force_kernel(self._ctx, matrix_idx, kernel, lanes), "force_kernel")

    def run_dev(self, matrix_idx: int, x, bias, y, alpha: float = 1.0, beta: float = 0.0, stream: int = 0) -> None:
        """y = alpha*A@x + beta*bias on torch CUDA tensors, asynchronous on `stream` (a cudaStream_t as int;
        0 is the CUDA default stream, `self.stream` the engine's own)."""
        px, pb, py = self._vecs(matrix_idx, x, bias, y, beta)
        check(lib.hispmv_run_dev(self._ctx, matrix_idx, px, pb, py, alpha, beta, C.c_void_p(stream)), "run_dev")

    def run_dev_phase(self, matrix_idx: int, x, bias, y, alpha: float, beta: float, phases: int, stream: int = 0) -> None:
        """run_dev one phase at a time (hispmv_run_dev_phase): 1 = products (blocked matrices), 2 = row sums, 3 = both."""
        px, pb, py = self._vecs(matrix_idx, x, bias, y, beta if phases & 2 else 0.0)
        check(lib.hispmv_run_dev_phase(self._ctx, matrix_idx, px, pb, py, alpha, beta, phases, C.c_void_p(stream)),
              "run_dev_phase")

    def linear_dev(self, matrix_idx: int, x, bias, y, relu: bool = False, stream: int = 0) -> None:
        px, pb, py = self._vecs(matrix_idx, x, bias, y, 0.0)
        check(lib.hispmv_linear_dev(self._ctx, matrix_idx, px, pb, py, int(relu), C.c_void_p(stream)), "linear_dev")

    def run_dev_mc(self, matrix_idx: int, x, bias, mc_y: int, alpha: float = 1.0, beta: float = 0.0,
                   relu: bool = False, stream: int = 0) -> None:
        """As run_dev, but y goes through `mc_y`, the NVSwitch multicast address of this rank's first row inside a
        vector replicated on every GPU: the results land on all ranks (the all-gather of a chained layer fused into
        the producing kernel, hispmv_run_dev_mc).  Peers need a barrier before they read."""
        px, pb, _ = self._vecs(matrix_idx, x, bias, None, beta)
        check(lib.hispmv_run_dev_mc(self._ctx, matrix_idx, px, pb, C.c_void_p(mc_y), alpha, beta, int(relu),
                                    C.c_void_p(stream)), "run_dev_mc")

    def run_dev_batch(self, matrix_idx: int, x, bias, y, alpha: float = 1.0, beta: float = 0.0, relu: bool = False,
                      stream: int = 0) -> None:
        """Several right-hand sides in one pass over the matrix: x is (num_vecs, cols), y (num_vecs, local rows), both
        contiguous CUDA tensors (hispmv_run_dev_batch)."""
        cols, n_y = self._shape(matrix_idx)
        if x.dim() != 2 or y.dim() != 2 or x.shape[0] != y.shape[0] or x.shape[1] != cols or y.shape[1] != n_y:
            raise ValueError(f"expected x (num_vecs, {cols}) and y (num_vecs, {n_y})")
        if bias is None and beta != 0.0:
            raise ValueError("bias is required when beta != 0")
        dev = self.device_id
        check(lib.hispmv_run_dev_batch(self._ctx, matrix_idx, _dptr(x, -1, "x", True, dev),
                                       _dptr(bias, n_y, "bias", True, dev), _dptr(y, -1, "y", True, dev),
                                       int(x.shape[0]), alpha, beta, int(relu), C.c_void_p(stream)), "run_dev_batch")

    @property
    def stream(self) -> int:
        return int(lib.hispmv_stream(self._ctx) or 0)

    def sync(self) -> None:
        check(lib.hispmv_sync(self._ctx), "sync")

    def launches_per_run(self, matrix_idx: int) -> int:
        return check(lib.hispmv_launches_per_run(self._ctx, matrix_idx), "launches_per_run")

    def num_matrices(self) -> int:
        return lib.hispmv_num_matrices(self._ctx)

    def matrix_info(self, matrix_idx: int) -> dict:
        info = capi.MatrixInfo()
        check(lib.hispmv_matrix_info_get(self._ctx, matrix_idx, C.byref(info)), "matrix_info")
        d = {name: getattr(info, name) for name, _ in capi.MatrixInfo._fields_ if name != "hist"}
        d["hist"] = list(info.hist)
        d["is_dense"] = bool(info.is_dense)
        d["kernel_name"] = capi.KERNEL_NAMES.get(info.kernel, "?")
        return d

    def plan_csr(self, matrix_idx: int):
        info = self.matrix_info(matrix_idx)
        n = info["row_end"] - info["row_begin"]
        rp = np.empty(n + 1, np.int32)
        ci = np.empty(info["nnz"], np.int32)
        v = np.empty(info["nnz"], np.float32)
        check(lib.hispmv_plan_csr(self._ctx, matrix_idx, _ptr(rp), _ptr(ci), _ptr(v)), "plan_csr")
        return rp, ci, v

    def plan_tiles(self, matrix_idx: int):
        info = self.matrix_info(matrix_idx)
        tr = np.empty(info["num_tiles"] + 1, np.int32)
        tn = np.empty(info["num_tiles"] + 1, np.int64)
        check(lib.hispmv_plan_tiles(self._ctx, matrix_idx, _ptr(tr), _ptr(tn)), "plan_tiles")
        return tr, tn

    def plan_tile_chunks(self, matrix_idx: int) -> np.ndarray:
        info = self.matrix_info(matrix_idx)
        out = np.empty(info["num_tiles"], np.int32)
        check(lib.hispmv_plan_tile_chunks(self._ctx, matrix_idx, _ptr(out)), "plan_tile_chunks")
        return out

    def plan_slab_csr(self, matrix_idx: int, slab: int):
        info = self.matrix_info(matrix_idx)
        n = info["row_end"] - info["row_begin"]
        nnz = check(lib.hispmv_plan_slab_nnz(self._ctx, matrix_idx, slab), "plan_slab_nnz")
        rp, ci, v = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        check(lib.hispmv_plan_slab_csr(self._ctx, matrix_idx, slab, _ptr(rp), _ptr(ci), _ptr(v)), "plan_slab_csr")
        return rp, ci, v

    def plan_blocked(self, matrix_idx: int, arrays: bool = True) -> dict:
        """The blocked strategy's plan (hispmv_plan_blocked_info / hispmv_plan_blocked) as host arrays."""
        o = np.zeros(8, np.int64)
        check(lib.hispmv_plan_blocked_info(self._ctx, matrix_idx, _ptr(o)), "plan_blocked_info")
        d = {"slab_cols": int(o[0]), "num_slabs": int(o[1]), "padded_nnz": int(o[2]), "num_seg": int(o[3]),
             "num_chunks": int(o[4]), "num_work": int(o[5]), "num_panels": int(o[6]), "num_pieces": int(o[7])}
        o4 = np.zeros(4, np.int64)
        check(lib.hispmv_plan_blocked_stage(self._ctx, matrix_idx, _ptr(o4), None, None, None, None, None), "plan_blocked_stage")
        d["stage_total"], d["bit_words"], d["reduce_words"] = int(o4[0]), int(o4[1]), int(o4[2])
        if arrays:
            n_y = self._shape(matrix_idx)[1]
            d["slab_ptr"] = np.empty(d["num_slabs"] + 1, np.int32)
            d["val"] = np.empty(d["padded_nnz"], np.float32)
            d["lcol"] = np.empty(d["padded_nnz"], np.uint16)
            d["flags"] = np.empty(d["padded_nnz"] // 16, np.uint16)
            d["group_base"] = np.empty(d["padded_nnz"] // 512 + 1, np.int32)
            d["prow_ptr"] = np.empty(n_y + 1, np.int32)
            d["perm"] = np.empty(d["num_pieces"], np.uint16)
            d["panel_seg"] = np.empty(d["num_panels"] + 1, np.int32)
            d["seg"] = np.empty((d["num_seg"], 2), np.int32)
            d["panel_chunk"] = np.empty(d["num_panels"] + 1, np.int32)
            d["chunk"] = np.empty((d["num_chunks"], 2), np.int32)
            d["work"] = np.empty((d["num_work"], 2), np.int32)
            check(lib.hispmv_plan_blocked(self._ctx, matrix_idx, _ptr(d["slab_ptr"]), _ptr(d["val"]), _ptr(d["lcol"]),
                                          _ptr(d["flags"]), _ptr(d["group_base"]), _ptr(d["prow_ptr"]), _ptr(d["perm"]),
                                          _ptr(d["panel_seg"]), _ptr(d["seg"]), _ptr(d["panel_chunk"]), _ptr(d["chunk"]),
                                          _ptr(d["work"])), "plan_blocked")
            d["seg_copy"] = np.empty((d["num_seg"], 2), np.int32)
            d["perm2"] = np.empty(d["stage_total"], np.uint16)
            d["panel_aux"] = np.empty((d["num_panels"] + 1, 2), np.int32)
            d["end_bits"] = np.empty(d["bit_words"], np.uint32)
            d["chunk_src"] = np.empty(d["stage_total"] // 4, np.int32)
            check(lib.hispmv_plan_blocked_stage(self._ctx, matrix_idx, None, _ptr(d["seg_copy"]), _ptr(d["perm2"]),
                                                _ptr(d["panel_aux"]), _ptr(d["end_bits"]), _ptr(d["chunk_src"])),
                  "plan_blocked_stage")
        return d

    def plan_split_rows(self, matrix_idx: int) -> np.ndarray:
        info = self.matrix_info(matrix_idx)
        out = np.empty(info["num_split_rows"], np.int32)
        check(lib.hispmv_plan_split_rows(self._ctx, matrix_idx, _ptr(out)), "plan_split_rows")
        return out


def parse_mtx(path: str):
    """Matrix Market file -> (rows_i32, cols_i32, vals_f32, n_rows, n_cols): the multithreaded host parser behind
    load_mtx (hispmv_parse_mtx), 0-based COO in file order with loadMtx's rules (common/src/spmv-helper.cpp:34-136).
    Needs no GPU."""
    nr, nc, nnz = C.c_int32(), C.c_int32(), C.c_int64()
    r, c_, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
    check(lib.hispmv_parse_mtx(path.encode(), C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(r), C.byref(c_),
                               C.byref(v)), "parse_mtx")
    try:
        n = nnz.value
        rows = np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        cols = np.ctypeslib.as_array(C.cast(c_, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        vals = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_float)), shape=(max(n, 1),))[:n].copy()
    finally:
        lib.hispmv_parse_mtx_free(r, c_, v)
    return rows, cols, vals, nr.value, nc.value


def shard_bounds(row_ptr, n_parts: int) -> np.ndarray:
    rp = _np(row_ptr, np.int32)
    out = np.empty(n_parts + 1, np.int32)
    check(lib.hispmv_shard_bounds(_ptr(rp), rp.size - 1, n_parts, _ptr(out)), "shard_bounds")
    return out
