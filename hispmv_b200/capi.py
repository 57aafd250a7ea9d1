"""ctypes view of the C-ABI in include/hispmv.h and include/hispmv_synth.h.

This is the only place Python touches libhispmv_cuda.so directly.  It is deliberately dumb: argument
types, return types, and a status check that raises with hispmv_last_error().  There is no fallback of any
kind -- if the shared library is missing the import fails, and if no B200 is visible hispmv_create fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhispmv_cuda.so")

HIST_BINS = 33

OK, FULL, ERR_ARG, ERR_INDEX, ERR_CUDA, ERR_STATE, ERR_IO = 0, -1, -2, -3, -4, -5, -6
KERNEL_AUTO, KERNEL_CSR_SCALAR, KERNEL_CSR_VECTOR, KERNEL_MERGE, KERNEL_GEMV, KERNEL_EMPTY, KERNEL_ADAPTIVE, KERNEL_ROWSTAGE, KERNEL_BLOCKED = 0, 1, 2, 3, 4, 5, 6, 7, 8
KERNEL_NAMES = {0: "auto", 1: "csr_scalar", 2: "csr_vector", 3: "merge", 4: "gemv", 5: "empty", 6: "adaptive", 7: "rowstage", 8: "blocked"}
FLAG_DENSE_OVERLAY, FLAG_ROW_DIST_NET = 1, 2
SYNTH_POWERLAW, SYNTH_UNIFORM, SYNTH_STENCIL27 = 1, 2, 3


class MatrixInfo(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
        ("nnz", C.c_int64), ("is_dense", C.c_int32), ("kernel", C.c_int32), ("vector_lanes", C.c_int32),
        ("tile_items", C.c_int32), ("num_tiles", C.c_int64), ("num_split_rows", C.c_int64),
        ("max_row_nnz", C.c_int32), ("empty_rows", C.c_int32), ("hist", C.c_int64 * HIST_BINS),
        ("device_bytes", C.c_int64), ("probe_near", C.c_int64), ("probe_cmp", C.c_int64),
        ("x_window_cols", C.c_int32), ("long_threshold", C.c_int32), ("chunk_nnz", C.c_int32),
        ("num_slabs", C.c_int32), ("slab_cols", C.c_int32), ("reserved_", C.c_int32), ("slab_runs", C.c_int64),
    ]


class HispmvError(RuntimeError):
    def __init__(self, status: int, where: str, message: str):
        super().__init__(f"{where}: {message} (status {status})")
        self.status = status


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
            "hispmv_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    p, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
    pp = C.POINTER(C.c_void_p)
    sig = {
        "hispmv_last_error": (C.c_char_p, []),
        "hispmv_version": (C.c_int, []),
        "hispmv_create": (C.c_int, [pp, C.c_int, C.c_int]),
        "hispmv_create_multi": (C.c_int, [pp, C.c_int, C.c_int, C.c_int]),
        "hispmv_multi_gpus": (C.c_int, [p]),
        "hispmv_multi_child": (C.c_void_p, [p, C.c_int]),
        "hispmv_destroy": (None, [p]),
        "hispmv_set_shard": (C.c_int, [p, C.c_int, C.c_int]),
        "hispmv_shard_bounds": (C.c_int, [p, i32, C.c_int, p]),
        "hispmv_set_memory_limit": (C.c_int, [p, i64]),
        "hispmv_add_sparse_coo": (C.c_int, [p, p, p, p, i64, i32, i32]),
        "hispmv_add_sparse_csr": (C.c_int, [p, p, p, p, i32, i32]),
        "hispmv_add_dense": (C.c_int, [p, p, i32, i32]),
        "hispmv_add_sparse_coo_dev": (C.c_int, [p, p, p, p, i64, i32, i32]),
        "hispmv_add_sparse_csr_dev": (C.c_int, [p, p, p, p, i32, i32]),
        "hispmv_add_dense_dev": (C.c_int, [p, p, i32, i32]),
        "hispmv_commit": (C.c_int, [p]),
        "hispmv_num_matrices": (C.c_int, [p]),
        "hispmv_select": (C.c_int, [p, u32]),
        "hispmv_force_kernel": (C.c_int, [p, C.c_int, C.c_int, C.c_int]),
        "hispmv_run": (C.c_int, [p, p, p, p, f32, f32]),
        "hispmv_linear": (C.c_int, [p, C.c_int, p, i64, p, p]),
        "hispmv_run_dev": (C.c_int, [p, C.c_int, p, p, p, f32, f32, p]),
        "hispmv_run_dev_phase": (C.c_int, [p, C.c_int, p, p, p, f32, f32, C.c_int, p]),
        "hispmv_linear_dev": (C.c_int, [p, C.c_int, p, p, p, C.c_int, p]),
        "hispmv_sync": (C.c_int, [p]),
        "hispmv_stream": (C.c_void_p, [p]),
        "hispmv_launches_per_run": (C.c_int, [p, C.c_int]),
        "hispmv_matrix_info_get": (C.c_int, [p, C.c_int, C.POINTER(MatrixInfo)]),
        "hispmv_plan_csr": (C.c_int, [p, C.c_int, p, p, p]),
        "hispmv_plan_tiles": (C.c_int, [p, C.c_int, p, p]),
        "hispmv_plan_split_rows": (C.c_int, [p, C.c_int, p]),
        "hispmv_plan_tile_chunks": (C.c_int, [p, C.c_int, p]),
        "hispmv_plan_blocked_info": (C.c_int, [p, C.c_int, p]),
        "hispmv_plan_blocked": (C.c_int, [p, C.c_int, p, p, p, p, p, p, p, p, p, p, p, p]),
        "hispmv_plan_blocked_stage": (C.c_int, [p, C.c_int, p, p, p, p, p, p]),
        "hispmv_plan_slab_nnz": (i64, [p, C.c_int, C.c_int]),
        "hispmv_plan_slab_csr": (C.c_int, [p, C.c_int, C.c_int, p, p, p]),
        "hispmv_load_mtx": (C.c_int, [p, C.c_char_p]),
        "hispmv_parse_mtx": (C.c_int, [C.c_char_p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), pp, pp, pp]),
        "hispmv_parse_mtx_free": (None, [p, p, p]),
        "hispmv_multicast_copy": (C.c_int, [p, p, i64, C.c_int, p]),
        "hispmv_peer_copy": (C.c_int, [p, C.c_int, p, i64, C.c_int, p]),
        "hispmv_run_xdev": (C.c_int, [p, p, p, p, p, f32, f32]),
        "hispmv_run_dev_mc": (C.c_int, [p, C.c_int, p, p, p, f32, f32, C.c_int, p]),
        "hispmv_run_dev_batch": (C.c_int, [p, C.c_int, p, p, p, i64, f32, f32, C.c_int, p]),
        "hispmv_synth_count": (C.c_int, [C.c_int, u64, i32, i32, p, i32, i32, C.POINTER(i64)]),
        "hispmv_synth_shard_bounds": (C.c_int, [C.c_int, u64, i32, i32, p, C.c_int, p, C.POINTER(i64)]),
        "hispmv_synth_csr": (C.c_int, [C.c_int, u64, i32, i32, p, i32, i32, pp, pp, pp, C.POINTER(i64)]),
        "hispmv_synth_free": (None, [p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED = [
    "hispmv_last_error", "hispmv_version", "hispmv_create", "hispmv_create_multi", "hispmv_multi_gpus", "hispmv_multi_child", "hispmv_destroy", "hispmv_set_shard",
    "hispmv_shard_bounds", "hispmv_set_memory_limit", "hispmv_add_sparse_coo", "hispmv_add_sparse_csr",
    "hispmv_add_dense", "hispmv_add_sparse_coo_dev", "hispmv_add_sparse_csr_dev", "hispmv_add_dense_dev",
    "hispmv_commit", "hispmv_num_matrices", "hispmv_select", "hispmv_force_kernel", "hispmv_run",
    "hispmv_linear", "hispmv_run_dev", "hispmv_run_dev_phase", "hispmv_linear_dev", "hispmv_sync", "hispmv_stream", "hispmv_launches_per_run",
    "hispmv_matrix_info_get", "hispmv_plan_csr", "hispmv_plan_tiles", "hispmv_plan_split_rows", "hispmv_plan_tile_chunks", "hispmv_plan_blocked_info", "hispmv_plan_blocked", "hispmv_plan_blocked_stage", "hispmv_plan_slab_nnz", "hispmv_plan_slab_csr", "hispmv_load_mtx", "hispmv_parse_mtx", "hispmv_parse_mtx_free", "hispmv_multicast_copy", "hispmv_peer_copy", "hispmv_run_xdev", "hispmv_run_dev_mc", "hispmv_run_dev_batch",
    "hispmv_synth_count", "hispmv_synth_shard_bounds", "hispmv_synth_csr", "hispmv_synth_free",
]


def last_error() -> str:
    return lib.hispmv_last_error().decode("utf-8", "replace")


def check(status: int, where: str) -> int:
    """Raise on any negative status (IndexError for a bad matrix index, as the plugin does)."""
    if status < 0:
        if status == ERR_INDEX:
            raise IndexError(f"{where}: {last_error()}")
        raise HispmvError(status, where, last_error())
    return status


def check_handle(status: int, where: str) -> int:
    """For the calls that hand out a matrix handle (create_*_handle, load_mtx): -1 = "device memory is full" is a
    return value there, exactly as in the reference (pyhispmv/src/fpga_handle.cpp:192-195,235-238); every other
    negative status raises."""
    if status == FULL:
        return status
    return check(status, where)
