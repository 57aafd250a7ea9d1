"""The synthetic workloads BASELINE.json names (C1..C5), as data -- no compute.

Large matrices (C2, C4, C5) are generated in HBM by libhispmv_cuda.so's counter-hash generators
(include/hispmv_synth.h); C1 and the model_test MLP weights are small and built with numpy / torch on the
host exactly as SURVEY.md 8(d) describes.  `scale` shrinks a config for parity tests without changing its
character.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import capi
from .capi import lib, check


from .workloads import (SynthSpec, c2_powerlaw, c4_stencil, c5_uniform, reference_vectors, REF_ALPHA, REF_BETA,  # noqa: F401
                        SYNTH_POWERLAW, SYNTH_UNIFORM, SYNTH_STENCIL27)


class DeviceCSR:
    """CSR arrays living in HBM, produced by hispmv_synth_csr (raw device pointers, freed on close)."""

    def __init__(self, spec: SynthSpec, row_begin: int = 0, row_end: int | None = None):
        row_end = spec.rows if row_end is None else row_end
        p = spec.params_array()
        rp, ci, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz = C.c_int64()
        check(lib.hispmv_synth_csr(spec.kind, spec.seed, spec.rows, spec.cols, C.c_void_p(p.ctypes.data), row_begin,
                                   row_end, C.byref(rp), C.byref(ci), C.byref(v), C.byref(nnz)), "synth_csr")
        self.spec, self.row_begin, self.row_end = spec, row_begin, row_end
        self.row_ptr, self.col, self.val, self.nnz = rp.value, ci.value, v.value, nnz.value
        self.rows = row_end - row_begin

    def close(self):
        for name in ("row_ptr", "col", "val"):
            ptr = getattr(self, name, None)
            if ptr:
                lib.hispmv_synth_free(C.c_void_p(ptr))
                setattr(self, name, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_count(spec: SynthSpec, row_begin: int = 0, row_end: int | None = None) -> int:
    row_end = spec.rows if row_end is None else row_end
    p = spec.params_array()
    nnz = C.c_int64()
    check(lib.hispmv_synth_count(spec.kind, spec.seed, spec.rows, spec.cols, C.c_void_p(p.ctypes.data), row_begin,
                                 row_end, C.byref(nnz)), "synth_count")
    return nnz.value


def synth_shard_bounds(spec: SynthSpec, n_parts: int):
    p = spec.params_array()
    out = np.empty(n_parts + 1, np.int32)
    total = C.c_int64()
    check(lib.hispmv_synth_shard_bounds(spec.kind, spec.seed, spec.rows, spec.cols, C.c_void_p(p.ctypes.data), n_parts,
                                        C.c_void_p(out.ctypes.data), C.byref(total)), "synth_shard_bounds")
    return out, total.value


# --------------------------------------------------------------------------------------------------
# Small host-built configs
# --------------------------------------------------------------------------------------------------
def c1_imbalanced_coo(n: int = 65536, seed: int = 0, dense_rows: int = 4, dense_len: int = 25000,
                      target_nnz: int = 900_000):
    """C1: n x n, row lengths min(zipf(1.8), 4096) scaled to ~target_nnz, plus `dense_rows`
    rows of `dense_len` nonzeros; columns uniform without repeats inside a row; values N(0,1).
    Returned as unsorted COO (int32, int32, float32) -- the form the plugin API takes."""
    rng = np.random.default_rng(seed)
    lens = np.minimum(rng.zipf(1.8, size=n), 4096).astype(np.int64)
    lens = np.minimum(np.maximum((lens * (target_nnz / lens.sum())).astype(np.int64), 0), n)
    heavy = rng.choice(n, size=dense_rows, replace=False)
    lens[heavy] = min(dense_len, n)
    rows = np.repeat(np.arange(n, dtype=np.int32), lens)
    cols = np.empty(rows.size, dtype=np.int32)
    off = 0
    for r in range(n):
        l = int(lens[r])
        if l:
            if l * 8 < n:
                c = np.unique(rng.integers(0, n, size=l))
                while c.size < l:  # top up the few collisions
                    c = np.unique(np.concatenate([c, rng.integers(0, n, size=l - c.size)]))
            else:
                c = rng.choice(n, size=l, replace=False)
            cols[off:off + l] = c[:l]
            off += l
    vals = rng.standard_normal(rows.size).astype(np.float32)
    perm = rng.permutation(rows.size)  # COO arrives unsorted
    return rows[perm], cols[perm], vals[perm], n, n


def write_mtx(path: str, rows, cols, vals, n_rows: int, n_cols: int) -> None:
    """Matrix Market 'coordinate real general', 1-based, %.9g values (round-trips fp32)."""
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{n_rows} {n_cols} {len(vals)}\n")
        for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist()):
            f.write(f"{r + 1} {c + 1} {v:.9g}\n")
