"""ctypes loaders for the checker libraries under oracle/ (TEST INFRASTRUCTURE).

  liboracle.so                 oracle/oracle.c, our CPU restatement
  _ref/libref_cpu.so           the reference's cpu/ path (cpu_spmv, mkl_spmv, naive_gemv, mkl_gemv, .mtx reader)
  _ref/libref_gpuhelper.so     the reference's gpu/src/spmvHelper.cpp (cooToCsr, cpuSpMV, loadMtx)
  _ref/libref_common.so        the reference's common/ host library (cpuSequential, tileAndPad, shared rows)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """(Re)build oracle/ with its Makefile.  _ref/ targets are only attempted where /root/reference exists."""
    if force or not os.path.exists(os.path.join(ORACLE_DIR, "liboracle.so")) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("libref_cpu.so", "libref_gpuhelper.so", "libref_common.so"))


_cache = {}


def oracle() -> C.CDLL:
    if "oracle" in _cache:
        return _cache["oracle"]
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    i, i64, f, d, vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_void_p
    lib.oracle_coo_to_csr.argtypes = [i, i64, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p]
    lib.oracle_coo_to_csr.restype = i
    lib.oracle_spmv_csr_f32.argtypes = [i, _i32p, _i32p, _f32p, _f32p, _f32p, f, f]
    lib.oracle_spmv_coo_f32.argtypes = [i, i64, _i32p, _i32p, _f32p, _f32p, _f32p, f, f, _f32p]
    lib.oracle_spmv_coo_inplace_f32.argtypes = [i, i64, _i32p, _i32p, _f32p, _f32p, _f32p, f, f]
    lib.oracle_gemv_f32.argtypes = [i, i, _f32p, _f32p, _f32p, f, f]
    lib.oracle_spmv_csr_f64.argtypes = [i, _i32p, _i32p, _f32p, _f32p, vp, f, f, _f64p, vp]
    lib.oracle_gemv_f64.argtypes = [i, i, _f32p, _f32p, vp, f, f, _f64p, vp]
    lib.oracle_max_scaled_error.argtypes = [i, _f32p, _f64p, _f64p, C.POINTER(i)]
    lib.oracle_max_scaled_error.restype = d
    lib.oracle_row_stats.argtypes = [i, _i32p, _i64p, C.POINTER(i), C.POINTER(i)]
    lib.oracle_select_kernel.argtypes = [i, i64, _i64p, i64, i64, i, C.POINTER(i), C.POINTER(i)]
    lib.oracle_col_probe.argtypes = [i, _i32p, _i32p, C.POINTER(i64), C.POINTER(i64)]
    lib.oracle_select_slab_cols.argtypes = [i, i64, i64, i64]
    lib.oracle_select_slab_cols.restype = i
    lib.oracle_rowstage_params.argtypes = [i, i64, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    lib.oracle_merge_tile_items.argtypes = [i, i64]
    lib.oracle_merge_tile_items.restype = i
    lib.oracle_merge_tiles.argtypes = [i, _i32p, i, vp, vp]
    lib.oracle_merge_tiles.restype = i64
    lib.oracle_split_rows.argtypes = [i, _i32p, i64, _i32p, _i64p, vp]
    lib.oracle_split_rows.restype = i64
    lib.oracle_shard_bounds.argtypes = [i, _i32p, i, _i32p]
    lib.oracle_adaptive_tiles.argtypes = [i, _i32p, i, i, i, vp, vp]
    lib.oracle_adaptive_tiles.restype = i64
    lib.oracle_adaptive_split_rows.argtypes = [i, _i32p, i, i, vp]
    lib.oracle_adaptive_split_rows.restype = i64
    lib.oracle_pb_order.argtypes = [i, i, _i32p, _i32p, _f32p, i, i, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.oracle_pb_order.restype = i64
    lib.oracle_pb_segments.argtypes = [i64, _i32p, _i32p, i, _i32p, i64, _i32p, _i32p, i, vp, vp, vp]
    lib.oracle_pb_segments.restype = i64
    lib.oracle_pb_work.argtypes = [i, _i32p, i, _i32p, i, i64, i64, _i32p]
    lib.oracle_select_blocked.argtypes = [i, i, i64, i64, i64, i64, i]
    lib.oracle_select_blocked.restype = i
    lib.oracle_pb_count_runs.argtypes = [i, _i32p, _i32p, i]
    lib.oracle_pb_count_runs.restype = i64
    lib.oracle_synth_row_len.argtypes = [i, C.c_uint64, _i64p, i64]
    lib.oracle_synth_row_len.restype = i
    lib.oracle_synth_csr.argtypes = [i, C.c_uint64, i, _i64p, i, i, vp, vp, vp]
    lib.oracle_synth_csr.restype = i64
    lib.oracle_load_mtx.argtypes = [C.c_char_p, C.POINTER(i), C.POINTER(i), vp, vp, vp]
    lib.oracle_load_mtx.restype = i64
    _u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    lib.oracle_peg_decode.argtypes = [_u64p, i, i, i64, i, i, i, vp, vp, vp, vp]
    lib.oracle_peg_decode.restype = i64
    lib.oracle_peg_spmv.argtypes = [_u64p, i, i, i64, i, i, i, i, i, _f32p, _f32p, f, f, _f32p]
    lib.oracle_peg_spmv.restype = i
    _cache["oracle"] = lib
    return lib


def _ref(name: str) -> C.CDLL:
    if name in _cache:
        return _cache[name]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            build(force=True)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (needs /root/reference at build time)")
    _cache[name] = C.CDLL(path)
    return _cache[name]


def ref_cpu() -> C.CDLL:
    lib = _ref("libref_cpu.so")
    i, i64, f, d = C.c_int, C.c_int64, C.c_float, C.c_double
    lib.ref_set_threads.argtypes = [i]
    lib.ref_get_threads.restype = i
    lib.ref_cpu_spmv.argtypes = [_i32p, _i32p, _f32p, i, i, i64, _f32p, _f32p, f, f]
    lib.ref_mkl_spmv.argtypes = [_i32p, _i32p, _f32p, i, i, i64, _f32p, _f32p, f, f, i]
    lib.ref_mkl_spmv.restype = d
    lib.ref_naive_gemv.argtypes = [_f32p, i, i, _f32p, _f32p, f, f]
    lib.ref_naive_gemv.restype = d
    lib.ref_mkl_gemv.argtypes = [_f32p, i, i, _f32p, _f32p, f, f, i]
    lib.ref_mkl_gemv.restype = d
    lib.ref_read_mtx.argtypes = [C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i64)]
    lib.ref_read_mtx.restype = i
    lib.ref_read_mtx_fetch.argtypes = [_i32p, _i32p, _f32p]
    return lib


def ref_gpuhelper() -> C.CDLL:
    lib = _ref("libref_gpuhelper.so")
    i, i64, f = C.c_int, C.c_int64, C.c_float
    lib.ref_gpu_coo_to_csr.argtypes = [i, i, i64, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p]
    lib.ref_gpu_cpu_spmv.argtypes = [i, i64, _i32p, _i32p, _f32p, i, _f32p, _f32p, f, f]
    lib.ref_gpu_load_mtx.argtypes = [C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i64)]
    lib.ref_gpu_load_mtx_fetch.argtypes = [_i32p, _i32p, _f32p]
    return lib


def ref_common() -> C.CDLL:
    lib = _ref("libref_common.so")
    i, i64, f = C.c_int, C.c_int64, C.c_float
    lib.ref_common_cpu_sequential.argtypes = [i, i, i64, _i32p, _i32p, _f32p, _f32p, _f32p, f, f, _f32p]
    lib.ref_common_coo_to_csr.argtypes = [i, i, i64, _i32p, _i32p, _f32p, _i32p, _i32p, _f32p]
    lib.ref_common_coo_to_csr.restype = i
    lib.ref_common_shared_rows.argtypes = [i, i, i, i, i64, _i32p, _i32p, _f32p, C.c_void_p, i]
    lib.ref_common_shared_rows.restype = i
    lib.ref_common_load_mtx.argtypes = [C.c_char_p, C.POINTER(i), C.POINTER(i), C.POINTER(i64)]
    lib.ref_common_load_mtx_fetch.argtypes = [_i32p, _i32p, _f32p]
    lib.ref_common_error_stats.argtypes = [i, _f32p, _f32p, C.c_char_p, i]
    lib.ref_common_error_stats.restype = i
    lib.ref_common_pack.argtypes = [i, i, i, i, i, i, i64, _i32p, _i32p, _f32p, _i64p]
    lib.ref_common_pack.restype = i
    lib.ref_common_pack_fetch.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), C.c_void_p]
    return lib


# ---- convenience wrappers (numpy in, numpy out) ---------------------------------------------------
def coo_to_csr(rows, r, c, v):
    r, c, v = (np.ascontiguousarray(r, np.int32), np.ascontiguousarray(c, np.int32), np.ascontiguousarray(v, np.float32))
    rp = np.zeros(rows + 1, np.int32)
    ci = np.zeros(r.size, np.int32)
    vv = np.zeros(r.size, np.float32)
    st = oracle().oracle_coo_to_csr(rows, r.size, r, c, v, rp, ci, vv)
    if st != 0:
        raise ValueError("oracle_coo_to_csr: row index out of range")
    return rp, ci, vv


def spmv_f64(rp, ci, v, x, y0=None, alpha=1.0, beta=0.0):
    rows = rp.size - 1
    y64 = np.zeros(rows, np.float64)
    scale = np.zeros(rows, np.float64)
    y0p = None if y0 is None else np.ascontiguousarray(y0, np.float32)
    oracle().oracle_spmv_csr_f64(rows, np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32),
                                 np.ascontiguousarray(v, np.float32), np.ascontiguousarray(x, np.float32),
                                 None if y0p is None else y0p.ctypes.data, alpha, beta, y64, scale.ctypes.data)
    return y64, scale


def gemv_f64(a, rows, cols, x, y0=None, alpha=1.0, beta=0.0):
    y64 = np.zeros(rows, np.float64)
    scale = np.zeros(rows, np.float64)
    y0p = None if y0 is None else np.ascontiguousarray(y0, np.float32)
    oracle().oracle_gemv_f64(rows, cols, np.ascontiguousarray(a, np.float32).reshape(-1),
                             np.ascontiguousarray(x, np.float32), None if y0p is None else y0p.ctypes.data, alpha,
                             beta, y64, scale.ctypes.data)
    return y64, scale


def max_scaled_error(y, y64, scale):
    at = C.c_int(-1)
    e = oracle().oracle_max_scaled_error(y.size, np.ascontiguousarray(y, np.float32), y64, scale, C.byref(at))
    return float(e), at.value


def merge_tiles(rp, tile_items):
    rp = np.ascontiguousarray(rp, np.int32)
    rows = rp.size - 1
    nt = oracle().oracle_merge_tiles(rows, rp, tile_items, None, None)
    tr = np.zeros(nt + 1, np.int32)
    tn = np.zeros(nt + 1, np.int64)
    oracle().oracle_merge_tiles(rows, rp, tile_items, tr.ctypes.data, tn.ctypes.data)
    return tr, tn


def split_rows(rp, tr, tn):
    rp = np.ascontiguousarray(rp, np.int32)
    rows = rp.size - 1
    nt = tr.size - 1
    n = oracle().oracle_split_rows(rows, rp, nt, tr, tn, None)
    out = np.zeros(n, np.int32)
    if n:
        oracle().oracle_split_rows(rows, rp, nt, tr, tn, out.ctypes.data)
    return out


def adaptive_tiles(rp, B=2048, T=1024, CH=4096):
    rp = np.ascontiguousarray(rp, np.int32)
    rows = rp.size - 1
    nt = oracle().oracle_adaptive_tiles(rows, rp, B, T, CH, None, None)
    tr = np.zeros(nt + 1, np.int32)
    tc = np.zeros(max(nt, 1), np.int32)
    oracle().oracle_adaptive_tiles(rows, rp, B, T, CH, tr.ctypes.data, tc.ctypes.data)
    tc = tc[:nt]
    tn = rp[tr].astype(np.int64)
    tn[:nt] += np.maximum(tc, 0).astype(np.int64) * CH
    ns = oracle().oracle_adaptive_split_rows(rows, rp, T, CH, None)
    sp = np.zeros(max(ns, 1), np.int32)
    oracle().oracle_adaptive_split_rows(rows, rp, T, CH, sp.ctypes.data)
    return tr, tc, tn, sp[:ns]


def row_stats(rp):
    rp = np.ascontiguousarray(rp, np.int32)
    hist = np.zeros(33, np.int64)
    mx, em = C.c_int(), C.c_int()
    oracle().oracle_row_stats(rp.size - 1, rp, hist, C.byref(mx), C.byref(em))
    return hist, mx.value, em.value


def col_probe(rp, ci):
    rp, ci = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32)
    if ci.size == 0:
        ci = np.zeros(1, np.int32)
    near, cmp_ = C.c_int64(), C.c_int64()
    oracle().oracle_col_probe(rp.size - 1, rp, ci, C.byref(near), C.byref(cmp_))
    return near.value, cmp_.value


def select_kernel(rp, ci, allow_split=1):
    """(kernel, lanes, probe_near, probe_cmp) the selector must choose for this CSR."""
    rp = np.ascontiguousarray(rp, np.int32)
    hist, mx, em = row_stats(rp)
    near, cmp_ = col_probe(rp, ci)
    k, l = C.c_int(), C.c_int()
    oracle().oracle_select_kernel(rp.size - 1, int(rp[-1]), hist, near, cmp_, allow_split, C.byref(k), C.byref(l))
    return k.value, l.value, near, cmp_


def select_slab_cols(cols, nnz, near, cmp_):
    return oracle().oracle_select_slab_cols(cols, nnz, near, cmp_)


def select_blocked(rows, cols, nnz, slab_runs, near, cmp_, allow_split=1):
    return oracle().oracle_select_blocked(rows, cols, nnz, slab_runs, near, cmp_, allow_split)


def pb_count_runs(rp, ci, W=49152):
    rp, ci = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32)
    return int(oracle().oracle_pb_count_runs(rp.size - 1, rp, ci if ci.size else np.zeros(1, np.int32), W))


def pb_plan(rp, ci, vv, cols, B, T, CH, W, align=512, n_cta=0, slab_cost=0, piece_cost16=0):
    """The blocked strategy's plan for this CSR (oracle_pb_order / oracle_adaptive_tiles over the pieces /
    oracle_pb_segments / oracle_pb_work) as a dict of numpy arrays."""
    rp, ci = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32)
    vv = np.ascontiguousarray(vv, np.float32)
    if ci.size == 0:
        ci, vv = np.zeros(1, np.int32), np.zeros(1, np.float32)
    rows = rp.size - 1
    S = (cols + W - 1) // W
    npieces = C.c_int64()
    slab_ptr = np.zeros(S + 1, np.int32)
    o = oracle()
    padded = o.oracle_pb_order(rows, cols, rp, ci, vv, W, align, slab_ptr.ctypes.data, None, None, None, None, None,
                               None, None, C.byref(npieces))
    val = np.zeros(padded, np.float32)
    lcol = np.zeros(padded, np.uint16)
    flags = np.zeros(padded // 16, np.uint16)
    group_base = np.zeros(padded // align + 1, np.int32)
    prow_ptr = np.zeros(rows + 1, np.int32)
    pcsr = np.zeros(max(npieces.value, 1), np.int32)
    pslab = np.zeros(max(npieces.value, 1), np.int32)
    o.oracle_pb_order(rows, cols, rp, ci, vv, W, align, slab_ptr.ctypes.data, val.ctypes.data, lcol.ctypes.data,
                      flags.ctypes.data, group_base.ctypes.data, prow_ptr.ctypes.data, pcsr.ctypes.data,
                      pslab.ctypes.data, C.byref(npieces))
    npc = int(npieces.value)
    tr, tc, tn, sp = adaptive_tiles(prow_ptr, B, T, CH)
    tcc = np.ascontiguousarray(tc if tc.size else np.zeros(1, np.int32), np.int32)
    npan = tr.size - 1
    perm = np.zeros(max(npc, 1), np.uint16)
    panel_seg = np.zeros(npan + 1, np.int32)
    nseg = o.oracle_pb_segments(npc, pcsr, pslab, S, prow_ptr, npan, tr, tcc, CH, None, None, None)
    seg = np.zeros((max(nseg, 1), 2), np.int32)
    o.oracle_pb_segments(npc, pcsr, pslab, S, prow_ptr, npan, tr, tcc, CH, perm.ctypes.data, panel_seg.ctypes.data,
                         seg.ctypes.data)
    d = {"slab_cols": W, "num_slabs": S, "padded_nnz": int(padded), "num_pieces": npc, "num_seg": int(nseg),
         "num_panels": npan, "slab_ptr": slab_ptr, "val": val, "lcol": lcol, "flags": flags, "group_base": group_base,
         "prow_ptr": prow_ptr, "perm": perm[:npc], "panel_seg": panel_seg, "seg": seg[:nseg],
         "tile_row": tr, "tile_chunk": tc, "tile_first": tn, "split_rows": sp}
    # the chunk table pass 2 walks (blocked.cu: pb_chunk_emit_kernel): every segment cut into runs of <= 32 pieces
    n0 = tn[:npan].astype(np.int64)
    n1 = np.where(tc >= 0, np.minimum(prow_ptr[tr[:npan] + 1], n0 + CH), prow_ptr[tr[1:npan + 1]]) if npan else n0
    chunks, panel_chunk = [], np.zeros(npan + 1, np.int32)
    for p_ in range(npan):
        panel_chunk[p_] = len(chunks)
        segs = d["seg"][panel_seg[p_]:panel_seg[p_ + 1]]
        offs = list(segs[:, 1]) + [int(n1[p_] - n0[p_])]
        for i_, (st, off) in enumerate(segs):
            ln = offs[i_ + 1] - off
            chunks += [(st + k_, min(32, ln - k_)) for k_ in range(0, ln, 32)]
    panel_chunk[npan] = len(chunks)
    d["panel_chunk"] = panel_chunk
    d["chunk"] = np.asarray(chunks, np.int32).reshape(-1, 2)
    d["num_chunks"] = len(chunks)
    # how pass 2 stages a STREAM panel (blocked.cu: pb_stage_len_kernel .. pb_end_bits_kernel): per segment one copy of
    # the 4-piece-aligned range of partial sums that covers it; perm2 = the slot of every staged position (0xFFFF = padding)
    seg_copy = np.zeros((nseg, 2), np.int32)
    panel_aux = np.zeros((npan + 1, 2), np.int32)
    perm2, bits, chunk_src = [], [], []
    for p_ in range(npan):
        panel_aux[p_] = (len(perm2), len(bits))
        if tc[p_] >= 0:
            continue                                   # LONG panel: nothing staged, no end marks
        segs = d["seg"][panel_seg[p_]:panel_seg[p_ + 1]]
        npc_p = int(n1[p_] - n0[p_])
        offs = list(segs[:, 1]) + [npc_p]
        base = len(perm2)
        for i_, (st, off) in enumerate(segs):
            ln = offs[i_ + 1] - off
            a0, a1 = st & ~3, (st + ln + 3) & ~3
            seg_copy[panel_seg[p_] + i_] = (a0, ((len(perm2) - base) >> 2) | (((a1 - a0) >> 2) << 16))
            perm2 += [int(perm[q]) if st <= q < st + ln else 0xFFFF for q in range(a0, a1)]
            chunk_src += list(range(a0, a1, 4))
        w = np.zeros((npc_p + 31) // 32, np.uint32)
        for r_ in range(tr[p_], tr[p_ + 1]):
            if prow_ptr[r_ + 1] > prow_ptr[r_]:
                j = int(prow_ptr[r_ + 1] - 1 - n0[p_])
                w[j >> 5] |= np.uint32(1 << (j & 31))
        bits += list(w)
    panel_aux[npan] = (len(perm2), len(bits))
    for p_ in range(npan):                             # LONG segments: the aligned start, no length
        if tc[p_] >= 0:
            for i_ in range(panel_seg[p_], panel_seg[p_ + 1]):
                seg_copy[i_] = (d["seg"][i_, 0] & ~3, 0)
    d["seg_copy"], d["perm2"] = seg_copy, np.asarray(perm2, np.uint16)
    d["panel_aux"], d["end_bits"] = panel_aux, np.asarray(bits, np.uint32)
    d["chunk_src"] = np.asarray(chunk_src, np.int32)
    d["stage_total"], d["bit_words"] = len(perm2), len(bits)
    if n_cta:
        work = np.zeros((n_cta, 2), np.int32)
        o.oracle_pb_work(S, slab_ptr, align, group_base, n_cta, slab_cost, piece_cost16, work.reshape(-1))
        d["work"] = work
    return d


def column_slab(rp, ci, vv, lo, hi):
    """CSR of the entries with lo <= col < hi, same rows (numpy restatement of csr_column_slab_device)."""
    keep = (ci >= lo) & (ci < hi)
    rows = rp.size - 1
    row_of = np.repeat(np.arange(rows), np.diff(rp))
    cnt = np.bincount(row_of[keep], minlength=rows)
    srp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    return srp, ci[keep], vv[keep]


def rowstage_params(rows, nnz, lanes_in=0):
    l, b, t, ch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    oracle().oracle_rowstage_params(rows, nnz, lanes_in, C.byref(l), C.byref(b), C.byref(t), C.byref(ch))
    return l.value, b.value, t.value, ch.value


def shard_bounds(rp, n_parts):
    rp = np.ascontiguousarray(rp, np.int32)
    out = np.zeros(n_parts + 1, np.int32)
    oracle().oracle_shard_bounds(rp.size - 1, rp, n_parts, out)
    return out


def synth_csr(kind, seed, cols, params, row_begin, row_end):
    p = np.asarray(params, np.int64)
    n = row_end - row_begin
    nnz = oracle().oracle_synth_csr(kind, seed, cols, p, row_begin, row_end, None, None, None)
    rp = np.zeros(n + 1, np.int32)
    ci = np.zeros(max(nnz, 1), np.int32)
    v = np.zeros(max(nnz, 1), np.float32)
    oracle().oracle_synth_csr(kind, seed, cols, p, row_begin, row_end, rp.ctypes.data, ci.ctypes.data, v.ctypes.data)
    return rp, ci[:nnz], v[:nnz]


PEG_META = ("num_pes", "pes_per_ch", "tile_rows", "tile_cols", "row_tiles", "col_tiles", "words_per_ch", "n_shared")


def ref_pack(num_ch_a, urams, fp_acc_latency, pre_acc, rows, cols, r, c, v):
    """The reference's packed PEG streams (channel-major uint64), its layout numbers and its shared-row ids."""
    r, c, v = (np.ascontiguousarray(r, np.int32), np.ascontiguousarray(c, np.int32), np.ascontiguousarray(v, np.float32))
    meta = np.zeros(8, np.int64)
    n_ch = ref_common().ref_common_pack(num_ch_a, urams, fp_acc_latency, int(pre_acc), rows, cols, r.size, r, c, v, meta)
    m = dict(zip(PEG_META, (int(t) for t in meta)))
    m["num_ch"] = n_ch
    stream = np.zeros(max(1, n_ch * m["words_per_ch"]), np.uint64)
    shared = np.zeros(max(1, m["n_shared"]), np.int32)
    ref_common().ref_common_pack_fetch(stream, shared.ctypes.data)
    return stream[:n_ch * m["words_per_ch"]], m, shared[:m["n_shared"]]


def peg_decode(stream, m):
    """(row, col, val, shared) of every real word of a packed PEG stream (oracle_peg_decode)."""
    stream = np.ascontiguousarray(stream, np.uint64)
    args = (stream, m["num_ch"], m["pes_per_ch"], m["words_per_ch"], m["tile_rows"], m["tile_cols"], m["col_tiles"])
    n = oracle().oracle_peg_decode(*args, None, None, None, None)
    if n < 0:
        raise ValueError(f"oracle_peg_decode: malformed stream ({n})")
    row, col = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    val, sh = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.uint8)
    oracle().oracle_peg_decode(*args, row.ctypes.data, col.ctypes.data, val.ctypes.data, sh.ctypes.data)
    return row[:n], col[:n], val[:n], sh[:n]


def peg_spmv(stream, m, rows, cols, x, c_in, alpha, beta):
    """y = beta*c_in + alpha*A x walked the way the accelerator consumes the stream (oracle_peg_spmv)."""
    y = np.zeros(rows, np.float32)
    st = oracle().oracle_peg_spmv(np.ascontiguousarray(stream, np.uint64), m["num_ch"], m["pes_per_ch"],
                                  m["words_per_ch"], m["tile_rows"], m["tile_cols"], m["col_tiles"], rows, cols,
                                  np.ascontiguousarray(x, np.float32), np.ascontiguousarray(c_in, np.float32),
                                  alpha, beta, y)
    if st < 0:
        raise ValueError(f"oracle_peg_spmv: malformed stream ({st})")
    return y, st


def load_mtx(path):
    r_, c_ = C.c_int(), C.c_int()
    n = oracle().oracle_load_mtx(path.encode(), C.byref(r_), C.byref(c_), None, None, None)
    if n < 0:
        raise IOError(f"oracle_load_mtx failed with {n}")
    r = np.zeros(max(n, 1), np.int32)
    c = np.zeros(max(n, 1), np.int32)
    v = np.zeros(max(n, 1), np.float32)
    oracle().oracle_load_mtx(path.encode(), C.byref(r_), C.byref(c_), r.ctypes.data, c.ctypes.data, v.ctypes.data)
    return r[:n], c[:n], v[:n], r_.value, c_.value
