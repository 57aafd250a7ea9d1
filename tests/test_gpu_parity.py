"""Parity tests proper: the CUDA path, called through the C-ABI / the pyhispmv plugin, against the oracle.

Bars: bit-exact for every integer artefact (CSR, tile coordinates, split rows, shard bounds, selector);
floating point per element |y - y64| / (|alpha| sum_j |a_ij x_j| + |beta y0_i|) <= 1e-5 (north_star), with the
float64 restatement as y64.  Where oracle/_ref is present the reference's own MKL result is held to the same
bar beside ours.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

TOL = 1e-5
ALPHA, BETA = np.float32(0.85), np.float32(-2.06)


@pytest.fixture(scope="module")
def eng():
    from hispmv_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _needs_research_kernels():
    """The three research kernels live in experimental.cu and are only in the library after `make EXPERIMENTAL=1`
    (hispmv_version() is odd then); the default build -- what ships -- does not carry them."""
    from hispmv_b200 import capi
    if not capi.lib.hispmv_version() & 1:
        pytest.skip("library built without the research kernels (make EXPERIMENTAL=1)")


def _rand_coo(rng, rows, cols, nnz, dup=0.0):
    r = rng.integers(0, rows, nnz).astype(np.int32)
    c = rng.integers(0, cols, nnz).astype(np.int32)
    v = rng.standard_normal(nnz).astype(np.float32)
    if dup and nnz:
        k = int(nnz * dup)
        src, dst = rng.integers(0, nnz, k), rng.integers(0, nnz, k)
        r[dst], c[dst] = r[src], c[src]
    return r, c, v


def _check_run(eng, idx, rp, ci, vv, rows, cols, rng, alpha=ALPHA, beta=BETA):
    x = rng.standard_normal(cols).astype(np.float32)
    y0 = rng.standard_normal(rows).astype(np.float32)
    y = np.full(rows, np.nan, np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(x, y0, y, float(alpha), float(beta))
    y64, scale = ol.spmv_f64(rp, ci, vv, x, y0, alpha, beta)
    err, at = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL, (err, at, eng.matrix_info(idx)["kernel_name"])
    return y


# ---------------------------------------------------------------------------------------------------
# integer contract
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols,nnz,dup", [(1, 1, 1, 0), (5, 9, 0, 0), (64, 64, 700, 0.3), (3000, 777, 50000, 0.05),
                                                (100000, 100000, 1000000, 0.0), (50000, 10000, 1000000, 0.01)])
def test_coo_to_csr_bit_exact(eng, rows, cols, nnz, dup):
    rng = np.random.default_rng(nnz + rows)
    r, c, v = _rand_coo(rng, rows, cols, nnz, dup)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    assert idx >= 0
    rp, ci, vv = eng.plan_csr(idx)
    rp2, ci2, vv2 = ol.coo_to_csr(rows, r, c, v)
    assert np.array_equal(rp, rp2)
    assert np.array_equal(ci, ci2)
    assert np.array_equal(vv.view(np.uint32), vv2.view(np.uint32))
    info = eng.matrix_info(idx)
    hist, mx, em = ol.row_stats(rp2)
    assert info["hist"] == hist.tolist() and info["max_row_nnz"] == mx and info["empty_rows"] == em
    k, l, near, cmp_ = ol.select_kernel(rp2, ci2, 1)
    assert (info["kernel"], info["probe_near"], info["probe_cmp"]) == (k, near, cmp_)


def test_csr_input_is_validated_and_unsorted_rows_are_resorted(eng):
    """hispmv_add_sparse_csr: a row_ptr that is not monotone from 0 to nnz or a column outside [0, cols) is refused; rows
    whose columns are out of order come out exactly as the COO path orders them (the reference's per-row sort)."""
    from hispmv_b200.capi import HispmvError
    rng = np.random.default_rng(3)
    rows, cols = 500, 900
    lens = rng.integers(0, 40, rows)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)          # unsorted inside every row, with duplicates
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle_csr(rp, c, v, rows, cols)
    got = eng.plan_csr(idx)
    want = ol.coo_to_csr(rows, r, c, v)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    _check_run(eng, idx, *want, rows, cols, rng)
    bad_rp = rp.copy()
    bad_rp[10], bad_rp[11] = bad_rp[11] + 5, bad_rp[10]
    with pytest.raises(HispmvError, match="row_ptr"):
        eng.create_sparse_handle_csr(bad_rp, c, v, rows, cols)
    bad_c = c.copy()
    bad_c[r.size // 2] = cols
    with pytest.raises(HispmvError, match="column index"):
        eng.create_sparse_handle_csr(rp, bad_c, v, rows, cols)
    n_before = eng.num_matrices()
    cs = want[1]                                               # sorted input is taken as it is
    assert eng.create_sparse_handle_csr(want[0], cs, want[2], rows, cols) == n_before


def test_coo_index_out_of_range_is_refused(eng):
    from hispmv_b200.capi import HispmvError
    with pytest.raises(HispmvError):
        eng.create_sparse_handle(np.array([0, 5], np.int32), np.array([0, 1], np.int32), np.ones(2, np.float32), 4, 4)


@pytest.mark.parametrize("tile", [128 * 7, 256 * 7, 256 * 11, 512 * 7])
def test_merge_plan_bit_exact(eng, tile, monkeypatch):
    from hispmv_b200 import capi
    monkeypatch.setenv("HISPMV_MERGE_TILE", str(tile))
    rng = np.random.default_rng(tile)
    rows, cols = 20000, 20000
    lens = np.minimum(rng.zipf(1.7, rows), 30000)
    lens[rng.integers(0, rows, 3)] = 15000
    lens[rng.integers(0, rows, 2000)] = 0
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_MERGE)
    info = eng.matrix_info(idx)
    assert info["kernel"] == capi.KERNEL_MERGE and info["tile_items"] == tile
    rp, ci, vv = eng.plan_csr(idx)
    tr, tn = eng.plan_tiles(idx)
    tr2, tn2 = ol.merge_tiles(rp, tile)
    assert np.array_equal(tr, tr2) and np.array_equal(tn, tn2)
    sp = eng.plan_split_rows(idx)
    sp2 = ol.split_rows(rp, tr2, tn2)
    assert np.array_equal(sp, sp2)
    assert info["num_split_rows"] == sp2.size
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_adaptive_plan_bit_exact(eng, seed):
    """Tile descriptors (first row, first-nonzero offset, chunk index) and the split-row list of the adaptive
    kernel against the sequential restatement; includes heavy rows, runs of empty rows and long rows back to back."""
    from hispmv_b200 import capi
    rng = np.random.default_rng(seed)
    rows, cols = 30000, 50000
    lens = np.minimum(rng.zipf(1.6, rows), 40000)
    heavy = rng.integers(1, rows - 2, 6)
    lens[heavy] = rng.integers(1024, 20000, 6)
    lens[heavy[0] + 1] = 5000           # two long rows in a row
    lens[0] = 9000 if seed == 1 else lens[0]
    lens[rows - 1] = 4096 if seed == 2 else lens[rows - 1]
    lens[rng.integers(0, rows, 3000)] = 0
    lens[1000:4500] = 0                 # > stream_items empty rows in a row
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
    info = eng.matrix_info(idx)
    assert info["kernel"] == capi.KERNEL_ADAPTIVE
    rp, ci, vv = eng.plan_csr(idx)
    tr2, tc2, tn2, sp2 = ol.adaptive_tiles(rp, info["tile_items"], info["long_threshold"], info["chunk_nnz"])
    assert info["num_tiles"] == tc2.size and info["num_split_rows"] == sp2.size
    tr, tn = eng.plan_tiles(idx)
    assert np.array_equal(tr, tr2) and np.array_equal(tn, tn2)
    assert np.array_equal(eng.plan_tile_chunks(idx), tc2)
    assert np.array_equal(eng.plan_split_rows(idx), sp2)
    y1 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))
    y2 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))   # counters reset themselves
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))


@pytest.mark.parametrize("lanes", [0, 1, 4, 32])
def test_rowstage_plan_bit_exact(eng, lanes):
    """The TMA-staged row-major kernel shares the adaptive tiling; its plan parameters (lanes, B, T, CH) and the
    resulting tile descriptors must equal the sequential restatement."""
    from hispmv_b200 import capi
    rng = np.random.default_rng(100 + lanes)
    rows, cols = 40000, 40000
    lens = rng.integers(20, 34, rows)
    lens[rng.integers(0, rows, 5)] = rng.integers(512, 9000, 5)     # a few LONG rows, some split
    lens[rng.integers(0, rows, 500)] = 0
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = np.clip(r + rng.integers(-40, 41, r.size), 0, cols - 1).astype(np.int32)   # banded
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_ROWSTAGE, lanes)
    info = eng.matrix_info(idx)
    assert info["kernel"] == capi.KERNEL_ROWSTAGE
    l, B, T, CH = ol.rowstage_params(rows, r.size, lanes)
    assert (info["vector_lanes"], info["tile_items"], info["long_threshold"], info["chunk_nnz"]) == (l, B, T, CH)
    rp, ci, vv = eng.plan_csr(idx)
    tr2, tc2, tn2, sp2 = ol.adaptive_tiles(rp, B, T, CH)
    assert info["num_tiles"] == tc2.size and info["num_split_rows"] == sp2.size
    tr, tn = eng.plan_tiles(idx)
    assert np.array_equal(tr, tr2) and np.array_equal(tn, tn2)
    assert np.array_equal(eng.plan_tile_chunks(idx), tc2)
    assert np.array_equal(eng.plan_split_rows(idx), sp2)
    y1 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))
    y2 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))


@pytest.mark.parametrize("window", [1, 777, 16000, 40960])
def test_persistent_window_kernel(eng, window, monkeypatch):
    """The persistent ADAPTIVE kernel (x[0, window) in shared memory, tiles pulled from a global counter): same
    plan artefacts as the one-tile-per-CTA kernel, results within tolerance, bit-identical from run to run even
    though the tile-to-group assignment is dynamic."""
    from hispmv_b200 import capi
    _needs_research_kernels()
    monkeypatch.setenv("HISPMV_PERSIST", str(window))
    rng = np.random.default_rng(window)
    rows, cols = 30000, 16000 if window == 16000 else 50000
    lens = np.minimum(rng.zipf(1.6, rows), 40000)
    heavy = rng.integers(1, rows - 2, 6)
    lens[heavy] = rng.integers(1024, 20000, 6)
    lens[rng.integers(0, rows, 3000)] = 0
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = (rng.random(r.size) ** 4 * cols).astype(np.int32)            # head-heavy columns
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
    info = eng.matrix_info(idx)
    assert info["kernel"] == capi.KERNEL_ADAPTIVE and info["x_window_cols"] == min(window, cols)
    rp, ci, vv = eng.plan_csr(idx)
    tr2, tc2, tn2, sp2 = ol.adaptive_tiles(rp, info["tile_items"], info["long_threshold"], info["chunk_nnz"])
    assert np.array_equal(eng.plan_tile_chunks(idx), tc2) and np.array_equal(eng.plan_split_rows(idx), sp2)
    ys = [_check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7)) for _ in range(3)]
    assert np.array_equal(ys[0].view(np.uint32), ys[1].view(np.uint32))
    assert np.array_equal(ys[0].view(np.uint32), ys[2].view(np.uint32))
    monkeypatch.setenv("HISPMV_PERSIST", "0")
    eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
    assert eng.matrix_info(idx)["x_window_cols"] == 0
    _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))


@pytest.mark.parametrize("seed", [0, 1])
def test_pipeline_kernel(eng, seed, monkeypatch):
    """The warp-specialised persistent pipeline (TMA producer / gather teams / reduce warps over an mbarrier ring):
    same tiles as the one-CTA-per-tile kernel, results within tolerance and bit-identical from run to run."""
    from hispmv_b200 import capi
    _needs_research_kernels()
    monkeypatch.setenv("HISPMV_PIPELINE", "1")
    rng = np.random.default_rng(50 + seed)
    rows, cols = 40000, 60000
    lens = np.minimum(rng.zipf(1.6, rows), 40000)
    heavy = rng.integers(1, rows - 2, 8)
    lens[heavy] = rng.integers(1024, 30000, 8)
    lens[rng.integers(0, rows, 3000)] = 0
    lens[5000:9000] = 0                         # tiles made of empty rows only (no bytes to stage)
    if seed == 1:
        lens[:] = np.minimum(lens, 3)           # very short rows: many rows per tile
        lens[17] = 20000
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = (rng.random(r.size) ** 3 * cols).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
    info = eng.matrix_info(idx)
    assert info["chunk_nnz"] <= 3072
    rp, ci, vv = eng.plan_csr(idx)
    tr2, tc2, tn2, sp2 = ol.adaptive_tiles(rp, info["tile_items"], info["long_threshold"], info["chunk_nnz"])
    assert np.array_equal(eng.plan_tile_chunks(idx), tc2) and np.array_equal(eng.plan_split_rows(idx), sp2)
    ys = [_check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7)) for _ in range(3)]
    assert np.array_equal(ys[0].view(np.uint32), ys[1].view(np.uint32))
    assert np.array_equal(ys[0].view(np.uint32), ys[2].view(np.uint32))
    _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(8), alpha=1.0, beta=0.0)


@pytest.mark.parametrize("kind,expect", [("banded", "rowstage"), ("banded_heavy", "adaptive"), ("random", "adaptive"),
                                         ("stencil", "rowstage")])
def test_selector_bit_exact(eng, kind, expect):
    """The runtime selector (row-length histogram + column-locality probe) against the oracle's restatement, on
    matrices built to land on either side of both rules."""
    rng = np.random.default_rng(sum(map(ord, kind)))
    rows = cols = 60000
    if kind == "stencil":
        from hispmv_b200 import synth
        spec = synth.c4_stencil(0.003)
        rp, ci, vv = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, spec.rows)
        rows = cols = spec.rows
        idx = eng.create_sparse_handle_csr(rp, ci, vv, rows, cols)
    else:
        lens = rng.integers(8, 24, rows)
        if kind == "banded_heavy":
            lens[rng.integers(0, rows, 300)] = 1500                   # > 1/8 of the nonzeros in rows >= 4 * mean
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        if kind == "random":
            c = rng.integers(0, cols, r.size).astype(np.int32)
        else:
            c = np.clip(r + rng.integers(-30, 31, r.size), 0, cols - 1).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
        idx = eng.create_sparse_handle(r, c, v, rows, cols)
        rp, ci, vv = eng.plan_csr(idx)
    info = eng.matrix_info(idx)
    k, l, near, cmp_ = ol.select_kernel(rp, ci, 1)
    assert (info["kernel"], info["probe_near"], info["probe_cmp"]) == (k, near, cmp_)
    assert info["kernel_name"] == expect
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng)


def test_selector_without_row_dist_net():
    """row_dist_net=False (the reference's switch for its shared-row network) forbids splitting rows across CTAs:
    the selector must stay on the one-row-per-lane-group kernels."""
    from hispmv_b200 import Engine
    e = Engine(0, row_dist_net=False)
    rng = np.random.default_rng(5)
    rows, cols = 20000, 20000
    for mean, want in ((1, "csr_scalar"), (12, "csr_vector")):
        lens = rng.integers(0, 2 * mean + 1, rows)
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
        idx = e.create_sparse_handle(r, c, v, rows, cols)
        rp, ci, vv = e.plan_csr(idx)
        info = e.matrix_info(idx)
        k, l, _, _ = ol.select_kernel(rp, ci, 0)
        assert (info["kernel"], info["vector_lanes"]) == (k, l) and info["kernel_name"] == want
        _check_run(e, idx, rp, ci, vv, rows, cols, rng)
    e.close()


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_shard_bounds_bit_exact(eng, parts):
    from hispmv_b200 import shard_bounds
    rng = np.random.default_rng(parts)
    lens = np.minimum(rng.zipf(1.6, 5000), 4000)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    assert np.array_equal(shard_bounds(rp, parts), ol.shard_bounds(rp, parts))


# ---------------------------------------------------------------------------------------------------
# floating point: every kernel strategy against the float64 restatement
# ---------------------------------------------------------------------------------------------------
def _matrix(rng, kind, rows, cols):
    if kind == "powerlaw":
        lens = np.minimum(rng.zipf(1.8, rows), cols)
        lens[rng.integers(0, rows, 4)] = min(cols, 25000)
    elif kind == "regular":
        lens = np.full(rows, 27)
    elif kind == "short":
        lens = rng.integers(0, 3, rows)
    elif kind == "hollow":
        lens = np.where(rng.random(rows) < 0.8, 0, rng.integers(1, 40, rows))
    else:
        raise ValueError(kind)
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    return r, c, v


@pytest.mark.parametrize("kind", ["powerlaw", "regular", "short", "hollow"])
@pytest.mark.parametrize("kernel,lanes", [(1, 0), (2, 2), (2, 8), (2, 32), (3, 0), (6, 0), (7, 0), (7, 1), (7, 2), (7, 8),
                                          (0, 0)])
def test_spmv_kernels_within_tolerance(eng, kind, kernel, lanes):
    rng = np.random.default_rng(sum(map(ord, kind)) * 131 + kernel * 17 + lanes)
    rows, cols = 30011, 40009
    r, c, v = _matrix(rng, kind, rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    if kernel:
        eng.force_kernel(idx, kernel, lanes)
    rp, ci, vv = eng.plan_csr(idx)
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng)
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng, alpha=1.0, beta=0.0)


def test_run_is_deterministic(eng):
    rng = np.random.default_rng(11)
    rows, cols = 20000, 20000
    r, c, v = _matrix(rng, "powerlaw", rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    rp, ci, vv = eng.plan_csr(idx)
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    eng.select_matrix(idx)
    ys = []
    for _ in range(3):
        y = np.zeros(rows, np.float32)
        eng.run_kernel(x, b, y, 0.55, -2.05)
        ys.append(y.copy())
    assert np.array_equal(ys[0].view(np.uint32), ys[1].view(np.uint32))
    assert np.array_equal(ys[0].view(np.uint32), ys[2].view(np.uint32))


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 5), (257, 1031), (1024, 4096), (1000, 10000), (64, 50000), (5, 70001)])
def test_gemv_within_tolerance(eng, rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    a = rng.standard_normal((rows, cols)).astype(np.float32)
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    idx = eng.create_dense_handle(a.reshape(-1), rows, cols)
    assert eng.matrix_info(idx)["kernel_name"] == "gemv"
    eng.select_matrix(idx)
    y = np.full(rows, np.nan, np.float32)
    eng.run_kernel(x, b, y, float(ALPHA), float(BETA))
    y64, scale = ol.gemv_f64(a, rows, cols, x, b, ALPHA, BETA)
    err, at = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL, (err, at)


def test_reference_closed_form_inputs_vs_mkl(eng):
    """The reference's own self-check protocol (cpu/src/main.cpp:147-148,173-178) on a C1-shaped matrix; our
    result and the reference's mkl_sparse_s_mv result are both held to the 1e-5 bar against float64."""
    from hispmv_b200.synth import c1_imbalanced_coo, reference_vectors
    r, c, v, n, _ = c1_imbalanced_coo(n=8192, target_nnz=110000, dense_rows=4, dense_len=3000)
    idx = eng.create_sparse_handle(r, c, v, n, n)
    rp, ci, vv = eng.plan_csr(idx)
    x, y0 = reference_vectors(n, n)
    y = np.zeros(n, np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(x, y0, y, float(ALPHA), float(BETA))
    y64, scale = ol.spmv_f64(rp, ci, vv, x, y0, ALPHA, BETA)
    err, _ = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL
    if ol.have_ref():
        y_mkl = y0.copy()
        ol.ref_cpu().ref_mkl_spmv(rp, ci, vv, n, n, ci.size, x, y_mkl, ALPHA, BETA, 1)
        err_mkl, _ = ol.max_scaled_error(y_mkl, y64, scale)
        assert err_mkl <= TOL
        # ours vs MKL directly, same normaliser
        d = np.abs(y.astype(np.float64) - y_mkl.astype(np.float64))
        assert np.all(d <= 2 * TOL * np.maximum(scale, 1e-30))


# ---------------------------------------------------------------------------------------------------
# the plugin surface (pyhispmv), mirroring apps/general_test.py:22-116
# ---------------------------------------------------------------------------------------------------
def test_pyhispmv_general_test_flow():
    import pyhispmv
    fpga = pyhispmv.FpgaHandle("unused.xclbin", 0, 24, 1, 1, 2, 5, True, False, True)
    rng = np.random.default_rng(0)
    rows, cols = 5000, 1000
    dense = rng.random((rows, cols), dtype=np.float32)
    x = rng.random(cols, dtype=np.float32)
    bias = rng.random(rows, dtype=np.float32)
    nnz = 100000
    cr = rng.integers(0, rows, nnz).astype(np.int32)
    cc = rng.integers(0, cols, nnz).astype(np.int32)
    cv = rng.random(nnz, dtype=np.float32)
    di = fpga.create_dense_handle(dense.flatten(), rows, cols)
    si = fpga.create_sparse_handle(cr, cc, cv, rows, cols)
    assert (di, si) == (0, 1)
    fpga.load_matrices()
    fpga.load_matrices()  # idempotent here
    y_d = np.zeros(rows, np.float32)
    y_s = np.zeros(rows, np.float32)
    fpga.select_matrix(di)
    fpga.run_kernel(x, bias, y_d, 1.0, 1.0)
    fpga.select_matrix(si)
    fpga.run_kernel(x, bias, y_s, 1.0, 1.0)
    from scipy.sparse import coo_matrix
    exp_d = dense.astype(np.float64) @ x + bias
    exp_s = coo_matrix((cv.astype(np.float64), (cr, cc)), shape=(rows, cols)).dot(x.astype(np.float64)) + bias
    assert np.allclose(y_d, exp_d, rtol=1e-3)   # the reference's own bar (general_test.py:106,113)
    assert np.allclose(y_s, exp_s, rtol=1e-3)
    assert np.max(np.abs(y_d - exp_d) / np.abs(exp_d)) < 1e-5
    assert np.max(np.abs(y_s - exp_s) / np.abs(exp_s)) < 1e-5
    # linear(): alpha = beta = 1, several vectors, new array back
    xs = rng.random((3, cols), dtype=np.float32)
    out = fpga.linear(si, xs.reshape(-1), bias)
    assert out.shape == (3 * rows,) and out.dtype == np.float32
    for k in range(3):
        e = coo_matrix((cv.astype(np.float64), (cr, cc)), shape=(rows, cols)).dot(xs[k].astype(np.float64)) + bias
        assert np.allclose(out[k * rows:(k + 1) * rows], e, rtol=1e-4, atol=1e-5)
    with pytest.raises(IndexError):
        fpga.select_matrix(7)
    with pytest.raises(TypeError):
        fpga.run_kernel(x, bias, np.zeros(rows, np.float64), 1.0, 1.0)  # y must be float32, written in place


def test_pyhispmv_misuse_errors():
    import pyhispmv
    fpga = pyhispmv.FpgaHandle("x", 0, 24, 1, 1, 2, 5, False, False, True)  # dense_overlay off
    with pytest.raises(RuntimeError):
        fpga.create_dense_handle(np.zeros(4, np.float32), 2, 2)
    with pytest.raises(RuntimeError):
        fpga.run_kernel(np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros(2, np.float32), 1.0, 1.0)
    with pytest.raises(RuntimeError):
        pyhispmv.FpgaHandle("x", 99, 24, 1, 1, 2, 5, True, False, True)  # no such device: error, not exit()


def test_memory_full_returns_minus_one():
    from hispmv_b200 import Engine
    e = Engine(0, memory_limit=1 << 20)
    rng = np.random.default_rng(0)
    assert e.create_dense_handle(np.zeros(16, np.float32), 4, 4) == 0
    assert e.create_dense_handle(rng.random(1 << 20, dtype=np.float32), 1024, 1024) == -1   # 4 MiB > 1 MiB cap
    r, c, v = _rand_coo(rng, 1000, 1000, 400000)
    assert e.create_sparse_handle(r, c, v, 1000, 1000) == -1
    assert e.num_matrices() == 1
    e.close()


def test_load_mtx_on_device(eng, tmp_path):
    from hispmv_b200.synth import c1_imbalanced_coo, write_mtx
    r, c, v, n, _ = c1_imbalanced_coo(n=2048, target_nnz=20000, dense_rows=2, dense_len=1500)
    p = tmp_path / "m.mtx"
    write_mtx(str(p), r, c, v, n, n)
    idx = eng.load_mtx(str(p))
    rp, ci, vv = eng.plan_csr(idx)
    r2, c2, v2, nr, nc = ol.load_mtx(str(p))
    rp2, ci2, vv2 = ol.coo_to_csr(nr, r2, c2, v2)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(vv, vv2)
    sym = tmp_path / "s.mtx"
    sym.write_text("%%MatrixMarket matrix coordinate real symmetric\n4 4 4\n1 1 1\n3 1 2.5\n4 2 -1\n4 4 3\n")
    idx = eng.load_mtx(str(sym))
    rp, ci, vv = eng.plan_csr(idx)
    r2, c2, v2, nr, nc = ol.load_mtx(str(sym))
    rp2, ci2, vv2 = ol.coo_to_csr(nr, r2, c2, v2)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(vv, vv2)


# ---------------------------------------------------------------------------------------------------
# synthetic generators: device output bit-exact against the CPU restatement
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["c2", "c2_weak3", "c4", "c5"])
def test_synth_generators_bit_exact(eng, which):
    from hispmv_b200 import synth, workloads
    spec = {"c2": synth.c2_powerlaw(0.002), "c2_weak3": workloads.c2_weak(3, 0.002), "c4": synth.c4_stencil(0.0004),
            "c5": synth.c5_uniform(0.0002)}[which]
    d = synth.DeviceCSR(spec)
    idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
    rp, ci, vv = eng.plan_csr(idx)
    rp2, ci2, vv2 = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, spec.rows)
    assert d.nnz == ci2.size
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(vv.view(np.uint32), vv2.view(np.uint32))
    # columns sorted inside each row
    row_of = np.repeat(np.arange(spec.rows), np.diff(rp2))
    same_row = row_of[1:] == row_of[:-1]
    assert np.all(ci2[1:][same_row] >= ci2[:-1][same_row])
    rng = np.random.default_rng(1)
    _check_run(eng, idx, rp2, ci2, vv2, spec.rows, spec.cols, rng)
    # a row block generated on its own equals the slice of the whole
    b0, b1 = spec.rows // 3, 2 * spec.rows // 3
    blk = synth.DeviceCSR(spec, b0, b1)
    assert blk.nnz == rp2[b1] - rp2[b0]
    bounds, total = synth.synth_shard_bounds(spec, 4)
    assert total == ci2.size and np.array_equal(bounds, ol.shard_bounds(rp2, 4))
    if which == "c2_weak3":   # three stacked blocks with the same row lengths: nnz-balanced thirds are the blocks
        b3, _ = synth.synth_shard_bounds(spec, 3)
        n = spec.rows // 3
        assert np.array_equal(np.diff(rp2)[:n], np.diff(rp2)[n:2 * n]) and rp2[n] * 3 == ci2.size
        assert b3[0] == 0 and b3[3] == spec.rows and rp2[b3[1]] == rp2[n] and rp2[b3[2]] == rp2[2 * n]
    d.close()
    blk.close()


# ---------------------------------------------------------------------------------------------------
# device-resident calls and sharding inside one process
# ---------------------------------------------------------------------------------------------------
def test_run_dev_and_linear_dev_relu(eng):
    import torch
    rng = np.random.default_rng(5)
    rows, cols = 20000, 9000
    r, c, v = _matrix(rng, "powerlaw", rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    rp, ci, vv = eng.plan_csr(idx)
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    xd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(b).cuda()
    yd = torch.empty(rows, dtype=torch.float32, device="cuda")
    for kernel in (0, 1, 2, 3, 6):
        eng.force_kernel(idx, kernel)
        stream = torch.cuda.current_stream().cuda_stream
        eng.linear_dev(idx, xd, bd, yd, relu=True, stream=stream)
        torch.cuda.synchronize()
        y64, scale = ol.spmv_f64(rp, ci, vv, x, b, 1.0, 1.0)
        y = yd.cpu().numpy()
        assert np.all(y >= 0)
        pos = y64 > 1e-4 * np.maximum(scale, 1e-30)
        err, _ = ol.max_scaled_error(y[pos], y64[pos], scale[pos])
        assert err <= TOL
        neg = y64 < -1e-4 * np.maximum(scale, 1e-30)
        assert np.all(y[neg] == 0)


@pytest.mark.parametrize("parts", [2, 3])
def test_row_block_shards_reassemble(parts):
    """Each shard context keeps only its nnz-balanced row block; concatenated y equals the unsharded y."""
    from hispmv_b200 import Engine
    rng = np.random.default_rng(parts)
    rows, cols = 30000, 12000
    r, c, v = _matrix(rng, "powerlaw", rows, cols)
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    full = Engine(0)
    fi = full.create_sparse_handle(r, c, v, rows, cols)
    rp, ci, vv = full.plan_csr(fi)
    bounds = ol.shard_bounds(rp, parts)
    pieces = []
    for p in range(parts):
        e = Engine(0, shard=(p, parts))
        i = e.create_sparse_handle(r, c, v, rows, cols)
        info = e.matrix_info(i)
        assert (info["row_begin"], info["row_end"]) == (bounds[p], bounds[p + 1])
        y = np.zeros(info["row_end"] - info["row_begin"], np.float32)
        e.select_matrix(i)
        e.run_kernel(x, b[info["row_begin"]:info["row_end"]], y, 0.55, -2.05)
        pieces.append(y)
        e.close()
    y = np.concatenate(pieces)
    y64, scale = ol.spmv_f64(rp, ci, vv, x, b, 0.55, -2.05)
    err, _ = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL
    full.close()


# ---------------------------------------------------------------------------------------------------
# the DNN-layer callers: apps/model_test.py flow (host plugin path) and the device-resident chain
# ---------------------------------------------------------------------------------------------------
def _mlp(seed=0, sizes=(512, 1024, 1024, 256)):
    import torch
    from hispmv_b200.layers import ThreeLayerFCModel, ThreeLayerFCModelConfig
    torch.manual_seed(seed)
    cfg = ThreeLayerFCModelConfig(sizes[0], sizes[1], sizes[2], sizes[3], 0.1, 0.25)
    m = ThreeLayerFCModel(cfg).eval()
    for p in m.parameters():
        p.requires_grad = False
    return m


@pytest.mark.parametrize("gpus", [1, 2])
def test_model_test_flow_through_plugin(gpus, monkeypatch):
    """apps/model_test.py:53-90: replace_layers(cpu_model, fpga), one forward of a random input, compare with the
    CPU model.  Tolerance as apps/general_test.py:106 (rtol 1e-3), plus the north-star bar per layer output.
    gpus = 2: the same unchanged caller with HISPMV_GPUS=2 in the environment -- one FpgaHandle, every layer row-sharded
    over two GPUs (over the one GPU twice, HISPMV_MULTI_WRAP, where the box has a single one)."""
    import torch
    import pyhispmv
    from hispmv_b200.layers import FpgaLayerManager
    if gpus > 1:
        monkeypatch.setenv("HISPMV_GPUS", str(gpus))
        monkeypatch.setenv("HISPMV_MULTI_WRAP", "1")
    cpu_model = _mlp()
    fpga = pyhispmv.FpgaHandle("unused.xclbin", 0, 24, 1, 1, 2, 5, True, False, True)
    fpga_model = FpgaLayerManager().replace_layers(cpu_model, fpga)
    x = torch.randn((1, 512))
    with torch.no_grad():
        ref = cpu_model(x).numpy()
        out = fpga_model(x).numpy()
    assert out.shape == ref.shape == (1, 256)
    assert np.allclose(out, ref, rtol=1e-3, atol=1e-4)
    # batch > 1 goes through linear()'s vector loop (fpga_handle.cpp:336, 366-379)
    xb = torch.randn((3, 512))
    with torch.no_grad():
        assert np.allclose(fpga_model(xb).numpy(), cpu_model(xb).numpy(), rtol=1e-3, atol=1e-4)


def test_one_handle_over_several_gpus(monkeypatch, tmp_path):
    """hispmv_create_multi: ONE handle in ONE process over several GPUs (the reference's single-handle model): matrices
    are row-sharded at the oracle's nnz-balanced split points, run_kernel / linear fan out and write the blocks of y side
    by side; results within the bar, handle indices aligned, -1 (memory full) rolls back on every GPU, device-pointer
    calls refuse.  On a one-GPU box the children share the GPU (HISPMV_MULTI_WRAP)."""
    import ctypes as C
    import torch
    from hispmv_b200 import Engine, capi
    from hispmv_b200.capi import lib, HispmvError
    monkeypatch.setenv("HISPMV_MULTI_WRAP", "1")
    rng = np.random.default_rng(77)
    e = Engine(0, n_gpus=3)
    try:
        assert lib.hispmv_multi_gpus(e._ctx) == 3
        rows, cols = 30011, 20011
        r, c, v = _matrix(rng, "powerlaw", rows, cols)
        idx = e.create_sparse_handle(r, c, v, rows, cols)
        a = rng.standard_normal((700, 1300)).astype(np.float32)
        di = e.create_dense_handle(a.reshape(-1), 700, 1300)
        assert (idx, di) == (0, 1) and e.num_matrices() == 2
        rp, ci, vv = ol.coo_to_csr(rows, r, c, v)
        info = e.matrix_info(idx)
        assert (info["row_begin"], info["row_end"], info["nnz"]) == (0, rows, ci.size)
        bounds = ol.shard_bounds(rp, 3)
        for k in range(3):   # every child holds exactly the oracle's block
            kid = lib.hispmv_multi_child(e._ctx, k)
            ki = capi.MatrixInfo()
            assert lib.hispmv_matrix_info_get(C.c_void_p(kid), idx, C.byref(ki)) == 0
            assert (ki.row_begin, ki.row_end) == (int(bounds[k]), int(bounds[k + 1]))
        e.load_matrices()
        x = rng.standard_normal(cols).astype(np.float32)
        b = rng.standard_normal(rows).astype(np.float32)
        y = np.full(rows, np.nan, np.float32)
        e.select_matrix(idx)
        e.run_kernel(x, b, y, float(ALPHA), float(BETA))
        y64, scale = ol.spmv_f64(rp, ci, vv, x, b, ALPHA, BETA)
        assert ol.max_scaled_error(y, y64, scale)[0] <= TOL
        X = rng.standard_normal((3, cols)).astype(np.float32)
        Y = e.linear(idx, X.reshape(-1), b).reshape(3, rows)
        for k in range(3):
            y64, scale = ol.spmv_f64(rp, ci, vv, X[k], b, 1.0, 1.0)
            assert ol.max_scaled_error(Y[k], y64, scale)[0] <= TOL
        xd = rng.standard_normal(1300).astype(np.float32)
        bd = rng.standard_normal(700).astype(np.float32)
        yd = e.linear(di, xd, bd)
        y64, scale = ol.gemv_f64(a, 700, 1300, xd, bd, 1.0, 1.0)
        assert ol.max_scaled_error(yd, y64, scale)[0] <= TOL
        e.force_kernel(idx, capi.KERNEL_MERGE)
        y2 = np.zeros(rows, np.float32)
        e.run_kernel(x, b, y2, float(ALPHA), float(BETA))
        y64, scale = ol.spmv_f64(rp, ci, vv, x, b, ALPHA, BETA)
        assert ol.max_scaled_error(y2, y64, scale)[0] <= TOL
        # device-pointer calls belong to the per-GPU contexts
        t = torch.zeros(cols, device="cuda")
        with pytest.raises(HispmvError, match="multi-GPU handle"):
            e.run_dev(idx, t, None, torch.zeros(rows, device="cuda"), 1.0, 0.0)
        # memory full: -1, and no child keeps a half-added matrix
        lib.hispmv_set_memory_limit(e._ctx, 1 << 20)
        assert e.create_sparse_handle(r, c, v, rows, cols) == -1
        assert e.num_matrices() == 2
        for k in range(3):
            assert lib.hispmv_num_matrices(C.c_void_p(lib.hispmv_multi_child(e._ctx, k))) == 2
    finally:
        e.close()


@pytest.mark.parametrize("graph", [False, True])
def test_device_chain_matches_cpu_model(eng, graph):
    import torch
    from hispmv_b200.layers import DeviceChain
    cpu_model = _mlp(1)
    chain = DeviceChain(eng, [cpu_model.dense, cpu_model.sparse1, cpu_model.sparse2], relu=[True, True, True], graph=graph)
    for _ in range(3):   # replays must not depend on leftover state
        x = torch.randn(512)
        with torch.no_grad():
            ref = cpu_model(x.view(1, -1)).numpy().reshape(-1)
        out = chain.forward(x.cuda()).cpu().numpy()
        assert np.allclose(out, ref, rtol=1e-3, atol=1e-4)


def test_device_chain_batched_forward(eng):
    """DeviceChain.forward_batch: a (batch, in) block through the three layers, eight vectors per pass over each matrix,
    against the CPU model and against the vector-by-vector chain."""
    import torch
    from hispmv_b200.layers import DeviceChain
    model = _mlp()
    chain = DeviceChain(eng, [model.dense, model.sparse1, model.sparse2], relu=[True, True, True])
    for batch in (1, 3, 8, 13):
        xb = torch.randn((batch, 512))
        with torch.no_grad():
            ref = model(xb).numpy()
        out = chain.forward_batch(xb.cuda()).cpu().numpy()
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        one = np.stack([chain.forward(xb[i].cuda()).cpu().numpy() for i in range(batch)])
        assert np.abs(out - one).max() <= 1e-4 * max(1.0, np.abs(one).max())


def _run_pinned(eng, x, y0, alpha, beta):
    """hispmv_run with PINNED host buffers (the range-pipelined path); plain numpy arrays are pageable and take the
    staged path."""
    import ctypes as C
    import torch
    from hispmv_b200.capi import lib, check
    xp, bp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y0).pin_memory()
    yp = torch.full((y0.size,), float("nan")).pin_memory()
    check(lib.hispmv_run(eng._ctx, C.c_void_p(xp.data_ptr()), C.c_void_p(bp.data_ptr()), C.c_void_p(yp.data_ptr()),
                         float(alpha), float(beta)), "hispmv_run")
    return yp.numpy().copy()


def test_host_run_pipelines_row_ranges(eng):
    """hispmv_run on a matrix with > 2^20 rows: from pinned memory the rows are cut into ranges at tile boundaries and
    bias upload, kernel and y download of different ranges overlap; from pageable memory (plain numpy arrays, the way
    the plugin is called) the vectors go through the context's pinned ring.  Both give the device-resident single
    launch's result, bit for bit."""
    import torch
    from hispmv_b200 import synth
    spec = synth.c2_powerlaw(0.13)                       # 1.3 M rows, ~13 M nnz, rows of up to 520 k nonzeros
    d = synth.DeviceCSR(spec)
    idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
    d.close()
    info = eng.matrix_info(idx)
    assert info["kernel_name"] == "adaptive" and info["num_split_rows"] > 0 and info["num_tiles"] >= 64
    x, y0 = synth.reference_vectors(spec.rows, spec.cols)
    y = np.full(spec.rows, np.nan, np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(x, y0, y, float(ALPHA), float(BETA))
    yd = torch.empty(spec.rows, device="cuda")
    eng.run_dev(idx, torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda(), yd, float(ALPHA), float(BETA),
                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(y.view(np.uint32), yd.cpu().numpy().view(np.uint32))
    assert np.array_equal(_run_pinned(eng, x, y0, ALPHA, BETA).view(np.uint32), y.view(np.uint32))
    n = 300000
    rp, ci, vv = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, n)
    y64, scale = ol.spmv_f64(rp, ci, vv, x, y0[:n], ALPHA, BETA)
    err, at = ol.max_scaled_error(y[:n], y64, scale)
    assert err <= TOL, (err, at)


def test_small_calls_are_one_graph_launch_and_stay_exact(eng, monkeypatch):
    """A DNN layer's vectors (tens of KB) go through one pinned block and one stream synchronisation; with
    HISPMV_SMALL_GRAPH=1, from the second call on, run_kernel / linear replay one captured graph (H2D, kernel, D2H).
    Results are bit-identical to the eager sequence (the default: checked in a second engine), for changing x / bias,
    two alpha-beta pairs and a re-planned matrix."""
    from hispmv_b200 import Engine, capi
    monkeypatch.setenv("HISPMV_SMALL_GRAPH", "1")
    rng = np.random.default_rng(31)
    rows, cols = 3000, 5000
    r, c, v = _matrix(rng, "powerlaw", rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    rp, ci, vv = eng.plan_csr(idx)
    eng.select_matrix(idx)
    outs = []
    for k in range(5):
        x = rng.standard_normal(cols).astype(np.float32)
        b = rng.standard_normal(rows).astype(np.float32)
        for alpha, beta in ((0.85, -2.06), (1.0, 0.0)):
            y = np.full(rows, np.nan, np.float32)
            eng.run_kernel(x, b, y, alpha, beta)
            y64, scale = ol.spmv_f64(rp, ci, vv, x, b, np.float32(alpha), np.float32(beta))
            assert ol.max_scaled_error(y, y64, scale)[0] <= TOL
            outs.append((x, b, alpha, beta, y))
        yl = eng.linear(idx, x, b)
        y64, scale = ol.spmv_f64(rp, ci, vv, x, b, 1.0, 1.0)
        assert ol.max_scaled_error(yl, y64, scale)[0] <= TOL
        if k == 2:
            eng.force_kernel(idx, capi.KERNEL_MERGE)     # the captured launches of the old plan are dropped
    monkeypatch.delenv("HISPMV_SMALL_GRAPH")              # the default: every call eager
    e2 = Engine(0)
    try:
        i2 = e2.create_sparse_handle(r, c, v, rows, cols)
        e2.select_matrix(i2)
        x, b, alpha, beta, y = outs[0]
        y2 = np.zeros(rows, np.float32)
        e2.run_kernel(x, b, y2, alpha, beta)
        assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))
    finally:
        e2.close()


def test_run_xdev_matches_run_and_respects_the_x_stream(eng, monkeypatch):
    """hispmv_run_xdev: x already in HBM, produced on another stream; host bias and y.  Same bits as hispmv_run, also
    with uneven row-range shares (HISPMV_RUN_SHARES) and when x is still being written when the call is made."""
    import ctypes as C
    import torch
    from hispmv_b200 import synth
    from hispmv_b200.capi import lib, check
    spec = synth.c2_powerlaw(0.11)
    d = synth.DeviceCSR(spec)
    idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
    d.close()
    x, y0 = synth.reference_vectors(spec.rows, spec.cols)
    y_ref = np.full(spec.rows, np.nan, np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(x, y0, y_ref, float(ALPHA), float(BETA))
    side = torch.cuda.Stream()
    xh = torch.from_numpy(x).pin_memory()
    bh = torch.from_numpy(y0).pin_memory()
    for shares in (None, "5,1,1,3,1,1,1,1,1,2,1"):
        if shares:
            monkeypatch.setenv("HISPMV_RUN_SHARES", shares)
        xd = torch.zeros(spec.cols, device="cuda")
        yh = torch.full((spec.rows,), float("nan")).pin_memory()
        with torch.cuda.stream(side):
            torch.cuda._sleep(20_000_000)                 # x arrives late: the call must wait for the side stream
            xd.copy_(xh, non_blocking=True)
        check(lib.hispmv_run_xdev(eng._ctx, C.c_void_p(xd.data_ptr()), C.c_void_p(side.cuda_stream),
                                  C.c_void_p(bh.data_ptr()), C.c_void_p(yh.data_ptr()), float(ALPHA), float(BETA)),
              "run_xdev")
        assert np.array_equal(yh.numpy().view(np.uint32), y_ref.view(np.uint32)), shares
    monkeypatch.delenv("HISPMV_RUN_SHARES", raising=False)
    st = lib.hispmv_run_xdev(eng._ctx, C.c_void_p(0), C.c_void_p(0), C.c_void_p(bh.data_ptr()),
                             C.c_void_p(yh.data_ptr()), 1.0, 0.0)
    assert st != 0                                         # null x is refused


def test_multicast_y_is_refused_where_y_is_read_back(eng):
    """hispmv_run_dev_mc: strategies that read y back (merge-path fix-up) cannot write through a write-only multicast
    address and must say so before launching anything.  (The store path itself needs two GPUs: test_multi_gpu.py.)"""
    import ctypes as C
    import torch
    from hispmv_b200 import capi
    from hispmv_b200.capi import lib
    rng = np.random.default_rng(2)
    r, c, v = _rand_coo(rng, 3000, 2000, 40000)
    idx = eng.create_sparse_handle(r, c, v, 3000, 2000)
    eng.force_kernel(idx, capi.KERNEL_MERGE)
    x = torch.zeros(2000, device="cuda")
    y = torch.full((3000,), 7.0, device="cuda")
    st = lib.hispmv_run_dev_mc(eng._ctx, idx, C.c_void_p(x.data_ptr()), C.c_void_p(0), C.c_void_p(y.data_ptr()), 1.0,
                               0.0, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert st == -5 and b"multicast" in lib.hispmv_last_error()           # HISPMV_ERR_STATE, nothing launched
    assert bool((y == 7.0).all())
    st = lib.hispmv_run_dev_mc(eng._ctx, idx, C.c_void_p(x.data_ptr()), C.c_void_p(0), C.c_void_p(0), 1.0, 0.0, 0,
                               C.c_void_p(0))
    assert st != 0                                                         # null multicast address


@pytest.mark.parametrize("shape", ["dnn", "short_rows", "with_empty_rows", "few_long_rows"])
def test_batched_vectors_one_pass(eng, shape):
    """Several right-hand sides per pass (hispmv_run_dev_batch, and linear() with num_vecs >= 2): every vector within
    the north_star bar of the float64 oracle, for 1..11 vectors (groups of 8 / 4 / 2 and a single left over), with
    alpha/beta and the fused ReLU on the device call."""
    import torch
    rng = np.random.default_rng({"dnn": 1, "short_rows": 2, "with_empty_rows": 3, "few_long_rows": 4}[shape])
    if shape == "dnn":
        rows, cols = 2048, 1536
        w = rng.standard_normal((rows, cols)).astype(np.float32) * (rng.random((rows, cols)) < 0.1)
        r, c = np.nonzero(w)
        v = w[r, c]
    elif shape == "few_long_rows":                       # one CTA per row (spmm_csr_cta_kernel)
        rows, cols = 300, 9000
        lens = rng.integers(2500, 3500, rows)
        lens[5] = 0
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
    elif shape == "short_rows":
        rows, cols = 30000, 5000
        r, c, v = _rand_coo(rng, rows, cols, 90000, dup=0.02)
    else:
        rows, cols = 5000, 7000
        lens = np.minimum(rng.zipf(1.6, rows), 4000)
        lens[rng.integers(0, rows, 1500)] = 0
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
    r, c, v = r.astype(np.int32), c.astype(np.int32), v.astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    rp, ci, vv = eng.plan_csr(idx)
    bias = rng.standard_normal(rows).astype(np.float32)
    for nv in (1, 2, 3, 5, 8, 11):
        X = rng.standard_normal((nv, cols)).astype(np.float32)
        # host path: linear() = alpha = beta = 1 (fpga_handle.cpp:351-352), returns num_vecs * rows values
        Y = eng.linear(idx, X.reshape(-1), bias).reshape(nv, rows)
        for k in range(nv):
            y64, scale = ol.spmv_f64(rp, ci, vv, X[k], bias, 1.0, 1.0)
            err, at = ol.max_scaled_error(Y[k], y64, scale)
            assert err <= TOL, (shape, nv, k, err, at)
        # device path with alpha / beta / ReLU
        Xd, bd = torch.from_numpy(X).cuda(), torch.from_numpy(bias).cuda()
        Yd = torch.full((nv, rows), float("nan"), device="cuda")
        eng.run_dev_batch(idx, Xd, bd, Yd, float(ALPHA), float(BETA), relu=True,
                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        Yh = Yd.cpu().numpy()
        for k in range(nv):
            y64, scale = ol.spmv_f64(rp, ci, vv, X[k], bias, ALPHA, BETA)
            pos = y64 > 1e-4 * np.maximum(scale, 1e-30)
            err, _ = ol.max_scaled_error(Yh[k][pos], y64[pos], scale[pos])
            assert err <= TOL and np.all(Yh[k] >= 0)
            assert np.all(Yh[k][y64 < -1e-4 * np.maximum(scale, 1e-30)] == 0)


@pytest.mark.parametrize("rows,cols", [(1, 1), (257, 1031), (1024, 4096), (37, 70001)])
def test_batched_vectors_dense_one_pass(eng, rows, cols):
    """The dense overlay with several vectors (gemm_lite_kernel): every vector within the bar of the float64 oracle,
    ragged shapes (cols not a multiple of 4, odd row counts), host and device calls."""
    import torch
    rng = np.random.default_rng(rows * 7 + cols)
    a = rng.standard_normal((rows, cols)).astype(np.float32)
    idx = eng.create_dense_handle(a.reshape(-1), rows, cols)
    bias = rng.standard_normal(rows).astype(np.float32)
    for nv in (2, 3, 8, 9):
        X = rng.standard_normal((nv, cols)).astype(np.float32)
        Y = eng.linear(idx, X.reshape(-1), bias).reshape(nv, rows)
        Xd, bd = torch.from_numpy(X).cuda(), torch.from_numpy(bias).cuda()
        Yd = torch.full((nv, rows), float("nan"), device="cuda")
        eng.run_dev_batch(idx, Xd, bd, Yd, float(ALPHA), float(BETA), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        Yh = Yd.cpu().numpy()
        for k in range(nv):
            y64, scale = ol.gemv_f64(a, rows, cols, X[k], bias, 1.0, 1.0)
            assert ol.max_scaled_error(Y[k], y64, scale)[0] <= TOL, (nv, k)
            y64, scale = ol.gemv_f64(a, rows, cols, X[k], bias, ALPHA, BETA)
            assert ol.max_scaled_error(Yh[k], y64, scale)[0] <= TOL, (nv, k)


def test_batch_rows_a_sub_warp_cannot_walk_get_a_cta_each(eng):
    """Rows above 65536 nonzeros stay in the one-pass batch: the sub-warp kernel skips them and one CTA per listed row
    takes them (hispmv_run_dev_batch / linear with several vectors): same contract, five vectors in one pass."""
    rng = np.random.default_rng(9)
    rows, cols = 300, 200000
    lens = np.full(rows, 50)
    lens[7] = 70000
    lens[299] = 140000
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    assert eng.matrix_info(idx)["max_row_nnz"] == 140000
    rp, ci, vv = eng.plan_csr(idx)
    bias = rng.standard_normal(rows).astype(np.float32)
    X = rng.standard_normal((5, cols)).astype(np.float32)
    Y = eng.linear(idx, X.reshape(-1), bias).reshape(5, rows)
    for k in range(5):
        y64, scale = ol.spmv_f64(rp, ci, vv, X[k], bias, 1.0, 1.0)
        assert ol.max_scaled_error(Y[k], y64, scale)[0] <= TOL


def test_column_slabs_bit_exact_and_within_tolerance(eng, monkeypatch):
    """x larger than L2: the matrix is cut into column slabs (CSR over the same rows, bit-exact against a numpy
    restatement), one launch per slab, y accumulating across them; bias / ReLU applied exactly once."""
    import torch
    from hispmv_b200 import capi
    monkeypatch.setenv("HISPMV_SLAB_COLS", "30016")
    rng = np.random.default_rng(77)
    rows, cols = 50000, 100003
    lens = np.minimum(rng.zipf(1.7, rows), 30000)
    lens[rng.integers(0, rows, 4)] = 9000
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
    info = eng.matrix_info(idx)
    assert (info["num_slabs"], info["slab_cols"]) == (4, 30016)
    rp, ci, vv = eng.plan_csr(idx)
    total = 0
    for sidx in range(4):
        srp, sci, svv = eng.plan_slab_csr(idx, sidx)
        srp2, sci2, svv2 = ol.column_slab(rp, ci, vv, sidx * 30016, min((sidx + 1) * 30016, cols))
        assert np.array_equal(srp, srp2) and np.array_equal(sci, sci2)
        assert np.array_equal(svv.view(np.uint32), svv2.view(np.uint32))
        total += sci.size
    assert total == ci.size
    y1 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(3))
    y2 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(3))
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))
    _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(4), alpha=1.0, beta=0.0)
    # fused ReLU only after the last slab
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    yd = torch.empty(rows, device="cuda")
    eng.linear_dev(idx, torch.from_numpy(x).cuda(), torch.from_numpy(b).cuda(), yd, relu=True,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    y64, scale = ol.spmv_f64(rp, ci, vv, x, b, 1.0, 1.0)
    y = yd.cpu().numpy()
    pos = y64 > 1e-4 * np.maximum(scale, 1e-30)
    err, _ = ol.max_scaled_error(y[pos], y64[pos], scale[pos])
    assert err <= TOL and np.all(y >= 0) and np.all(y[y64 < -1e-4 * np.maximum(scale, 1e-30)] == 0)
    # linear() through the host path
    out = eng.linear(idx, x, b)
    err, _ = ol.max_scaled_error(out, y64, scale)
    assert err <= TOL


def test_slab_selector_bit_exact(eng):
    """select_slab_cols: x of 80 MB with random columns -> two slabs of 10 M columns; the same matrix with 16 M
    columns (64 MB) stays whole."""
    rng = np.random.default_rng(9)
    rows, nnz = 300000, 4200000
    for cols, want in ((20000000, 2), (16000000, 0)):
        r = rng.integers(0, rows, nnz).astype(np.int32)
        r[:50000] = 11                                              # a heavy row
        c = rng.integers(0, cols, nnz).astype(np.int32)
        v = rng.standard_normal(nnz).astype(np.float32)
        idx = eng.create_sparse_handle(r, c, v, rows, cols)
        info = eng.matrix_info(idx)
        w = ol.select_slab_cols(cols, nnz, info["probe_near"], info["probe_cmp"])
        assert info["kernel_name"] == "adaptive" and info["slab_cols"] == w and info["num_slabs"] == want
        if want:
            assert w == 10000000 and (info["num_slabs"] - 1) * w < cols <= info["num_slabs"] * w
        rp, ci, vv = eng.plan_csr(idx)
        _check_run(eng, idx, rp, ci, vv, rows, cols, rng)
