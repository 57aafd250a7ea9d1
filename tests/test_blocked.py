"""The blocked (column-slab x row-panel, two-pass) strategy: plan bit-exact against the oracle's sequential
restatement, results within the north_star tolerance of the float64 restatement, bit-reproducible runs, the
host-buffer pipeline and the selector rule.  CPU part: the oracle's plan, walked the way the two passes walk it,
reproduces the CSR product (so the restatement itself is pinned to the CSR semantics the other tests pin)."""
import numpy as np
import pytest

import oracle_lib as ol

TOL = 1e-5
ALPHA, BETA = np.float32(0.85), np.float32(-2.06)


def _powerlaw(rng, rows, cols, heavy=3, max_len=20000):
    lens = np.minimum(rng.zipf(1.6, rows), max_len)
    if heavy:
        lens[rng.integers(0, rows, heavy)] = rng.integers(max_len // 4, max_len, heavy)
    lens[rng.integers(0, rows, rows // 10)] = 0
    lens[rows // 3: rows // 3 + 3000] = 0        # more empty rows in a row than a panel's item budget
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    # power-law columns (a hot head and a thin tail, as in C2) so that slabs differ in size by orders of magnitude
    c = np.minimum((rng.random(r.size) ** 5 * cols).astype(np.int32), cols - 1)
    v = rng.standard_normal(r.size).astype(np.float32)
    return r, c, v


def _walk_plan(d, rows, CH, W, cols, x):
    """numpy emulation of pass 1 + pass 2 over a plan dict: returns A x in float64."""
    G = 512
    slab_of = np.zeros(d["padded_nnz"], np.int64)
    for s in range(d["num_slabs"]):
        slab_of[d["slab_ptr"][s]:d["slab_ptr"][s + 1]] = s
    # storage inside a group: entry e = 16*lane + w sits where coalesced vector loads hand lane `lane` its 16 entries
    k = np.arange(d["padded_nnz"], dtype=np.int64)
    g, e = k // G, k % G
    lane, w = e // 16, e % 16
    val = d["val"][g * G + ((w // 4) * 32 + lane) * 4 + w % 4]
    lcol = d["lcol"][g * G + ((w // 8) * 32 + lane) * 8 + w % 8]
    prod = val.astype(np.float64) * x[np.minimum(slab_of * W + lcol, cols - 1)]
    # pass 1: one partial per piece; pieces never cross a group of G entries
    fl = np.unpackbits(d["flags"].view(np.uint8)[:, None], axis=1, bitorder="little").reshape(-1).astype(bool)
    part = np.zeros(d["num_pieces"])
    q, run = 0, 0.0
    for k in range(d["padded_nnz"]):
        if k % G == 0:
            assert d["group_base"][k // G] == q
            run = 0.0
        run += prod[k]
        if fl[k]:
            part[q] = run
            q, run = q + 1, 0.0
    assert q == d["num_pieces"] == d["group_base"][-1]
    # pass 2: panels over the per-row piece counts
    y = np.zeros(rows)
    tr, tc, tn, pr = d["tile_row"], d["tile_chunk"], d["tile_first"], d["prow_ptr"]
    for p in range(d["num_panels"]):
        n0 = int(tn[p])
        n1 = min(int(pr[tr[p] + 1]), n0 + CH) if tc[p] >= 0 else int(pr[tr[p + 1]])
        n = n1 - n0
        buf = np.full(n, np.nan)
        segs = d["seg"][d["panel_seg"][p]:d["panel_seg"][p + 1]]
        offs = list(segs[:, 1]) + [n]
        for i, (st, off) in enumerate(segs):
            ln = offs[i + 1] - off
            assert ln > 0
            buf[d["perm"][st:st + ln]] = part[st:st + ln]
        assert not np.isnan(buf).any()           # every piece slot of the panel is filled exactly once
        if tc[p] >= 0:
            y[tr[p]] += buf.sum()
            assert np.all(d["seg_copy"][d["panel_seg"][p]:d["panel_seg"][p + 1], 1] == 0)
            continue
        # the way the kernel does it: the aligned quads of partial sums that cover the panel's segments (chunk_src; the
        # same ranges per segment in seg_copy), every position dropped at its slot (perm2), rows closed at the end marks
        pos0, bit0 = d["panel_aux"][p]
        pos1, bit1 = d["panel_aux"][p + 1]
        stage = np.full(pos1 - pos0, np.nan)
        padded = np.concatenate([part, np.full(8, np.nan)])      # the quads may read past the last piece
        for q, a0 in enumerate(d["chunk_src"][pos0 // 4:pos1 // 4]):
            assert a0 % 4 == 0
            stage[4 * q:4 * q + 4] = padded[a0:a0 + 4]
        for a0, packed in d["seg_copy"][d["panel_seg"][p]:d["panel_seg"][p + 1]]:
            so, ln = 4 * (packed & 0xFFFF), 4 * ((packed >> 16) & 0xFFFF)
            assert a0 % 4 == 0 and ln > 0
            assert np.array_equal(stage[so:so + ln], padded[a0:a0 + ln], equal_nan=True)
        slots = np.full(n, np.nan)
        pl = d["perm2"][pos0:pos1]
        keep = pl != 0xFFFF
        assert keep.sum() == n and len(set(pl[keep])) == n
        slots[pl[keep]] = stage[keep]
        assert np.array_equal(slots, buf)
        words = d["end_bits"][bit0:bit1]
        assert words.size == (n + 31) // 32
        ends = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool) if n else np.zeros(0, bool)
        want_ends = np.zeros(n, bool)
        for rr in range(tr[p], tr[p + 1]):
            y[rr] = buf[pr[rr] - n0:pr[rr + 1] - n0].sum()
            if pr[rr + 1] > pr[rr]:
                want_ends[pr[rr + 1] - 1 - n0] = True
        assert np.array_equal(ends, want_ends)
    return y


@pytest.mark.parametrize("seed,W,B,T,CH", [(0, 1024, 2048, 256, 512), (1, 4096, 512, 64, 128), (2, 20000, 4096, 1024, 4096)])
def test_oracle_blocked_plan_reproduces_the_csr_product(seed, W, B, T, CH):
    rng = np.random.default_rng(seed)
    rows, cols = 4000, 20000 + seed
    r, c, v = _powerlaw(rng, rows, cols, max_len=6000)
    rp, ci, vv = ol.coo_to_csr(rows, r, c, v)
    d = ol.pb_plan(rp, ci, vv, cols, B, T, CH, W, n_cta=7, slab_cost=300, piece_cost16=24)
    x = rng.standard_normal(cols).astype(np.float32)
    y = _walk_plan(d, rows, CH, W, cols, x)
    y64, scale = ol.spmv_f64(rp, ci, vv, x)
    assert np.max(np.abs(y - y64) / np.maximum(scale, 1e-30)) < 1e-12
    # layout facts: slab starts aligned, ascending; lcol inside the slab; work ranges tile the blocked order
    assert np.all(d["slab_ptr"] % 512 == 0) and np.all(np.diff(d["slab_ptr"]) >= 0)
    assert d["lcol"].max() < W and d["num_pieces"] <= ci.size and np.all(np.diff(d["prow_ptr"]) >= 0)
    assert np.all((np.diff(d["prow_ptr"]) > 0) == (np.diff(rp) > 0))     # a row has pieces iff it has nonzeros
    w = d["work"]
    assert w[0, 0] == 0 and w[-1, 1] == d["padded_nnz"] and np.array_equal(w[1:, 0], w[:-1, 1])
    assert np.all(w % 512 == 0)


def test_oracle_select_blocked_rule():
    assert ol.select_blocked(10_000_000, 10_000_000, 100_000_000, 30_000_000, 10, 1000) == 1    # C2: 0.3 runs per nonzero
    assert ol.select_blocked(10_000_000, 10_000_000, 100_000_000, 60_000_000, 10, 1000) == 1
    assert ol.select_blocked(10_000_000, 10_000_000, 100_000_000, 60_000_001, 10, 1000) == 0
    assert ol.select_blocked(100_000_000, 100_000_000, 1_000_000_000, 999_000_000, 0, 1000) == 0  # C5: every nonzero a run
    assert ol.select_blocked(20_000_000, 20_000_000, 540_000_000, 20_000_000, 990, 1000) == 0   # C4: banded
    assert ol.select_blocked(65536, 65536, 1_000_000, 100_000, 10, 1000) == 0                   # C1: small
    assert ol.select_blocked(8192, 8192, 6_700_000, 8192, 10, 1000) == 0                        # C3b: x fits L1
    assert ol.select_blocked(10_000_000, 99_999, 100_000_000, 20_000_000, 10, 1000) == 0        # x below 100 000 columns
    assert ol.select_blocked(10_000_000, 100_000, 100_000_000, 20_000_000, 10, 1000) == 1
    assert ol.select_blocked(10_000_000, 10_000_000, 100_000_000, 30_000_000, 10, 1000, 0) == 0  # rows may not be split


def test_oracle_run_count():
    rp = np.array([0, 3, 3, 7], np.int32)
    ci = np.array([1, 2, 50000, 0, 49151, 49152, 98304], np.int32)
    assert ol.pb_count_runs(rp, ci, 49152) == 2 + 3
    assert ol.pb_count_runs(rp, ci, 4) == 2 + 4


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from hispmv_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


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


@pytest.mark.gpu
@pytest.mark.parametrize("seed,params", [(0, "1024,2048,256,512,300,0"), (1, "4096,512,64,128,0,16"),
                                         (2, "20000,4096,1024,4096,5000,40"), (3, "49152,8192,4096,8192,32768,8")])
def test_blocked_plan_bit_exact(eng, seed, params, monkeypatch):
    """Every integer artefact of the blocked plan (slab starts, blocked order through val / lcol / perm, segment table,
    pass-1 work ranges) equals the sequential restatement; then the run is within tolerance and bit-reproducible."""
    from hispmv_b200 import capi
    monkeypatch.setenv("HISPMV_BLOCKED", params)
    W, B, T, CH, cost, pcost = (int(t) for t in params.split(","))
    rng = np.random.default_rng(seed)
    rows, cols = 30000, 200003 + seed
    r, c, v = _powerlaw(rng, rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_BLOCKED)
    info = eng.matrix_info(idx)
    assert info["kernel_name"] == "blocked" and (info["tile_items"], info["long_threshold"], info["chunk_nnz"]) == (B, T, CH)
    rp, ci, vv = eng.plan_csr(idx)
    got = eng.plan_blocked(idx)
    want = ol.pb_plan(rp, ci, vv, cols, B, T, CH, W, n_cta=got["num_work"], slab_cost=cost, piece_cost16=pcost)
    for k in ("slab_cols", "num_slabs", "padded_nnz", "num_pieces", "num_seg", "num_panels", "num_chunks"):
        assert got[k] == want[k], k
    for k in ("stage_total", "bit_words"):
        assert got[k] == want[k], k
    for k in ("slab_ptr", "lcol", "flags", "group_base", "prow_ptr", "perm", "panel_seg", "seg", "panel_chunk", "chunk",
              "seg_copy", "perm2", "chunk_src", "panel_aux", "end_bits", "work"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["val"].view(np.uint32), want["val"].view(np.uint32))
    tr, tn = eng.plan_tiles(idx)        # the panels: adaptive tiles over the per-row piece counts
    assert np.array_equal(tr, want["tile_row"]) and np.array_equal(tn, want["tile_first"])
    assert np.array_equal(eng.plan_tile_chunks(idx), want["tile_chunk"])
    assert np.array_equal(eng.plan_split_rows(idx), want["split_rows"])
    assert eng.launches_per_run(idx) == 2
    y1 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))
    y2 = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(7))
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))
    _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(8), alpha=1.0, beta=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["powerlaw", "regular", "short", "hollow", "one_row", "one_col"])
def test_blocked_within_tolerance(eng, kind, monkeypatch):
    from hispmv_b200 import capi
    monkeypatch.setenv("HISPMV_BLOCKED", "8192,4096,512,1024")
    rng = np.random.default_rng(sum(map(ord, kind)))
    rows, cols = 30011, 100003
    if kind == "powerlaw":
        r, c, v = _powerlaw(rng, rows, cols)
    else:
        if kind == "regular":
            lens = np.full(rows, 27)
        elif kind == "short":
            lens = rng.integers(0, 3, rows)
        elif kind == "hollow":
            lens = np.where(rng.random(rows) < 0.8, 0, rng.integers(1, 40, rows))
        elif kind == "one_row":
            rows = 1
            lens = np.array([70000])
        else:
            cols = 1
            lens = rng.integers(0, 2, rows)
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_BLOCKED)
    assert eng.matrix_info(idx)["kernel_name"] == "blocked"
    rp, ci, vv = eng.plan_csr(idx)
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng)
    _check_run(eng, idx, rp, ci, vv, rows, cols, rng, alpha=1.0, beta=0.0)


@pytest.mark.gpu
def test_blocked_device_calls_relu_linear_and_unaligned_x(eng, monkeypatch):
    """run_dev / linear_dev(+ReLU) / linear() on a blocked matrix; x at an address that is not 16-byte aligned takes
    the plain-load staging path and gives the same bits."""
    import torch
    from hispmv_b200 import capi
    monkeypatch.setenv("HISPMV_BLOCKED", "16384,8192,1024,2048")
    rng = np.random.default_rng(5)
    rows, cols = 40000, 150001
    r, c, v = _powerlaw(rng, rows, cols)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_BLOCKED)
    rp, ci, vv = eng.plan_csr(idx)
    x = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(rows).astype(np.float32)
    s = torch.cuda.current_stream().cuda_stream
    xd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(b).cuda()
    y = torch.empty(rows, device="cuda")
    eng.run_dev(idx, xd, bd, y, float(ALPHA), float(BETA), s)
    torch.cuda.synchronize()
    y64, scale = ol.spmv_f64(rp, ci, vv, x, b, ALPHA, BETA)
    assert ol.max_scaled_error(y.cpu().numpy(), y64, scale)[0] <= TOL
    big = torch.empty(cols + 1, device="cuda")
    big[1:].copy_(xd)
    y_un = torch.empty(rows, device="cuda")
    eng.run_dev(idx, big[1:], bd, y_un, float(ALPHA), float(BETA), s)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int32), y_un.view(torch.int32))
    yr = torch.empty(rows, device="cuda")
    eng.linear_dev(idx, xd, bd, yr, relu=True, stream=s)
    torch.cuda.synchronize()
    y64, scale = ol.spmv_f64(rp, ci, vv, x, b, 1.0, 1.0)
    yh = yr.cpu().numpy()
    pos = y64 > 1e-4 * np.maximum(scale, 1e-30)
    assert ol.max_scaled_error(yh[pos], y64[pos], scale[pos])[0] <= TOL
    assert np.all(yh >= 0) and np.all(yh[y64 < -1e-4 * np.maximum(scale, 1e-30)] == 0)
    xs = np.concatenate([x, x[::-1].copy(), 2 * x])          # three vectors: both stream lanes of linear()
    out = eng.linear(idx, xs, b).reshape(3, rows)
    for k, xv in enumerate((x, x[::-1].copy(), 2 * x)):
        y64, scale = ol.spmv_f64(rp, ci, vv, xv, b, 1.0, 1.0)
        assert ol.max_scaled_error(out[k], y64, scale)[0] <= TOL


@pytest.mark.gpu
def test_blocked_host_run_pipelines_panel_ranges(eng):
    """> 1 M rows: hispmv_run cuts the panels into row ranges (pass 1 once, pass 2 range by range around the bias / y
    copies); the result is bit-identical to the device-resident call."""
    import torch
    from hispmv_b200 import capi
    rng = np.random.default_rng(21)
    rows, cols = 1_300_000, 1_100_000
    lens = rng.integers(0, 7, rows)
    lens[[5, 700_000]] = [40000, 3000]
    r = np.repeat(np.arange(rows, dtype=np.int32), lens)
    c = rng.integers(0, cols, r.size).astype(np.int32)
    v = rng.standard_normal(r.size).astype(np.float32)
    idx = eng.create_sparse_handle(r, c, v, rows, cols)
    eng.force_kernel(idx, capi.KERNEL_BLOCKED)
    info = eng.matrix_info(idx)
    assert info["kernel_name"] == "blocked" and info["num_tiles"] >= 64
    rp, ci, vv = eng.plan_csr(idx)
    y_host = _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(3))
    g = np.random.default_rng(3)
    x = g.standard_normal(cols).astype(np.float32)
    y0 = g.standard_normal(rows).astype(np.float32)
    yd = torch.empty(rows, device="cuda")
    eng.run_dev(idx, torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda(), yd, float(ALPHA), float(BETA),
                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy().view(np.uint32), y_host.view(np.uint32))
    # pinned host buffers: pass 1 once, then pass 2 range by range around the bias / y copies
    import ctypes as C
    from hispmv_b200.capi import lib, check
    xp, bp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y0).pin_memory()
    yp = torch.full((rows,), float("nan")).pin_memory()
    eng.select_matrix(idx)
    check(lib.hispmv_run(eng._ctx, C.c_void_p(xp.data_ptr()), C.c_void_p(bp.data_ptr()), C.c_void_p(yp.data_ptr()),
                         float(ALPHA), float(BETA)), "hispmv_run")
    assert np.array_equal(yp.numpy().view(np.uint32), y_host.view(np.uint32))


@pytest.mark.gpu
def test_blocked_selector_bit_exact(eng):
    """The selector sends a scattered-column matrix with >= 100 000 columns, >= 16 M nonzeros and at most 0.6 (row, slab)
    runs per nonzero to the blocked strategy and leaves the others on the one-pass kernels, exactly as
    oracle_select_blocked says; the run count itself is bit-exact against the oracle's."""
    import torch
    cases = ((400_000, 1_500_000, 16_500_000, 40),     # rows concentrated in 40-column bursts: few runs -> blocked
             (2_000_000, 1_500_000, 16_500_000, 0),    # uniform columns: every nonzero its own run -> one-pass
             (400_000, 90_000, 16_500_000, 40),        # x too small (below 100 000 columns)
             (200_000, 1_500_000, 8_000_000, 40))      # too few nonzeros
    for rows, cols, nnz, burst in cases:
        g = torch.Generator(device="cuda").manual_seed(rows + cols)
        r = torch.randint(0, rows, (nnz,), device="cuda", generator=g, dtype=torch.int32)
        if burst:   # a row's entries sit in a few bursts of `burst` consecutive columns
            start = (r.long() * 7919) % (cols - 64)
            c = (start + torch.randint(0, burst, (nnz,), device="cuda", generator=g)).to(torch.int32)
        else:
            c = torch.randint(0, cols, (nnz,), device="cuda", generator=g, dtype=torch.int32)
        v = torch.randn(nnz, device="cuda", generator=g)
        idx = eng.create_sparse_handle_coo_dev(r, c, v, rows, cols)
        info = eng.matrix_info(idx)
        rp, ci, vv = eng.plan_csr(idx)
        runs = ol.pb_count_runs(rp, ci) if info["slab_runs"] else 0
        assert info["slab_runs"] == runs
        want = ol.select_blocked(rows, cols, nnz, runs, info["probe_near"], info["probe_cmp"]) if runs else 0
        assert (info["kernel_name"] == "blocked") == bool(want), (rows, cols, nnz, info["kernel_name"], runs)
        if want:
            pb = eng.plan_blocked(idx, arrays=False)
            assert pb["slab_cols"] == 49152 and pb["num_slabs"] == (cols + 49151) // 49152
        _check_run(eng, idx, rp, ci, vv, rows, cols, np.random.default_rng(1))
