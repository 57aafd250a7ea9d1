"""BASELINE.json's configurations at FULL size on one B200 (C2 100 M nnz, C4 540 M nnz, C5 1 B nnz, the 8192^2 GeMV).

The CPU oracle cannot walk these in seconds, so the checks are the size-independent ones the domain offers, plus one
that is not size-limited at all:
  * every element of y against a float64 reference accumulated on the device by plain torch (gather + index_add_,
    chunked) with the north_star bar |y - y64| / (|alpha| sum|a_ij x_j| + |beta y0_i|) <= 1e-5; that torch reference
    is itself held to the C oracle (and through it to the reference's arithmetic) on the first rows of the matrix,
    which the oracle regenerates bit-exactly on the CPU;
  * checksum of checksums: sum_i y_i == sum_k a_k x[col_k] (+ beta sum y0) in float64;
  * linearity: A(2 xa - 3 xb) == 2 A xa - 3 A xb within the same bar;
  * determinism: two runs are bit-identical; a second strategy (merge-path) agrees within the bar;
  * integer artefacts: row_ptr equals the prefix sum of the generator's row lengths, tiles cover the nonzeros in
    order, the split-row list is sorted and holds only rows at or above the long-row threshold.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

TOL = 1e-5
ALPHA, BETA = float(np.float32(0.85)), float(np.float32(-2.06))   # what the C-ABI's float arguments hold
CHUNK = 1 << 27          # nonzeros per torch reference chunk (keeps the float64 temporaries near 4 GB)
SAMPLE_ROWS = 100_000    # rows the C oracle re-generates and re-computes on the CPU


class _DevArray:
    """A raw device pointer as something torch.as_tensor can wrap without copying."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def _wrap(ptr, n, typestr):
    import torch
    return torch.as_tensor(_DevArray(ptr, n, typestr), device="cuda")


def _torch_reference(rp, col, val, x, y0, alpha, beta):
    """float64 y and the normaliser, both on the device, by plain torch ops."""
    import torch
    rows = rp.numel() - 1
    lens = (rp[1:] - rp[:-1]).long()
    acc = torch.zeros(rows, dtype=torch.float64, device="cuda")
    mag = torch.zeros(rows, dtype=torch.float64, device="cuda")
    rp64 = rp.long()
    nnz = int(rp64[-1])
    r0 = 0
    while r0 < rows:
        # rows [r0, r1) with at most CHUNK nonzeros (a single longer row goes alone)
        target = int(rp64[r0]) + CHUNK
        r1 = int(torch.searchsorted(rp64, torch.tensor([target], device="cuda"), right=True)[0]) - 1
        r1 = min(max(r1, r0 + 1), rows)
        k0, k1 = int(rp64[r0]), int(rp64[r1])
        if k1 > k0:
            row_of = torch.repeat_interleave(torch.arange(r0, r1, device="cuda"), lens[r0:r1])
            p = val[k0:k1].double() * x[col[k0:k1].long()].double()
            acc.index_add_(0, row_of, p)
            mag.index_add_(0, row_of, p.abs())
            del row_of, p
        r0 = r1
    assert nnz == int(lens.sum())
    y64 = alpha * acc + beta * y0.double()
    scale = abs(alpha) * mag + (abs(beta) * y0.double().abs())
    return y64, scale


def _max_err(y, y64, scale):
    import torch
    s = torch.where(scale > 0, scale, torch.ones_like(scale))
    e = (y.double() - y64).abs() / s
    e = torch.where((scale == 0) & (y.double() == y64), torch.zeros_like(e), e)
    return float(e.max())


def _run(eng, idx, x, b, y, alpha=ALPHA, beta=BETA):
    import torch
    eng.run_dev(idx, x, b, y, alpha, beta, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


@pytest.mark.parametrize("which", ["c2", "c4", "c5"])
def test_full_size_spmv_properties(which):
    import torch
    from hispmv_b200 import Engine, capi, synth
    spec = {"c2": synth.c2_powerlaw, "c4": synth.c4_stencil, "c5": synth.c5_uniform}[which](1.0)
    d = synth.DeviceCSR(spec)
    rp = _wrap(d.row_ptr, spec.rows + 1, "<i4")
    col = _wrap(d.col, d.nnz, "<i4")
    val = _wrap(d.val, d.nnz, "<f4")
    eng = Engine(0)
    try:
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        info = eng.matrix_info(idx)
        assert info["nnz"] == d.nnz and info["rows"] == spec.rows
        # C2's rows are concentrated (0.3 (row, slab) runs per nonzero): the two-pass blocked strategy; C5's uniform
        # columns make every nonzero its own run: one-pass kernel over column slabs; C4 is banded: row-major staging
        expect_kernel = {"c2": "blocked", "c4": "rowstage", "c5": "adaptive"}[which]
        assert info["kernel_name"] == expect_kernel
        if which == "c5":
            assert info["num_slabs"] >= 2          # x (400 MB) does not fit L2: column slabs
        xh, y0h = synth.reference_vectors(spec.rows, spec.cols)
        g = torch.Generator(device="cuda").manual_seed(7)
        xa = torch.from_numpy(xh).cuda() * (torch.rand(spec.cols, device="cuda", generator=g) - 0.5)
        y0 = torch.from_numpy(y0h).cuda()
        y = torch.full((spec.rows,), float("nan"), device="cuda")
        _run(eng, idx, xa, y0, y)

        # --- every element against the float64 torch reference ---
        y64, scale = _torch_reference(rp, col, val, xa, y0, ALPHA, BETA)
        assert not bool(torch.isnan(y).any())
        assert _max_err(y, y64, scale) <= TOL

        # --- the torch reference against the C oracle on the first rows (generator restated on the CPU) ---
        n = SAMPLE_ROWS
        rph, cih, vvh = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, n)
        assert np.array_equal(rph, rp[:n + 1].cpu().numpy())
        k = int(rph[-1])
        assert np.array_equal(cih, col[:k].cpu().numpy())
        assert np.array_equal(vvh.view(np.uint32), val[:k].cpu().numpy().view(np.uint32))
        o64, oscale = ol.spmv_f64(rph, cih, vvh, xa.cpu().numpy(), y0h[:n], np.float32(ALPHA), np.float32(BETA))
        assert np.allclose(o64, y64[:n].cpu().numpy(), rtol=0, atol=1e-9 * max(1.0, float(oscale.max())))
        err, at = ol.max_scaled_error(y[:n].cpu().numpy(), o64, oscale)
        assert err <= TOL, (err, at)

        # --- checksum of checksums ---
        total = float(y.double().sum())
        want = float(y64.sum())
        assert abs(total - want) <= TOL * float(scale.sum())

        # --- determinism ---
        y2 = torch.empty_like(y)
        _run(eng, idx, xa, y0, y2)
        assert torch.equal(y.view(torch.int32), y2.view(torch.int32))

        # --- linearity (alpha = 1, beta = 0; bias unused) ---
        xb = torch.rand(spec.cols, device="cuda", generator=g) + 0.5
        ya, yb, yc = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
        _run(eng, idx, xa, y0, ya, 1.0, 0.0)
        _run(eng, idx, xb, y0, yb, 1.0, 0.0)
        xc = 2.0 * xa - 3.0 * xb
        _run(eng, idx, xc, y0, yc, 1.0, 0.0)
        _, s_a = _torch_reference(rp, col, val, xa.abs() * 2.0 + xb.abs() * 3.0, torch.zeros_like(y0), 1.0, 0.0)
        lin = (yc.double() - (2.0 * ya.double() - 3.0 * yb.double())).abs()
        assert float((lin / torch.where(s_a > 0, s_a, torch.ones_like(s_a))).max()) <= 4 * TOL
        del ya, yb, yc, s_a, lin

        # --- a second strategy agrees (merge-path tiles + carry fix-up; rows may be split differently) ---
        if which != "c5":
            eng.force_kernel(idx, capi.KERNEL_MERGE)
            _run(eng, idx, xa, y0, y2)
            assert _max_err(y2, y64, scale) <= TOL
            eng.force_kernel(idx, capi.KERNEL_AUTO)

        # --- integer artefacts ---
        blocked = info["kernel_name"] == "blocked"
        if info["num_slabs"] == 0:
            tr, tn = eng.plan_tiles(idx)
            # tiles / panels cover the nonzeros (blocked: the PIECES, runs of one row inside a 512-entry group) in order
            covered = eng.plan_blocked(idx, arrays=False)["num_pieces"] if blocked else d.nnz
            assert tn[0] == 0 and tn[-1] == covered and np.all(np.diff(tn) >= 0) and covered <= d.nnz
            assert tr[0] == 0 and tr[-1] == spec.rows and np.all(np.diff(tr) >= 0)
        split = eng.plan_split_rows(idx)
        assert np.all(np.diff(split) > 0)
        lens = (rp[1:] - rp[:-1])
        if which == "c2":
            assert int(lens.max()) == 1_000_000                                  # the clipped head of the power law
            if blocked:
                pb = eng.plan_blocked(idx, arrays=False)
                assert info["slab_runs"] * 5 <= d.nnz * 2 and info["slab_runs"] <= pb["num_pieces"] <= d.nnz // 2
                # the same matrix on the one-pass kernel: heavy rows are cut into LONG tiles and meet through carries
                eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
                info = eng.matrix_info(idx)
                split = eng.plan_split_rows(idx)
                _run(eng, idx, xa, y0, y2)
                assert _max_err(y2, y64, scale) <= TOL
            long_rows = torch.nonzero(lens >= info["long_threshold"]).flatten().cpu().numpy()   # rows cut into LONG tiles
            assert split.size > 0 and np.isin(split, long_rows).all()
        else:
            assert split.size == 0
    finally:
        eng.close()
        d.close()


def test_full_size_gemv_8192():
    """8192 x 8192 GeMV with the reference's closed-form inputs (cpu/src/main.cpp:213-226) against float64 torch."""
    import torch
    from hispmv_b200 import Engine
    n = 8192
    i = torch.arange(n, device="cuda", dtype=torch.float32)
    a = ((i + 1)[:, None] / (i + 2)[None, :]).contiguous()
    x = (i + 1) / (i + 2)
    y0 = -2.0 * (i + 1) / (i + 2)
    eng = Engine(0)
    try:
        idx = eng.create_dense_handle_dev(a, n, n)
        y = torch.empty(n, device="cuda")
        _run(eng, idx, x, y0, y)
        y64 = ALPHA * (a.double() @ x.double()) + BETA * y0.double()
        scale = abs(ALPHA) * (a.double().abs() @ x.double().abs()) + abs(BETA) * y0.double().abs()
        assert _max_err(y, y64, scale) <= TOL
        y2 = torch.empty_like(y)
        _run(eng, idx, x, y0, y2)
        assert torch.equal(y.view(torch.int32), y2.view(torch.int32))
    finally:
        eng.close()
