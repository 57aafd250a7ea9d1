"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the unmodified reference
compiled into oracle/_ref) against (a) the CPU restatement in oracle/ -- runs everywhere, no GPU, no /root/reference --
and (b) the CUDA path through the C-ABI (-m gpu).

Bars: bit-exact where the reference fixes the arithmetic order (COO->CSR, cpu_spmv, cpuSequential, naive_gemv);
|y - y64| / (|alpha| sum_j |a_ij x_j| + |beta y0_i|) <= 1e-5 against MKL outputs (north_star tolerance)."""
import importlib.util
import os
import zlib

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

TOL = 1e-5


def _gold(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# ---- (a) the oracle ------------------------------------------------------------------------------------
def test_oracle_spmv_c1_against_golden():
    d, g = make_golden.case_inputs("spmv_c1"), _gold("spmv_c1")
    rp, ci, vv = ol.coo_to_csr(d["rows"], d["r"], d["c"], d["v"])
    assert np.array_equal(rp, g["row_ptr"])
    assert _crc(ci) == int(g["col_idx_crc"]) and _crc(vv) == int(g["vals_crc"])
    y = d["y0"].copy()
    ol.oracle().oracle_spmv_csr_f32(d["rows"], rp, ci, vv, d["x"], y, d["alpha"], d["beta"])
    assert np.array_equal(y.view(np.uint32), g["y_cpu_spmv"].view(np.uint32))        # cpu_spmv: bit-exact
    y64, scale = ol.spmv_f64(rp, ci, vv, d["x"], d["y0"], d["alpha"], d["beta"])
    err, at = ol.max_scaled_error(g["y_mkl"], y64, scale)                             # MKL: within tolerance
    assert err <= TOL, (err, at)


def test_oracle_spmv_host_against_golden():
    d, g = make_golden.case_inputs("spmv_host"), _gold("spmv_host")
    out = np.zeros(d["rows"], np.float32)
    ol.oracle().oracle_spmv_coo_f32(d["rows"], d["r"].size, d["r"], d["c"], d["v"], d["x"], d["y0"], d["alpha"],
                                    d["beta"], out)
    assert np.array_equal(out.view(np.uint32), g["cout_cpu_sequential"].view(np.uint32))  # cpuSequential: bit-exact


def test_oracle_gemv_against_golden():
    d, g = make_golden.case_inputs("gemv"), _gold("gemv")
    y = d["y0"].copy()
    ol.oracle().oracle_gemv_f32(d["rows"], d["cols"], d["a"].reshape(-1), d["x"], y, d["alpha"], d["beta"])
    assert np.array_equal(y.view(np.uint32), g["y_naive_gemv"].view(np.uint32))      # naive_gemv: bit-exact
    y64, scale = ol.gemv_f64(d["a"], d["rows"], d["cols"], d["x"], d["y0"], d["alpha"], d["beta"])
    err, at = ol.max_scaled_error(g["y_mkl"], y64, scale)
    assert err <= TOL, (err, at)


# ---- (b) the CUDA path ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from hispmv_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _near(y, y_ref, scale, k=2.0):
    """both sides are fp32 results within TOL of the float64 value, so they differ by at most 2*TOL*scale"""
    return np.all(np.abs(y.astype(np.float64) - y_ref.astype(np.float64)) <= k * TOL * np.maximum(scale, 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 6, 7])
def test_cuda_spmv_c1_against_golden(eng, kernel):
    d, g = make_golden.case_inputs("spmv_c1"), _gold("spmv_c1")
    idx = eng.create_sparse_handle(d["r"], d["c"], d["v"], d["rows"], d["cols"])
    if kernel:
        eng.force_kernel(idx, kernel)
    rp, ci, vv = eng.plan_csr(idx)
    assert np.array_equal(rp, g["row_ptr"]) and _crc(ci) == int(g["col_idx_crc"]) and _crc(vv) == int(g["vals_crc"])
    y = np.zeros(d["rows"], np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(d["x"], d["y0"], y, float(d["alpha"]), float(d["beta"]))
    y64, scale = ol.spmv_f64(rp, ci, vv, d["x"], d["y0"], d["alpha"], d["beta"])
    err, at = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL, (err, at)
    assert _near(y, g["y_mkl"], scale) and _near(y, g["y_cpu_spmv"], scale)


@pytest.mark.gpu
def test_cuda_spmv_host_against_golden(eng):
    d, g = make_golden.case_inputs("spmv_host"), _gold("spmv_host")
    idx = eng.create_sparse_handle(d["r"], d["c"], d["v"], d["rows"], d["cols"])
    rp, ci, vv = eng.plan_csr(idx)
    y = np.zeros(d["rows"], np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(d["x"], d["y0"], y, float(d["alpha"]), float(d["beta"]))
    y64, scale = ol.spmv_f64(rp, ci, vv, d["x"], d["y0"], d["alpha"], d["beta"])
    assert _near(y, g["cout_cpu_sequential"], scale)


@pytest.mark.gpu
def test_cuda_gemv_against_golden(eng):
    d, g = make_golden.case_inputs("gemv"), _gold("gemv")
    idx = eng.create_dense_handle(d["a"].reshape(-1), d["rows"], d["cols"])
    y = np.zeros(d["rows"], np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(d["x"], d["y0"], y, float(d["alpha"]), float(d["beta"]))
    y64, scale = ol.gemv_f64(d["a"], d["rows"], d["cols"], d["x"], d["y0"], d["alpha"], d["beta"])
    err, at = ol.max_scaled_error(y, y64, scale)
    assert err <= TOL, (err, at)
    assert _near(y, g["y_mkl"], scale) and _near(y, g["y_naive_gemv"], scale)
