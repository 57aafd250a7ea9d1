"""tools/spmv_host.py's error report against the reference's HiSpmvHandle::printErrorStats
(common/src/spmv-helper.cpp:835-895, compiled unmodified into oracle/_ref): the same text, byte for byte."""
import contextlib
import importlib.util
import io
import os

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("spmv_host", os.path.join(os.path.dirname(HERE), "tools", "spmv_host.py"))
spmv_host = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(spmv_host)

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")


def _reference_text(cpu, out):
    import ctypes as C
    b = C.create_string_buffer(8192)
    ol.ref_common().ref_common_error_stats(cpu.size, np.ascontiguousarray(cpu, np.float32),
                                           np.ascontiguousarray(out, np.float32), b, 8192)
    return b.value.decode()


def _ours(cpu, out):
    s = io.StringIO()
    with contextlib.redirect_stdout(s):
        spmv_host.print_error_stats(cpu, out)
    return s.getvalue()


@pytest.mark.parametrize("case", ["identical", "few", "many", "signs", "spread"])
def test_same_text_as_the_reference(case):
    rng = np.random.default_rng({"identical": 0, "few": 1, "many": 2, "signs": 3, "spread": 4}[case])
    n = 5000
    cpu = (rng.standard_normal(n) * 100).astype(np.float32)
    out = cpu.copy()
    if case == "few":
        k = rng.choice(n, 7, replace=False)
        out[k] *= np.float32(1.0 + 1e-5)
    elif case == "many":
        out = (cpu * (1 + rng.standard_normal(n).astype(np.float32) * np.float32(1e-6))).astype(np.float32)
    elif case == "signs":                                  # the reference compares magnitudes: a flipped sign is a match
        out = -cpu
        out[:40] *= np.float32(1.001)
    elif case == "spread":
        out = (cpu * (1 + np.float32(10.0) ** rng.integers(-7, -1, n).astype(np.float32))).astype(np.float32)
    assert _ours(cpu, out) == _reference_text(cpu, out)
