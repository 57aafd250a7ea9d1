"""The C-ABI library loads without a GPU and exports exactly what include/*.h declares; the plugin module imports;
the product has no CPU fallback (hispmv_create fails without an sm_100 device) and never touches oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("hispmv.h", "hispmv_synth.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hispmv_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(ROOT, "hispmv_b200", "libhispmv_cuda.so"))
    declared = _declared()
    assert len(declared) >= 30
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    from hispmv_b200 import capi
    assert sorted(capi.EXPORTED) == declared, sorted(set(capi.EXPORTED) ^ set(declared))


def test_plugin_module_imports_with_reference_surface():
    import pyhispmv
    for m in ("create_dense_handle", "create_sparse_handle", "load_matrices", "select_matrix", "run_kernel", "linear"):
        assert hasattr(pyhispmv.FpgaHandle, m)      # pyhispmv/src/pyhispmv_bindings.cpp:13-39


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from hispmv_b200 import capi
    ctx = C.c_void_p()
    assert capi.lib.hispmv_create(C.byref(ctx), 0, 3) == capi.ERR_CUDA
    import pyhispmv
    with pytest.raises(RuntimeError):
        pyhispmv.FpgaHandle("x.xclbin", 0, 24, 1, 1, 2, 5, True, False, True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hispmv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in src and "liboracle" not in src and "oracle/_ref" not in src, f
