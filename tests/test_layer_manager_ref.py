"""hispmv_b200.layers.LayerManager / FpgaLinear against the reference's own apps/fpga_layer_manager.py, imported from
/root/reference in the build container (it cannot travel: on the GPU box these tests skip).  Both managers are driven
with the same models and a recording stand-in for the accelerator handle; the calls they make on it -- which handle
type, which arrays, in which order -- and what the replaced model computes must be identical."""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from hispmv_b200 import layers as L

REF = "/root/reference/apps/fpga_layer_manager.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="/root/reference is not present on this machine")


def _reference_manager():
    stub = types.ModuleType("model")            # apps/model.py needs sparse_dot_mkl (absent); only SparseLinear is used
    stub.SparseLinear = L.SparseLinear
    saved = sys.modules.get("model")
    sys.modules["model"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_fpga_layer_manager", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("model", None)
        else:
            sys.modules["model"] = saved
    return mod.FpgaLayerManager()


class RecordingHandle:
    """Stands in for pyhispmv.FpgaHandle: records every call, computes linear() with numpy."""

    def __init__(self, full_after=None):
        self.calls, self.mats, self.full_after = [], [], full_after

    def _add(self, kind, dense):
        if self.full_after is not None and len(self.mats) >= self.full_after:
            return -1
        self.mats.append(dense)
        return len(self.mats) - 1

    def create_dense_handle(self, flat, rows, cols):
        self.calls.append(("dense", np.array(flat, np.float32).copy(), int(rows), int(cols)))
        return self._add("dense", np.array(flat, np.float32).reshape(rows, cols))

    def create_sparse_handle(self, r, c, v, rows, cols):
        self.calls.append(("sparse", np.array(r).copy(), np.array(c).copy(), np.array(v, np.float32).copy(), int(rows), int(cols)))
        d = np.zeros((rows, cols), np.float32)
        np.add.at(d, (np.array(r), np.array(c)), np.array(v, np.float32))
        return self._add("sparse", d)

    def load_matrices(self):
        self.calls.append(("load",))

    def linear(self, idx, x, bias):
        a = self.mats[idx]
        xs = np.asarray(x, np.float32).reshape(-1, a.shape[1])
        self.calls.append(("linear", idx, xs.shape[0]))
        return (xs @ a.T + np.asarray(bias, np.float32)).reshape(-1).astype(np.float32)


def _same_calls(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for x, y in zip(a, b):
        assert x[0] == y[0]
        for u, v in zip(x[1:], y[1:]):
            if isinstance(u, np.ndarray):
                assert u.dtype == v.dtype or (u.dtype.kind == v.dtype.kind), (x[0], u.dtype, v.dtype)
                assert np.array_equal(u, v), x[0]
            else:
                assert u == v, x[0]


def _mlp():
    torch.manual_seed(3)
    m = L.ThreeLayerFCModel(L.ThreeLayerFCModelConfig(48, 96, 64, 24, 0.1, 0.25)).eval()
    for p in m.parameters():
        p.requires_grad = False
    return m


def test_process_weights_makes_the_same_calls():
    ref_mgr, ours = _reference_manager(), L.LayerManager()
    torch.manual_seed(0)
    dense = nn.Linear(40, 30)
    sparse_ish = nn.Linear(50, 20)
    with torch.no_grad():
        sparse_ish.weight.mul_((torch.rand(20, 50) < 0.3).float())      # density 0.3 -> COO handle
    half = nn.Linear(8, 4)
    with torch.no_grad():
        half.weight.copy_(torch.tensor([[1.0, 0] * 4] * 4))             # density exactly 0.5 -> still COO (> 0.5 is dense)
    for layer in (dense, sparse_ish, half, L.SparseLinear(60, 30, 0.2)):
        a, b = RecordingHandle(), RecordingHandle()
        with contextlib.redirect_stdout(io.StringIO()):
            ia, ba = ref_mgr.process_weights(layer, a)
        ib, bb = ours.process_weights(layer, b)
        assert ia == ib and np.array_equal(np.asarray(ba), np.asarray(bb))
        _same_calls(a.calls, b.calls)


def test_memory_full_raises_the_same_error():
    ref_mgr, ours = _reference_manager(), L.LayerManager()
    layer = nn.Linear(6, 5)
    for mgr in (ref_mgr, ours):
        with pytest.raises(RuntimeError, match="FPGA memory is full"):
            mgr.process_weights(layer, RecordingHandle(full_after=0))


def test_replaced_model_makes_the_same_calls_and_results():
    ref_mgr, ours = _reference_manager(), L.LayerManager()
    model = _mlp()
    a, b = RecordingHandle(), RecordingHandle()
    with contextlib.redirect_stdout(io.StringIO()):
        ref_model = ref_mgr.replace_layers(model, a)
    our_model = ours.replace_layers(model, b)
    if ref_model is None:        # the reference's replace_layers builds the model in place and returns it at the end
        pytest.skip("reference replace_layers returned nothing")
    for batch in (1, 3):
        x = torch.randn(batch, 48)
        with torch.no_grad():
            ya = ref_model(x)
            yb = our_model(x)
            y0 = model(x)
        assert ya.shape == yb.shape == y0.shape
        assert torch.equal(ya, yb)
        assert torch.allclose(yb, y0, rtol=1e-4, atol=1e-4)
    _same_calls(a.calls, b.calls)


def _reference_model_module():
    """apps/model.py with sparse_dot_mkl (absent here) replaced by scipy's CSR product -- same arithmetic, MKL aside."""
    stub = types.ModuleType("sparse_dot_mkl")
    stub.dot_product_mkl = lambda a, b: np.asarray(a @ b, dtype=np.float32)
    saved = sys.modules.get("sparse_dot_mkl")
    sys.modules["sparse_dot_mkl"] = stub
    try:
        spec = importlib.util.spec_from_file_location("ref_apps_model", "/root/reference/apps/model.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("sparse_dot_mkl", None)
        else:
            sys.modules["sparse_dot_mkl"] = saved
    return mod


def test_model_classes_draw_the_same_weights_and_compute_the_same_outputs():
    """layers.SparseLinear / ThreeLayerFCModel against apps/model.py:10-80: under the same torch seed the layers hold
    the same weights (same order of random draws, same mask rule) and the forward pass gives the same activations."""
    ref = _reference_model_module()
    sizes = (40, 72, 56, 20, 0.1, 0.25)
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(11)
        a = ref.ThreeLayerFCModel(ref.ThreeLayerFCModelConfig(*sizes), rp_time=1).eval()
    torch.manual_seed(11)
    b = L.ThreeLayerFCModel(L.ThreeLayerFCModelConfig(*sizes)).eval()
    assert torch.equal(a.dense.weight, b.dense.weight) and torch.equal(a.dense.bias, b.dense.bias)
    for name in ("sparse1", "sparse2"):
        wa, wb = getattr(a, name).weight.coalesce(), getattr(b, name).weight.coalesce()
        assert wa.shape == wb.shape and torch.equal(wa.indices(), wb.indices()) and torch.equal(wa.values(), wb.values())
        assert torch.equal(getattr(a, name).bias, getattr(b, name).bias)
    x = torch.randn(3, 40)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ya = a(x)
        yb = b(x)
    assert ya.shape == yb.shape == (3, 20)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-5)


def test_compare_model_outputs_prints_the_same_report():
    ref = _reference_model_module()
    torch.manual_seed(5)
    a = torch.randn(4, 300) * 50
    b = a * (1 + 1e-4 * torch.randn(4, 300))
    b[0, 7] += 0.5
    texts = []
    for fn in (ref.compare_model_outputs, L.compare_model_outputs):
        s = io.StringIO()
        with contextlib.redirect_stdout(s):
            fn(a, b)
        texts.append(s.getvalue())
    assert texts[0] == texts[1] and "Max Relative Error" in texts[1]
