"""model_test's MLP (apps/model_test.py:42-48: 4096 -> 8192 dense -> 8192 (d=0.1) -> 1024 (d=0.25), batch 1) on one GPU:
the plugin path the reference apps use (FpgaLinear -> FpgaHandle.linear, host buffers every layer) beside the
device-resident chain (bias + ReLU fused, activations stay in HBM), with and without CUDA-graph replay."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hispmv_b200 import Engine  # noqa: E402
from hispmv_b200.layers import DeviceChain, ThreeLayerFCModel, ThreeLayerFCModelConfig  # noqa: E402


def main():
    torch.manual_seed(0)
    model = ThreeLayerFCModel(ThreeLayerFCModelConfig()).eval()
    for p in model.parameters():
        p.requires_grad = False
    x = torch.randn(4096)
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(5):
            ref = model(x.view(1, -1))
        cpu_us = (time.perf_counter() - t0) / 5 * 1e6
    ref = ref.numpy().reshape(-1)
    eng = Engine(0)
    layers = [model.dense, model.sparse1, model.sparse2]
    out = {}
    for name, graph in (("device chain", False), ("device chain, CUDA graph", True)):
        ch = DeviceChain(eng, layers, relu=[True, True, True], graph=graph)
        xd = x.cuda()
        y = ch.forward(xd).cpu().numpy()
        assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max(), name
        for _ in range(10):
            ch.forward(xd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 500
        e0.record()
        for _ in range(iters):
            ch.forward(xd)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters * 1e3
    ch = DeviceChain(eng, layers, relu=[True, True, True])
    for batch in (8, 64):
        xb = torch.randn(batch, 4096, device="cuda")
        for _ in range(5):
            ch.forward_batch(xb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            ch.forward_batch(xb)
        e1.record()
        torch.cuda.synchronize()
        out[f"device chain, batch {batch} (us per vector)"] = e0.elapsed_time(e1) / 100 * 1e3 / batch
    # the plugin path: host numpy in, host numpy out, per layer (apps/fpga_layer_manager.py:58-67)
    import pyhispmv
    from hispmv_b200.layers import FpgaLayerManager
    import copy
    fpga = pyhispmv.FpgaHandle("unused.xclbin", 0, 24, 1, 1, 2, 5, True, False, True)
    fmodel = FpgaLayerManager().replace_layers(copy.deepcopy(model), fpga)    # apps/model_test.py:53-60
    if fmodel is not None:
        with torch.no_grad():
            yp = fmodel(x.view(1, -1)).numpy().reshape(-1)
            assert np.abs(yp - ref).max() <= 1e-4 * np.abs(ref).max()
            for _ in range(20):
                fmodel(x.view(1, -1))
            t0 = time.perf_counter()
            for _ in range(200):
                fmodel(x.view(1, -1))
            out["plugin path (host buffers per layer)"] = (time.perf_counter() - t0) / 200 * 1e6
            # where it goes: the three fpga.linear calls on their own (numpy in, numpy out), the rest is torch glue
            from hispmv_b200.layers import FpgaLinear
            total = 0.0
            for name, mod in fmodel.named_modules():
                if isinstance(mod, FpgaLinear):
                    cols = {"dense": 4096, "sparse1": 8192, "sparse2": 8192}.get(name.split(".")[-1], 4096)
                    xin = np.random.default_rng(1).standard_normal(cols).astype(np.float32)
                    for _ in range(20):
                        fpga.linear(mod.matrix_idx, xin, mod.bias_npy)
                    t0 = time.perf_counter()
                    for _ in range(300):
                        fpga.linear(mod.matrix_idx, xin, mod.bias_npy)
                    us = (time.perf_counter() - t0) / 300 * 1e6
                    total += us
                    out[f"  fpga.linear alone: {name}"] = us
            out["  three fpga.linear calls"] = total
    flops = 2 * (8192 * 4096 + int(model.sparse1.weight._nnz()) + int(model.sparse2.weight._nnz()))
    print(f"chain_one_gpu: CPU model (torch, {torch.get_num_threads()} threads) {cpu_us:.0f} us per pass")
    for name, us in out.items():
        print(f"  {name:40s} {us:8.1f} us per pass  ({flops / us / 1e3:.1f} GFLOP/s)")
    eng.close()


if __name__ == "__main__":
    main()
