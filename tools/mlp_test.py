#!/usr/bin/env python
"""The reference's DNN-layer demo on the GPU engine (apps/model_test.py:38-90): build the three-layer model, hand its
layers to the accelerator through the layer manager, run one batch through both models and print the timing and the
error report.  Same command-line flags as the reference script; the accelerator handle is `pyhispmv.FpgaHandle`.

    python tools/mlp_test.py [--batch_size 1] [--input_size 4096] [--hidden_size_1 8192] [--hidden_size_2 8192]
                             [--output_size 1024] [--density1 0.1] [--density2 0.25]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLAGS = (("batch_size", int, 1, "Batch size"), ("input_size", int, 4096, "Input size"),
         ("hidden_size_1", int, 8192, "Size of the first hidden layer"),
         ("hidden_size_2", int, 8192, "Size of the second hidden layer"), ("output_size", int, 1024, "Output size"),
         ("density1", float, 0.1, "Density for the first sparse layer"),
         ("density2", float, 0.25, "Density for the second sparse layer"))


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    for name, typ, default, text in FLAGS:
        ap.add_argument("--" + name, type=typ, default=default, help=text)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()

    import torch
    import pyhispmv
    from hispmv_b200.layers import (FpgaLayerManager, ThreeLayerFCModel, ThreeLayerFCModelConfig, compare_model_outputs)

    # the reference's build parameters (apps/model_test.py:19-35); only dense_overlay and row_dist_net mean anything here
    fpga = pyhispmv.FpgaHandle("unused.xclbin", a.device, 24, 1, 1, 2, 5, True, False, True)
    cpu_model = ThreeLayerFCModel(ThreeLayerFCModelConfig(a.input_size, a.hidden_size_1, a.hidden_size_2, a.output_size,
                                                          a.density1, a.density2)).eval()
    for p in cpu_model.parameters():
        p.requires_grad = False
    x = torch.randn((a.batch_size, a.input_size))
    gpu_model = FpgaLayerManager().replace_layers(cpu_model, fpga).eval()
    with torch.no_grad():
        gpu_model(x)                                   # first call: staging buffers
        t0 = time.time()
        gpu_out = gpu_model(x)
        t_gpu = time.time() - t0
        t0 = time.time()
        cpu_out = cpu_model(x)
        t_cpu = time.time() - t0
    print("\n")
    print(f"FPGA Inference Time: {t_gpu:.6f} seconds")           # the reference's line names, for its log readers
    print(f"CPU Inference Time (single thread): {t_cpu:.6f} seconds")
    compare_model_outputs(cpu_out, gpu_out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
