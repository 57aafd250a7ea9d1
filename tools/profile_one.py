"""Run one kernel strategy on one synthetic config a few times -- the command ncu wraps.

    python tools/profile_one.py --config c2 --kernel merge [--tile 1792] [--lanes 8] [--scale 1.0] [--runs 3]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hispmv_b200 import Engine, capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--kernel", default="auto", choices=["auto", "adaptive", "rowstage", "merge", "vector", "scalar", "gemv"])
    ap.add_argument("--tile", default="")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--cols", type=int, default=8192)
    args = ap.parse_args()
    if args.tile:
        os.environ["HISPMV_MERGE_TILE"] = args.tile
    eng = Engine(0)
    if args.kernel == "gemv":
        a = torch.rand(args.rows, args.cols, device="cuda")
        idx = eng.create_dense_handle_dev(a, args.rows, args.cols)
        rows, cols = args.rows, args.cols
    else:
        spec = {"c2": synth.c2_powerlaw, "c4": synth.c4_stencil, "c5": synth.c5_uniform}[args.config](args.scale)
        d = synth.DeviceCSR(spec)
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        d.close()
        k = {"auto": capi.KERNEL_AUTO, "adaptive": capi.KERNEL_ADAPTIVE, "rowstage": capi.KERNEL_ROWSTAGE, "merge": capi.KERNEL_MERGE, "vector": capi.KERNEL_CSR_VECTOR,
             "scalar": capi.KERNEL_CSR_SCALAR}[args.kernel]
        eng.force_kernel(idx, k, args.lanes)
        rows, cols = spec.rows, spec.cols
    x = torch.rand(cols, device="cuda") + 0.5
    b = torch.rand(rows, device="cuda")
    y = torch.empty(rows, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(args.runs):
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
    torch.cuda.synchronize()
    print(eng.matrix_info(idx)["kernel_name"], float(y.sum()))
    eng.close()


if __name__ == "__main__":
    main()
