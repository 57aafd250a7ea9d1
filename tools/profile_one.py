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
    ap.add_argument("--kernel", default="auto", choices=["auto", "adaptive", "rowstage", "merge", "vector", "scalar", "gemv", "blocked"])
    ap.add_argument("--tile", default="")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--cols", type=int, default=8192)
    ap.add_argument("--flush", action="store_true", help="write a 256 MB buffer between runs (cold L2)")
    args = ap.parse_args()
    if args.tile:
        os.environ["HISPMV_MERGE_TILE"] = args.tile
    eng = Engine(0)
    if args.kernel == "gemv":
        a = torch.rand(args.rows, args.cols, device="cuda")
        idx = eng.create_dense_handle_dev(a, args.rows, args.cols)
        rows, cols = args.rows, args.cols
    elif args.config in ("c1", "c3b", "c3c"):   # the small host-built configs (as tools/sweep.py builds them)
        import numpy as np
        if args.config == "c1":
            r, c, v, rows, cols = synth.c1_imbalanced_coo()
        else:
            rows, cols, dens = {"c3b": (8192, 8192, 0.1), "c3c": (1024, 8192, 0.25)}[args.config]
            g0 = torch.Generator().manual_seed(0)
            w = torch.randn(rows, cols, generator=g0) * (torch.rand(rows, cols, generator=g0) < dens)
            nzr, nzc = torch.nonzero(w, as_tuple=True)
            r, c, v = nzr.numpy().astype(np.int32), nzc.numpy().astype(np.int32), w[nzr, nzc].numpy()
        idx = eng.create_sparse_handle(r, c, v, rows, cols)
        k = {"auto": capi.KERNEL_AUTO, "adaptive": capi.KERNEL_ADAPTIVE, "rowstage": capi.KERNEL_ROWSTAGE, "merge": capi.KERNEL_MERGE,
             "vector": capi.KERNEL_CSR_VECTOR, "scalar": capi.KERNEL_CSR_SCALAR, "blocked": capi.KERNEL_BLOCKED}[args.kernel]
        eng.force_kernel(idx, k, args.lanes)
    else:
        spec = {"c2": synth.c2_powerlaw, "c4": synth.c4_stencil, "c5": synth.c5_uniform}[args.config](args.scale)
        d = synth.DeviceCSR(spec)
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        d.close()
        k = {"auto": capi.KERNEL_AUTO, "adaptive": capi.KERNEL_ADAPTIVE, "rowstage": capi.KERNEL_ROWSTAGE, "merge": capi.KERNEL_MERGE, "vector": capi.KERNEL_CSR_VECTOR,
             "scalar": capi.KERNEL_CSR_SCALAR, "blocked": capi.KERNEL_BLOCKED}[args.kernel]
        if k != capi.KERNEL_AUTO:
            eng.force_kernel(idx, k, args.lanes)
        rows, cols = spec.rows, spec.cols
    x = torch.rand(cols, device="cuda") + 0.5
    b = torch.rand(rows, device="cuda")
    y = torch.empty(rows, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(256 * 1024 * 1024 // 4, device="cuda") if args.flush else None
    flush_r = torch.zeros(256 * 1024 * 1024 // 4, device="cuda") if args.flush else None
    for _ in range(args.runs):
        if flush is not None:
            flush.add_(1.0)          # cold L2 for the next run ...
            _ = flush_r.sum()        # ... and clean (the flush's dirty lines are written back before the kernel)
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
    torch.cuda.synchronize()
    print(eng.matrix_info(idx)["kernel_name"], float(y.sum()))
    eng.close()


if __name__ == "__main__":
    main()
