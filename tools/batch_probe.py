"""Batch > 1 on the model_test sparse layers: eight vectors in one pass (hispmv_run_dev_batch) against eight launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hispmv_b200 import Engine  # noqa: E402


def timed(fn, iters=50):
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def main():
    eng = Engine(0)
    st = torch.cuda.current_stream().cuda_stream
    for name, nr, nc, dens in (("C3b 8192x8192 d=0.1", 8192, 8192, 0.1), ("C3c 1024x8192 d=0.25", 1024, 8192, 0.25)):
        g0 = torch.Generator().manual_seed(0)
        w = torch.randn(nr, nc, generator=g0) * (torch.rand(nr, nc, generator=g0) < dens)
        nzr, nzc = torch.nonzero(w, as_tuple=True)
        idx = eng.create_sparse_handle(nzr.numpy().astype(np.int32), nzc.numpy().astype(np.int32), w[nzr, nzc].numpy(), nr, nc)
        b = torch.rand(nr, device="cuda")
        for nv in (2, 4, 8, 16):
            X = torch.rand(nv, nc, device="cuda")
            Y = torch.empty(nv, nr, device="cuda")
            Y1 = torch.empty(nv, nr, device="cuda")
            t_batch = timed(lambda: eng.run_dev_batch(idx, X, b, Y, 1.0, 1.0, stream=st))

            def one_by_one():
                for k in range(nv):
                    eng.run_dev(idx, X[k], b, Y1[k], 1.0, 1.0, st)
            t_seq = timed(one_by_one)
            torch.cuda.synchronize()
            rel = float(((Y - Y1).abs() / (Y1.abs() + 1.0)).max())
            print(f"{name}: {nv:2d} vectors  one pass {t_batch:7.1f} us   vector by vector {t_seq:7.1f} us   "
                  f"({t_seq / t_batch:.1f}x)  max diff {rel:.1e}", flush=True)
    for name, nr, nc in (("dense 8192x4096 (MLP layer 1)", 8192, 4096), ("dense 8192x8192", 8192, 8192)):
        a = torch.rand(nr, nc, device="cuda") - 0.5
        idx = eng.create_dense_handle_dev(a, nr, nc)
        b = torch.rand(nr, device="cuda")
        for nv in (2, 8, 16):
            X = torch.rand(nv, nc, device="cuda")
            Y = torch.empty(nv, nr, device="cuda")
            Y1 = torch.empty(nv, nr, device="cuda")
            t_batch = timed(lambda: eng.run_dev_batch(idx, X, b, Y, 1.0, 1.0, stream=st))

            def one_by_one():
                for k in range(nv):
                    eng.run_dev(idx, X[k], b, Y1[k], 1.0, 1.0, st)
            t_seq = timed(one_by_one)
            torch.cuda.synchronize()
            rel = float(((Y - Y1).abs() / (Y1.abs() + 1.0)).max())
            print(f"{name}: {nv:2d} vectors  one pass {t_batch:7.1f} us   vector by vector {t_seq:7.1f} us   "
                  f"({t_seq / t_batch:.1f}x)  max diff {rel:.1e}", flush=True)
        del a
    eng.close()


if __name__ == "__main__":
    main()
