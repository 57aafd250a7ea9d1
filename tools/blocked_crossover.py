"""Where the blocked strategy stops paying: C2-shaped matrices (10 M x 10 M, ~100 M nnz) with column skew gamma = 5 .. 1
(1 = uniform).  Prints (row, slab) runs per nonzero -- the selector's input -- and both kernels' times (cold L2)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hispmv_b200 import Engine, capi, synth  # noqa: E402


def timed(eng, idx, x, b, y, flush, iters=12):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
    ts = []
    for _ in range(iters):
        flush[0].add_(1.0)
        _ = flush[1].sum()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    k_c2 = int(round(0.6912 * 2 ** 32))
    flush = (torch.zeros(64 << 20, device="cuda"), torch.zeros(64 << 20, device="cuda"))
    os.environ["HISPMV_BLOCKED_AUTO"] = "1"
    for gamma in (5, 4, 3, 2, 1):
        spec = synth.SynthSpec("C2", 1, 1, 10_000_000, 10_000_000, (k_c2, 1_000_000, gamma))
        eng = Engine(0)
        d = synth.DeviceCSR(spec)
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        nnz = d.nnz
        d.close()
        info = eng.matrix_info(idx)
        x = torch.rand(spec.cols, device="cuda")
        b = torch.rand(spec.rows, device="cuda")
        y = torch.empty(spec.rows, device="cuda")
        out = {}
        for kname, k in (("blocked", capi.KERNEL_BLOCKED), ("one-pass", capi.KERNEL_ADAPTIVE)):
            eng.force_kernel(idx, k)
            out[kname] = timed(eng, idx, x, b, y, flush)
            if kname == "blocked":
                pieces = eng.plan_blocked(idx, arrays=False)["num_pieces"]
        print(f"gamma={gamma} nnz={nnz} auto={info['kernel_name']:8s} runs/nnz={info.get('slab_runs', 0) / nnz:.3f} pieces/nnz={pieces / nnz:.3f} "
              f"blocked {out['blocked']:.4f} ms  one-pass {out['one-pass']:.4f} ms  ratio {out['one-pass'] / out['blocked']:.2f}", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
