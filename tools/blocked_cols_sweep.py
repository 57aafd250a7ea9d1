"""Blocked against one-pass as x shrinks (10 M rows, ~100 M nnz, power-law rows, gamma 5 / 1): where the selector's
column threshold should sit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hispmv_b200 import Engine, capi, synth  # noqa: E402
from blocked_crossover import timed  # noqa: E402


def main():
    k_c2 = int(round(0.6912 * 2 ** 32))
    flush = (torch.zeros(64 << 20, device="cuda"), torch.zeros(64 << 20, device="cuda"))
    for cols in (2_000_000, 1_000_000, 500_000, 250_000, 100_000):
        for gamma in (5, 1):
            spec = synth.SynthSpec("C2", 1, 1, 10_000_000, cols, (k_c2, min(1_000_000, cols), gamma))
            eng = Engine(0)
            d = synth.DeviceCSR(spec)
            idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
            nnz = d.nnz
            d.close()
            auto = eng.matrix_info(idx)["kernel_name"]
            x = torch.rand(spec.cols, device="cuda")
            b = torch.rand(spec.rows, device="cuda")
            y = torch.empty(spec.rows, device="cuda")
            out = {}
            for kname, k in (("blocked", capi.KERNEL_BLOCKED), ("one-pass", capi.KERNEL_ADAPTIVE)):
                eng.force_kernel(idx, k)
                out[kname] = timed(eng, idx, x, b, y, flush)
                if kname == "blocked":
                    pieces = eng.plan_blocked(idx, arrays=False)["num_pieces"]
            print(f"cols={cols:8d} gamma={gamma} nnz={nnz} auto={auto:8s} pieces/nnz={pieces / nnz:.3f} blocked {out['blocked']:.4f} ms  "
                  f"one-pass {out['one-pass']:.4f} ms  ratio {out['one-pass'] / out['blocked']:.2f}", flush=True)
            eng.close()


if __name__ == "__main__":
    main()
