"""Robustness sweep of the blocked strategy at large sizes: C2-shaped matrices with other seeds, sizes, column skews and
row-length tails than the benchmark's.  For each: the blocked result against the float64 oracle on the first 200 000
rows, and against the one-pass kernel on EVERY row (two independent kernels over independent plans)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from hispmv_b200 import Engine, capi, synth  # noqa: E402


def main():
    ol.build()
    k_c2 = int(round(0.6912 * 2 ** 32))
    cases = [
        ("seed7", synth.SynthSpec("C2", 1, 7, 10_000_000, 10_000_000, (k_c2, 1_000_000, 5))),
        ("7Mx9M_seed3", synth.SynthSpec("C2", 1, 3, 7_000_003, 9_000_001, (k_c2, 1_000_000, 5))),
        ("gamma2_seed5", synth.SynthSpec("C2", 1, 5, 8_000_000, 12_345_677, (k_c2, 1_000_000, 2))),
        ("heavy_tail", synth.SynthSpec("C2", 1, 11, 6_000_000, 6_000_000, (2 * k_c2, 3_000_000, 6))),
        ("short_rows", synth.SynthSpec("C2", 1, 13, 30_000_000, 5_000_000, (k_c2 // 4, 20_000, 4))),
    ]
    worst = 0.0
    for name, spec in cases:
        eng = Engine(0)
        d = synth.DeviceCSR(spec)
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        nnz = d.nnz
        d.close()
        auto = eng.matrix_info(idx)["kernel_name"]
        x = torch.rand(spec.cols, device="cuda") - 0.3
        b = torch.rand(spec.rows, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ys = {}
        for kname, k in (("blocked", capi.KERNEL_BLOCKED), ("adaptive", capi.KERNEL_ADAPTIVE)):
            eng.force_kernel(idx, k)
            y = torch.full((spec.rows,), float("nan"), device="cuda")
            eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
            torch.cuda.synchronize()
            ys[kname] = y
        n = 200_000
        rp, ci, vv = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, n)
        y64, scale = ol.spmv_f64(rp, ci, vv, x.cpu().numpy(), b[:n].cpu().numpy(), 0.85, -2.06)
        err, _ = ol.max_scaled_error(ys["blocked"][:n].cpu().numpy(), y64, scale)
        # every row: the two kernels against each other, scaled by what the rows add up (|y| + |beta b| as a floor)
        diff = (ys["blocked"] - ys["adaptive"]).abs()
        ref = ys["adaptive"].abs() + 2.06 * b.abs() + 1e-6
        rel = float((diff / ref).max())
        nan = int(torch.isnan(ys["blocked"]).sum())
        worst = max(worst, err)
        print(f"{name:14s} rows={spec.rows} cols={spec.cols} nnz={nnz} auto={auto:8s} blocked vs oracle (200k rows) {err:.2e}  "
              f"blocked vs one-pass (all rows, relative to |y|+|beta b|) {rel:.2e}  NaN rows {nan}", flush=True)
        eng.close()
        assert nan == 0 and err <= 1e-5, name
    print("worst", worst)


if __name__ == "__main__":
    main()
