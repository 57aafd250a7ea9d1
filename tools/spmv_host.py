#!/usr/bin/env python
"""spmv-host for the GPU engine: the reference's stand-alone self-checking host program (common/src/spmv-host.cpp:41-191)
with the same protocol and the same log lines, so that builds/collect_data.py:8-23 scrapes its output unchanged.

    python tools/spmv_host.py <matrix.mtx>            [--exec_ms 1000] [--power_s 0] [--device 0]
    python tools/spmv_host.py <dense rows> <dense cols> [--exec_ms 1000]

Protocol (spmv-host.cpp:17-23,43-44,92-100,181-189): x = c_in = (i+2)/(i+1), alpha = 0.55, beta = -2.05; a CPU result
(scipy CSR product in fp32 here; the reference uses its own cpuSequential) is compared with the accelerator's through
the relative-error histogram of printErrorStats (common/src/spmv-helper.cpp:835-895); GFLOPS = 2 (nnz + rows) / time.
The accelerator time is the mean of as many back-to-back device-resident runs as fit in --exec_ms ("rp_time").
--power_s S keeps the kernel running for S seconds while the board power is sampled (spmv-host.cpp:138-147,
common/src/spmv-helper.cpp:1027-1049) and prints the reference's Average Power / Max Power / Number of Samples lines,
plus one watt value per line in ./power_logs/<matrix>.log as the V100 benchmark does (gpu/src/nvmlPower.cpp:51-91).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def generate_vector(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float32)
    return ((i + np.float32(2.0)) / (i + np.float32(1.0))).astype(np.float32)   # spmv-host.cpp:17-23


def print_error_stats(cpu_ref: np.ndarray, out: np.ndarray) -> None:
    """HiSpmvHandle::printErrorStats (common/src/spmv-helper.cpp:835-895), line for line: relative error of the
    magnitudes against the CPU result (no floor under the reference value: its `epsilon` is numeric_limits::lowest(),
    i.e. negative), exact matches dropped, at most ten mismatches listed, otherwise ten equal-width bins between the
    smallest and the largest error with the last bin closed."""
    a, b = np.abs(out.astype(np.float32)).astype(np.float64), np.abs(cpu_ref.astype(np.float32)).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(a - b) / b
    rel = rel[rel != 0]                                   # NaN != 0 holds, as in the C++ test
    if rel.size == 0:
        print("No mismatch found")
        return
    if rel.size <= 10:
        print("Found atmost 10 mismatches, Relative Errors:")
        for e in rel:
            print(f"\t{e:g}")
        return
    lo, hi = rel.min(), rel.max()
    width = (hi - lo) / 10
    idx = np.minimum(((rel - lo) / width).astype(np.int64), 9)
    counts = np.bincount(idx, minlength=10)
    print("Relative Error Range:\tCount")
    for k in range(10):
        start = lo + k * width
        print(f"[{start:.3e}, {start + width:.3e}):\t{counts[k]}")


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("args", nargs="+", help="<matrix.mtx>  |  <dense rows> <dense cols>")
    ap.add_argument("--exec_ms", type=float, default=1000.0, help="time budget for the repeated runs (reference: --exec_ms)")
    ap.add_argument("--power_s", type=float, default=0.0, help="seconds of sampled execution (reference: --power_s)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--dump_y", default="", help="also save the accelerator's result vector (.npy), for the tests")
    a = ap.parse_args()
    import scipy.sparse as sp
    import torch
    from hispmv_b200 import Engine

    alpha, beta = 0.55, -2.05                                   # spmv-host.cpp:43-44
    eng = Engine(a.device)
    print("\nPreparing A Mtx...")
    t0 = time.perf_counter()
    if len(a.args) == 1:
        idx = eng.load_mtx(a.args[0])
        dense = None
    elif len(a.args) == 2:
        rows, cols = int(a.args[0]), int(a.args[1])
        dense = generate_vector(rows * cols).reshape(rows, cols)  # spmv-host.cpp:71-79
        idx = eng.create_dense_handle(dense.reshape(-1), rows, cols)
    else:
        print(f"Sparse Mode Usage: {sys.argv[0]} <sparse mtx>\nDense Mode Usage: {sys.argv[0]} <rows> <cols>", file=sys.stderr)
        return 1
    eng.load_matrices()
    print(f"Pre-processing Time: {time.perf_counter() - t0:.6f} secs")
    info = eng.matrix_info(idx)
    rows, cols, nnz = info["rows"], info["cols"], info["nnz"]
    print(f"Matrix A Length: {nnz}")
    print(f"Kernel: {info['kernel_name']} lanes={info['vector_lanes']} tiles={info['num_tiles']} "
          f"split rows={info['num_split_rows']} column slabs={info.get('num_slabs', 0)}")
    x, c_in = generate_vector(cols), generate_vector(rows)

    print("\nComputing on CPU... ")
    if dense is None:
        # the CPU side reads the file with an independent reader (scipy), not the engine's own CSR, so that an ingest
        # error shows up in the error report instead of cancelling out
        import scipy.io
        m = sp.csr_matrix(scipy.io.mmread(a.args[0])).astype(np.float32)
        if m.shape != (rows, cols):
            print(f"Error: the engine holds {rows}x{cols}, the file says {m.shape[0]}x{m.shape[1]}", file=sys.stderr)
            return 1
    t0 = time.perf_counter()
    ax = (m @ x) if dense is None else (dense @ x)
    cpu = (np.float32(alpha) * ax.astype(np.float32) + np.float32(beta) * c_in).astype(np.float32)
    t_cpu = time.perf_counter() - t0
    print(f"CPU TIME: {t_cpu * 1e3:.6f} ms")
    print(f"CPU GFLOPS: {2.0 * (nnz + rows) / (t_cpu * 1e9):.6f}")

    print("\nComputing on GPU... ")
    xd, cd = torch.from_numpy(x).cuda(a.device), torch.from_numpy(c_in).cuda(a.device)
    yd = torch.empty(rows, device=xd.device)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.run_dev(idx, xd, cd, yd, alpha, beta, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_dev(idx, xd, cd, yd, alpha, beta, st)
    e1.record()
    e1.synchronize()
    one_ms = max(e0.elapsed_time(e1), 1e-3)
    clock_mhz = getattr(torch.cuda.get_device_properties(a.device), "clock_rate", 1965000) / 1e3
    print(f"Approx. Clock Cycles: {int(one_ms * 1e3 * clock_mhz)}")   # SM cycles of one product (reference: FPGA cycles)
    # spmv-host.cpp:120-150: rp_time fills exec_ms and is capped at 2^15 (a uint16 kernel argument there); what does
    # not fit becomes whole repeats of the batch ("Num samples"), scaled up to power_s seconds for a power run
    rp_wanted = max(1.0, a.exec_ms / one_ms)
    rp_time = int(min(rp_wanted, 1 << 15))
    num_samples = rp_wanted / rp_time if int(rp_wanted) != rp_time else 1.0
    if a.power_s > 0:
        num_samples *= a.power_s * 1000.0 / a.exec_ms
    num_samples = max(1, int(num_samples))
    print(f"Using Repeat Time: {rp_time}")
    print(f"Using Num samples: {num_samples}")
    monitor = None
    if a.power_s > 0:
        from hispmv_b200.power import GpuPowerMonitor, report
        name = os.path.splitext(os.path.basename(a.args[0]))[0] if len(a.args) == 1 else f"dense_{a.args[0]}x{a.args[1]}"
        monitor = GpuPowerMonitor(period_s=min(1.0, max(0.05, a.power_s / 10)))
        monitor.start_monitoring(a.device, debug=True, log_path=os.path.join("power_logs", name + ".log"))
    print("Kernel Launched")
    e0.record()
    for _ in range(num_samples):
        for _ in range(rp_time):
            eng.run_dev(idx, xd, cd, yd, alpha, beta, st)
        torch.cuda.current_stream().synchronize()                # one batch in flight at a time, as run.wait() there
    e1.record()
    e1.synchronize()
    print("Kernel Finished")
    if monitor is not None:
        monitor.stop_monitoring()
        print(report(monitor))
    rp_time *= num_samples
    total_ms = e0.elapsed_time(e1)
    t_us = total_ms * 1e3 / rp_time
    print(f"Total Kernel Runtime: {total_ms:.6f}ms")
    print(f"FPGA TIME: {t_us:.6f}us")                              # key names kept for builds/collect_data.py
    print(f"FPGA GFLOPS: {2.0 * (nnz + rows) / (t_us * 1e3):.6f}")
    if monitor is not None and monitor.get_average_power()[1]:
        avg_w = monitor.get_average_power()[0]
        print(f"GFLOPS per Watt: {2.0 * (nnz + rows) / (t_us * 1e3) / avg_w:.6f}")
    # the host-buffer call the plugin makes (copies inside)
    y = np.zeros(rows, np.float32)
    eng.select_matrix(idx)
    eng.run_kernel(x, c_in, y, alpha, beta)
    print()
    print_error_stats(cpu, y)
    if a.dump_y:
        np.save(a.dump_y, y)
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
