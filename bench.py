#!/usr/bin/env python
"""bench.py -- the headline measurement: SpMV GFLOP/s (and HBM GB/s against the roofline) on BASELINE.json's
power-law workload (configs[1], "C2": 10M x 10M, ~100M nnz fp32, highly imbalanced rows), plus -- in the same JSON
line, under "configs" -- the other BASELINE shapes measured in the same run by the same code.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass y = alpha*A*x + beta*y0 over the whole matrix.
  headline (top-level keys)
    N = 1   the C2 matrix on one B200.
    N > 1   weak scaling: the matrix has N x 10M rows (same generator, same 10M columns), split into nnz-balanced
            contiguous row blocks, one per rank (one process per GPU, torchrun).  Every step's x is replicated on all
            ranks: by default every rank contributes its 1/N block (the shape of an SpMV chain, where each rank produces a
            block of the next x; --x-source root: rank 0 holds all of it) through NVSwitch multicast stores, NCCL above
            64 MB; the exchange of step k+1 runs on a second stream under the SpMV of step k.
  "configs"
    c5      BASELINE configs[4]: ONE 100M x 100M, 1B-nnz uniform matrix, strong-scaled over the N ranks (at N = 1 the
            whole matrix on one GPU), x exchanged every step as in the headline.
    c4, c1, gemv8192 (N = 1 only): configs[3], configs[0] and the 8192^2 GeMV of configs[2].
`value` is device-resident whole-job throughput (CUDA events, max over ranks).  `e2e` is the same metric through the
plugin's host-buffer call (hispmv_run: x and bias from pinned host memory, y back to the host, every step; at N > 1
every rank sends 1/N of x across PCIe and the slices meet over NVLink, then hispmv_run_xdev); `e2e_pageable` is the
same call with plain (pageable) numpy arrays, the way pyhispmv.FpgaHandle.run_kernel is called.
`vs_cusparse`: cuSPARSE's generic SpMV -- the reference's GPU baseline call, gpu/src/spmv.cu:83-103, made by
tools/cusparse_ref.cu -- on the same device-resident CSR, same GPU, same run.
`parity`: a seeded sample of every rank's y against the float64 oracle (rows regenerated on the CPU by the oracle's
restatement of the generator); the run FAILS when the north_star bar (1e-5) is exceeded.
`--impl reference` times the reference's own CPU path (mkl_sparse_s_mv exactly as cpu/src/main.cpp:26-49 calls
it, compiled unmodified into oracle/_ref) on the host cores, on the same matrix.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALPHA, BETA = 0.85, -2.06  # cpu/src/main.cpp:147-148
METRIC, UNIT = "spmv_gflops", "GFLOP/s"
TOL = 1e-5                 # north_star: |y - y64| / (|alpha| sum|a_ij x_j| + |beta y0_i|)


def workloads():
    """hispmv_b200/workloads.py loaded on its own: shapes and generator parameters only, no CUDA library."""
    spec = importlib.util.spec_from_file_location("hispmv_workloads", os.path.join(ROOT, "hispmv_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hispmv_workloads"] = mod
    spec.loader.exec_module(mod)
    return mod


_OUT = None


def emit(line) -> None:
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(spec, n_gpus):
    if spec.name.startswith("C5"):
        base = f"C5 uniform-random CSR {spec.rows}x{spec.cols} fp32, 6 + popcount(8 bits) nnz per row, seed {spec.seed}"
    elif spec.name.startswith("C4"):
        base = f"C4 27-point stencil CSR {spec.rows}x{spec.cols} fp32, seed {spec.seed}"
    else:
        base = (f"C2 power-law CSR {spec.rows}x{spec.cols} fp32, row len ~ min(1M, 0.6912/u), cols ~ Zipf(0.8), "
                f"seed {spec.seed}")
        if spec.params[2] >> 8:
            base += f", {spec.rows // (spec.params[2] >> 8)} stacked blocks with C2's row lengths (same nnz per block)"
    return base + (f", {n_gpus} nnz-balanced row blocks" if n_gpus > 1 else "")


class L2Flush:
    """Cold L2 between timed steps of the small configs: 256 MB written (larger than the 126 MB L2), then 256 MB of
    another buffer read, so the write-back of the flush's own dirty lines is over before the timed kernel starts (left
    in L2, ~100 MB of dirty lines drain to DRAM underneath a 30 us kernel: 8 us on the 8192 x 4096 GeMV)."""
    WHAT = "256 MB written then 256 MB read between steps (cold, clean L2)"

    def __init__(self):
        import torch
        self.w = torch.zeros(256 * 1024 * 1024 // 4, device="cuda")
        self.r = torch.zeros(256 * 1024 * 1024 // 4, device="cuda")

    def __call__(self):
        self.w.add_(1.0)
        self.sink = self.r.sum()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on the host cores.  Imports nothing from hispmv_b200 (the package
# loads libhispmv_cuda.so on import); the only native code this arm loads lives under oracle/.
# ----------------------------------------------------------------------------------------------------
def time_reference_mkl(rp, ci, vv, rows, cols, x, steps, warmup, threads):
    """ns per mkl_sparse_s_mv call, measured by the reference's own loop (cpu/src/main.cpp:37-41).  beta = 0 so the
    in-place rp_time loop cannot overflow."""
    import numpy as np
    import oracle_lib as ol
    lib = ol.ref_cpu()
    lib.ref_set_threads(threads)
    y = np.zeros(rows, np.float32)
    if warmup:
        lib.ref_mkl_spmv(rp, ci, vv, rows, cols, ci.size, x, y, ALPHA, 0.0, warmup)
    return lib.ref_mkl_spmv(rp, ci, vv, rows, cols, ci.size, x, y, ALPHA, 0.0, steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import oracle_lib as ol
    wl = workloads()
    ol.build()
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    if args.workload == "c5":
        full = wl.c5_uniform(args.scale)
        # bounded sample: the first tenth of the rows (rows are independent; 100 M nnz, about 30 ms per MKL call)
        sample_rows = max(1, full.rows // 10)
    else:
        base = wl.c2_powerlaw(args.scale)
        full = wl.c2_weak(world, args.scale)
        sample_rows = base.rows   # bounded sample: the first 10 M rows = exactly the N=1 matrix (rows are hash-generated)
    spec = wl.SynthSpec(full.name, full.kind, full.seed, sample_rows, full.cols, full.params)
    threads = os.cpu_count() or 1
    t0 = time.time()
    rp, ci, vv = ol.synth_csr(full.kind, full.seed, full.cols, full.params, 0, sample_rows)
    t_gen = time.time() - t0
    x, _ = wl.reference_vectors(spec.rows, spec.cols)
    kind = "reference" if ol.have_ref() else "port"
    if kind == "reference":
        ns = time_reference_mkl(rp, ci, vv, spec.rows, spec.cols, x, args.steps, args.warmup, threads)
    else:  # the reference could not be compiled here: time the oracle's restatement of cpu_spmv (1 thread)
        threads = 1
        y = np.zeros(spec.rows, np.float32)
        t0 = time.time()
        for _ in range(args.steps):
            ol.oracle().oracle_spmv_csr_f32(spec.rows, rp, ci, vv, x, y, ALPHA, 0.0)
        ns = (time.time() - t0) / args.steps * 1e9
    nnz = int(ci.size)
    gflops = 2.0 * (nnz + spec.rows) / ns
    sample = (f"rows [0, {sample_rows}) of the {full.rows}-row workload ({nnz} nnz; rows are independent and "
              f"hash-generated, so this is the N=1 matrix -- a RATE on one rank's worth of rows, whatever N is), "
              f"{args.steps} mkl_sparse_s_mv calls after {args.warmup} warm-up, beta = 0 (our arm runs beta = -2.06, "
              f"i.e. more work), host-generated in {t_gen:.1f}s")
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ns * 1e-6, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(full, world), "rows": full.rows, "cols": full.cols, "sample_nnz": nnz,
                   "what": "mkl_sparse_s_mv via the reference's mkl_spmv (cpu/src/main.cpp:26-49), libtorch's MKL"
                           if kind == "reference" else "oracle port of cpu_spmv"},
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
class Cusparse:
    """tools/libcusparse_ref.so: the reference's gpu/ baseline call (comparator, not product)."""

    def __init__(self):
        path = os.path.join(ROOT, "tools", "libcusparse_ref.so")
        self.lib = None
        if os.path.exists(path):
            lib = C.CDLL(path)
            lib.cusparse_ref_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                              C.POINTER(C.c_float), C.POINTER(C.c_int64)]
            lib.cusparse_ref_error.restype = C.c_char_p
            lib.cusparse_ref_version.restype = C.c_int
            self.lib = lib

    def time(self, dcsr, cols, x_dev, steps):
        import torch
        if self.lib is None:
            return {"unavailable": "tools/libcusparse_ref.so not built"}
        y = torch.zeros(max(dcsr.rows, 1), device="cuda")
        ms, buf = C.c_float(), C.c_int64()
        st = self.lib.cusparse_ref_spmv(dcsr.row_ptr, dcsr.col, dcsr.val, dcsr.rows, cols, dcsr.nnz, x_dev.data_ptr(),
                                        y.data_ptr(), ALPHA, 0.0, 3, steps, None, C.byref(ms), C.byref(buf))
        if st != 0:
            return {"unavailable": self.lib.cusparse_ref_error().decode()}
        return {"cusparse_ms": float(ms.value), "calls": steps, "version": int(self.lib.cusparse_ref_version()),
                "what": "cusparseSpMV(CUSPARSE_SPMV_ALG_DEFAULT, CSR 32-bit indices, fp32), beta = 0, as "
                        "gpu/src/spmv.cu:83-103 calls it; same device CSR, CUDA events"}


def parity_sample(spec, rb, re, xh, b_local, y_dev, rows_per=20000):
    """Max scaled error of y over three row windows of this rank's block (first, middle, last), the windows' rows
    regenerated on the CPU by the oracle's restatement of the generator and multiplied in float64."""
    import oracle_lib as ol
    n = re - rb
    starts = sorted({0, max(0, n // 2 - rows_per // 2), max(0, n - rows_per)})
    worst, checked = 0.0, 0
    for s0 in starts:
        r0, r1 = rb + s0, min(re, rb + s0 + rows_per)
        if r1 <= r0:
            continue
        rp, ci, vv = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, r0, r1)
        y64, scale = ol.spmv_f64(rp, ci, vv, xh, b_local[s0:s0 + (r1 - r0)], ALPHA, BETA)
        err, _ = ol.max_scaled_error(y_dev[s0:s0 + (r1 - r0)].cpu().numpy(), y64, scale)
        worst, checked = max(worst, err), checked + (r1 - r0)
    return worst, checked


def run_sparse(args, dist_ctx, spec, label, steps, warmup, do_e2e, do_cpu, clocks_for=None):
    """One sparse workload on the ranks of dist_ctx: returns the record (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from hispmv_b200 import Engine, synth
    world, rank, local = dist_ctx
    wl = sys.modules["hispmv_workloads"]
    if world > 1:
        bounds, total_nnz = synth.synth_shard_bounds(spec, world)
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
    else:
        rb, re, total_nnz = 0, spec.rows, None
    dcsr = synth.DeviceCSR(spec, rb, re)
    eng = Engine(local)
    t0 = time.time()
    idx = eng.create_sparse_handle_csr_dev(dcsr.row_ptr, dcsr.col, dcsr.val, re - rb, spec.cols)
    plan_s = time.time() - t0
    local_nnz = dcsr.nnz
    if total_nnz is None:
        total_nnz = local_nnz
    info = eng.matrix_info(idx)
    n_local = re - rb

    xh, y0h = wl.reference_vectors(spec.rows, spec.cols)
    x_host = torch.from_numpy(xh).pin_memory()
    b_local_h = y0h[rb:re].copy()
    b_host = torch.from_numpy(b_local_h).pin_memory()
    y_host = torch.empty(n_local, dtype=torch.float32).pin_memory()
    x_src = x_host.cuda()                     # x as rank 0 produces it
    cus = Cusparse().time(dcsr, spec.cols, x_src, min(steps, 50)) if rank == 0 else None
    dcsr.close()
    x_source = "none"
    if world > 1:
        from hispmv_b200.sharded import XReplicator
        x_source = args.x_source if args.x_source != "auto" else ("distributed" if 4 * spec.cols <= (64 << 20) else "root")
        xrep = XReplicator(spec.cols, torch.device("cuda", local), mode=args.x_exchange,
                           distributed=x_source == "distributed")
        xbuf = [xrep.buffer(0), xrep.buffer(1)]
    else:
        xrep = None
        xbuf = [x_src, x_src]                 # no exchange at N=1: one resident x
    bias = b_host.cuda()
    y = torch.empty(n_local, device="cuda")
    comp = torch.cuda.Stream()
    comm = torch.cuda.Stream(priority=-1)   # the exchange's CTAs take freed SM slots ahead of the SpMV's queued CTAs
    ev_x = [torch.cuda.Event() for _ in range(2)]      # x buffer k is filled
    ev_done = [torch.cuda.Event() for _ in range(2)]   # SpMV reading x buffer k has finished

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    two_pass = info["kernel_name"] == "blocked"
    hold_exchange = os.environ.get("HISPMV_BENCH_HOLD_EXCHANGE", "1") != "0"   # development switch
    ev_p1 = [torch.cuda.Event() for _ in range(2)]     # pass 1 of the step reading x buffer k has finished

    def steps_device(n):
        """n pipelined steps; work is on the comp / comm streams.  The exchange of step k+1 runs under the SpMV of step
        k; with the two-pass strategy it is held back until pass 1 of step k has finished: pass 1 fills every SM's
        shared memory (one CTA per SM), so an exchange CTA that got there first would push a pass-1 CTA into a second
        wave -- under pass 2 (four small CTAs per SM) it fits beside them."""
        for k in range(n):
            cur = k & 1
            if world > 1:
                comm.wait_event(ev_done[cur])              # this rank's replica is free again (SpMV k-2 done)
                if two_pass and k > 0 and hold_exchange:
                    comm.wait_event(ev_p1[cur ^ 1])        # pass 1 of step k-1 is out of the way
                if x_source == "distributed":
                    xrep.allgather_slices(k, x_src, comm)   # every rank contributes its 1/N of x
                else:
                    xrep.replicate(k, x_src, comm)          # rank 0 holds x
                ev_x[cur].record(comm)
                comp.wait_event(ev_x[cur])
            if two_pass and world > 1:
                eng.run_dev_phase(idx, xbuf[cur], bias, y, ALPHA, BETA, 1, comp.cuda_stream)
                ev_p1[cur].record(comp)
                eng.run_dev_phase(idx, xbuf[cur], bias, y, ALPHA, BETA, 2, comp.cuda_stream)
            else:
                eng.run_dev(idx, xbuf[cur], bias, y, ALPHA, BETA, comp.cuda_stream)
            ev_done[cur].record(comp)

    # -------- device-resident timing -------------------------------------------------------------------
    steps_device(max(warmup, 3))
    barrier()
    sampler = ClockSampler(local) if (rank == 0 and clocks_for) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(comp)
    t_host = time.perf_counter()
    steps_device(steps)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / steps   # CPU time to enqueue one step
    comm.synchronize()
    e1.record(comp)
    barrier()
    ms_total = e0.elapsed_time(e1)

    # -------- parity: this rank's y against the float64 oracle on sampled row windows ----------------------------
    perr, pchecked = parity_sample(spec, rb, re, xh, b_local_h, y)

    # -------- the kernels alone (roofline numerator): CUDA events on the launching stream, every step ---------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b_ in kev:
        a.record(comp)
        eng.run_dev(idx, xbuf[0], bias, y, ALPHA, BETA, comp.cuda_stream)
        b_.record(comp)
    barrier()
    k_ms = [a.elapsed_time(b_) for a, b_ in kev]
    kernel_ms = sum(k_ms) / len(k_ms)
    phase_ms = None
    if info["kernel_name"] == "blocked":    # the two passes on their own
        phase_ms = []
        for which in (1, 2):
            pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(steps, 50))]
            for a, b_ in pe:
                a.record(comp)
                eng.run_dev_phase(idx, xbuf[0], bias, y, ALPHA, BETA, which, comp.cuda_stream)
                b_.record(comp)
            barrier()
            phase_ms.append(sum(a.elapsed_time(b_) for a, b_ in pe) / len(pe))

    # -------- the x exchange alone (N > 1), timed on the communication stream ----------------
    bcast_ms = 0.0
    if world > 1:
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(comm)
        for k in range(steps):
            if x_source == "distributed":
                xrep.allgather_slices(k, x_src, comm)
            else:
                xrep.replicate(k, x_src, comm)
        b1.record(comm)
        barrier()
        bcast_ms = b0.elapsed_time(b1) / steps

    # -------- end to end through the host-buffer plugin call ------------------------------------------------
    e2e_ms, e2e_exact, pageable_ms = 0.0, True, None
    if do_e2e:
        from hispmv_b200.capi import lib, check
        eng.select_matrix(idx)

        def step_e2e(k):
            if world > 1:
                # x is the same host vector on every rank: each rank carries 1/N of it across PCIe and the slices meet
                # over NVLink in every rank's replica; bias and y are this rank's row block (pipelined inside the call)
                xrep.gather_from_host(k, x_host, comm)
                check(lib.hispmv_run_xdev(eng._ctx, C.c_void_p(xbuf[k & 1].data_ptr()), C.c_void_p(comm.cuda_stream),
                                          C.c_void_p(b_host.data_ptr()), C.c_void_p(y_host.data_ptr()), ALPHA, BETA),
                      "hispmv_run_xdev")
            else:
                check(lib.hispmv_run(eng._ctx, C.c_void_p(x_host.data_ptr()), C.c_void_p(b_host.data_ptr()),
                                     C.c_void_p(y_host.data_ptr()), ALPHA, BETA), "hispmv_run")

        for k in range(3):
            step_e2e(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            step_e2e(k)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        # the host-buffer path must reproduce the device-resident result bit for bit (same kernels, same x)
        eng.run_dev(idx, x_src if world == 1 else xbuf[0], bias, y, ALPHA, BETA, comp.cuda_stream)
        comp.synchronize()
        e2e_exact = bool(torch.equal(y.cpu().view(torch.int32), y_host.view(torch.int32)))
        if world == 1:
            # the same call the way the plugin's callers make it: plain (pageable) numpy arrays
            y_np = np.empty(n_local, np.float32)
            n_pg = max(3, min(steps, 20))
            for _ in range(2):
                eng.run_kernel(xh, b_local_h, y_np, ALPHA, BETA)
            t0 = time.perf_counter()
            for _ in range(n_pg):
                eng.run_kernel(xh, b_local_h, y_np, ALPHA, BETA)
            pageable_ms = (time.perf_counter() - t0) * 1e3 / n_pg
            e2e_exact = e2e_exact and bool(np.array_equal(y_np.view(np.int32), y_host.numpy().view(np.int32)))
    clocks = sampler.stop() if sampler else None

    # -------- reduce over ranks ----------------------------------------------------------------------------
    vals = torch.tensor([ms_total, kernel_ms, e2e_ms, bcast_ms, 0.0 if e2e_exact else 1.0, perr], device="cuda",
                        dtype=torch.float64)
    cnt = torch.tensor([float(pchecked)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, kernel_ms_max, e2e_ms, bcast_ms, e2e_bad, perr = [float(v) for v in vals.tolist()]
    flops_step = 2.0 * (total_nnz + spec.rows)
    ms_step = ms_total / steps
    rec = None
    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_alg_local = 8 * local_nnz + 4 * spec.cols + 4 * n_local      # nnz*(val+idx) + x + y
        achieved = bytes_alg_local / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if world == 1 and args.scale == 1.0:
                traffic = tj.get(label + "_" + info["kernel_name"] + "_dram_bytes_per_launch")
        except Exception:
            pass
        launches = int(eng.launches_per_run(idx))
        rec = {
            "value": flops_step / (ms_step * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_step, "steps": steps,
            "config": {"workload": workload_name(spec, world), "rows": spec.rows, "cols": spec.cols, "nnz": int(total_nnz),
                       "alpha": ALPHA, "beta": BETA, "kernel": info["kernel_name"], "tile_items": info["tile_items"],
                       "column_slabs": info.get("num_slabs", 0), "split_rows": info["num_split_rows"],
                       "plan_seconds": plan_s,
                       "l2": f"matrix stream is {8 * local_nnz / 1e6:.0f} MB per step per GPU, larger than the 126 MB L2 "
                             "(no flush needed between steps)",
                       "x_exchange": "none (N=1)" if world == 1 else (
                           (("every rank stores its 1/N block of x" if x_source == "distributed" else
                             "one store of x from rank 0") + " to the NVSwitch multicast address each step "
                            f"(hispmv_multicast_copy, {'copy engine' if xrep.mc_ctas < 0 else str(xrep.slice_cta_count() if x_source == 'distributed' else (xrep.mc_ctas or 32)) + ' CTAs of multimem.st'}, "
                            "symmetric-memory replicas, two device barriers)"
                            if xrep.mode == "multicast" else
                            ("NCCL all-gather of the ranks' blocks of x each step" if x_source == "distributed" else
                             "NCCL broadcast of x from rank 0 each step"))
                           + ", double-buffered under the previous step's SpMV")},
            "gb_per_s": (8 * total_nnz + 4 * spec.cols + 4 * spec.rows) / (ms_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_alg_local, "launches_per_step": launches,
                         "note": "rank 0's row block; kernel_ms = all launches of one step (CUDA events on the launching "
                                 "stream); achieved = algorithmic bytes of the SpMV / kernel_ms; traffic = ncu dram "
                                 "read+write of the same launches on the N=1 matrix (profiles/)"},
            "parity": {"rows_checked": int(cnt.item()), "max_scaled_err": perr, "tolerance": TOL,
                       "against": "float64 oracle on three row windows per rank, rows regenerated on the CPU"},
            "vs_cusparse": None,
            "phases": {"spmv_ms_max_over_ranks": kernel_ms_max, "x_broadcast_ms": bcast_ms,
                       "host_enqueue_ms_per_step": host_enqueue_ms,
                       "spmv_only_gflops": flops_step / (kernel_ms_max * 1e-3) / 1e9,
                       "note": "value includes the per-step x exchange (pipelined under the previous SpMV); "
                               "spmv_only is the same step with x already resident"},
            "gpu_launches": launches * steps,
        }
        if phase_ms:
            pb = eng.plan_blocked(idx, arrays=False)
            b1 = 6.25 * pb["padded_nnz"] + 4 * pb["num_pieces"] + 4 * spec.cols
            # pass 2: staged partial sums (aligned covers) + their 16-bit places, segment descriptors, end marks, and per
            # row bias + piece extent in, y out
            b2 = 6 * pb["stage_total"] + 8 * pb["num_seg"] + 4 * pb["bit_words"] + 12 * n_local
            rec["kernels"] = [
                {"name": "pb_expand_kernel", "ms": phase_ms[0], "actual_bytes": b1,
                 "actual_gbs": b1 / (phase_ms[0] * 1e-3) / 1e9, "frac_of_peak": b1 / (phase_ms[0] * 1e-3) / 1e9 / peak},
                {"name": "pb_reduce_kernel", "ms": phase_ms[1], "actual_bytes": b2,
                 "actual_gbs": b2 / (phase_ms[1] * 1e-3) / 1e9, "frac_of_peak": b2 / (phase_ms[1] * 1e-3) / 1e9 / peak}]
            rec["config"]["pieces"] = pb["num_pieces"]
            rec["config"]["segments"] = pb["num_seg"]
            rec["config"]["slabs"] = pb["num_slabs"]
        if cus and "cusparse_ms" in cus:
            cus["ours_ms"] = kernel_ms
            cus["speedup"] = cus["cusparse_ms"] / kernel_ms
        rec["vs_cusparse"] = cus
        if do_e2e:
            rec["e2e"] = {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                          "h2d_bytes_per_step": int(4 * spec.cols + 4 * spec.rows),
                          "d2h_bytes_per_step": int(4 * spec.rows), "ms_per_step": e2e_ms,
                          "bit_identical_to_device_path": e2e_bad == 0.0,
                          "api": "hispmv_run (host x, bias -> host y), pinned host memory" if world == 1 else
                                 "per rank: 1/N of the host x up + slices exchanged over NVLink (XReplicator."
                                 f"gather_from_host, {xrep.mode}), then hispmv_run_xdev (host bias block -> host y block), "
                                 "pinned host memory"}
            if pageable_ms:
                rec["e2e_pageable"] = {"value": flops_step / (pageable_ms * 1e-3) / 1e9, "unit": UNIT,
                                       "ms_per_step": pageable_ms,
                                       "api": "Engine.run_kernel = hispmv_run with plain numpy arrays (pageable memory), the "
                                              "way pyhispmv.FpgaHandle.run_kernel is called"}
        if clocks is not None:
            rec["clocks"] = clocks
        if do_cpu:
            rec["cpu_baseline"] = cpu_baseline(eng, idx, spec, xh)
    eng.close()
    del x_src, bias, y, xbuf
    torch.cuda.empty_cache()
    return rec


def run_gemv(args, local, rows, cols, steps):
    """The 8192^2 GeMV of configs[2] with the reference's closed-form matrix (cpu/src/main.cpp:213-218)."""
    import numpy as np
    import torch
    from hispmv_b200 import Engine
    i = torch.arange(rows, device="cuda", dtype=torch.float32).unsqueeze(1)
    j = torch.arange(cols, device="cuda", dtype=torch.float32).unsqueeze(0)
    a = ((i + 1) / (j + 2)).contiguous()
    x = ((torch.arange(cols, device="cuda", dtype=torch.float32) + 1) / (torch.arange(cols, device="cuda", dtype=torch.float32) + 2))
    b = -2.0 * (torch.arange(rows, device="cuda", dtype=torch.float32) + 1) / (torch.arange(rows, device="cuda", dtype=torch.float32) + 2)
    y = torch.empty(rows, device="cuda")
    eng = Engine(local)
    idx = eng.create_dense_handle_dev(a, rows, cols)
    st = torch.cuda.current_stream().cuda_stream
    flush = L2Flush()
    for _ in range(3):
        eng.run_dev(idx, x, b, y, ALPHA, BETA, st)
    ts = []
    for _ in range(steps):
        flush()            # 268 MB of A would partly stay in the 126 MB L2: cold L2 between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_dev(idx, x, b, y, ALPHA, BETA, st)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ref = ALPHA * (a.double() @ x.double()) + BETA * b.double()
    scale = abs(ALPHA) * (a.double().abs() @ x.double().abs()) + abs(BETA) * b.double().abs()
    err = float(((y.double() - ref).abs() / scale).max())
    peak, _ = measured_peak()
    bytes_alg = 4 * rows * cols + 4 * cols + 4 * rows
    eng.close()
    return {"value": (2.0 * rows * cols + rows) / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "config": {"workload": f"dense {rows}x{cols} fp32 GeMV, A[i,j]=(i+1)/(j+2) (cpu/src/main.cpp:213-218)",
                       "kernel": "gemv", "l2": L2Flush.WHAT},
            "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak, "kernel_ms": ms,
                         "algorithmic_bytes_per_launch": bytes_alg},
            "parity": {"rows_checked": rows, "max_scaled_err": err, "tolerance": TOL,
                       "against": "float64 torch product on the device"}}


def run_c1(args, local, steps):
    """configs[0]: the 65,536^2 imbalanced matrix through the plugin's COO entry, checked against the oracle in full."""
    import numpy as np
    import torch
    import oracle_lib as ol
    from hispmv_b200 import Engine, synth
    r, c, v, n, _ = synth.c1_imbalanced_coo()
    eng = Engine(local)
    idx = eng.create_sparse_handle(r, c, v, n, n)
    info = eng.matrix_info(idx)
    wl = sys.modules["hispmv_workloads"]
    xh, y0h = wl.reference_vectors(n, n)
    x, b = torch.from_numpy(xh).cuda(), torch.from_numpy(y0h).cuda()
    y = torch.empty(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    flush = L2Flush()
    for _ in range(3):
        eng.run_dev(idx, x, b, y, ALPHA, BETA, st)
    cold, warm = [], []
    for it in range(2 * steps):
        if it < steps:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_dev(idx, x, b, y, ALPHA, BETA, st)
        e1.record()
        e1.synchronize()
        (cold if it < steps else warm).append(e0.elapsed_time(e1))
    ms, ms_warm = float(np.median(cold)), float(np.median(warm))
    rp, ci, vv = eng.plan_csr(idx)
    y64, scale = ol.spmv_f64(rp, ci, vv, xh, y0h, ALPHA, BETA)
    err, _ = ol.max_scaled_error(y.cpu().numpy(), y64, scale)
    peak, _ = measured_peak()
    bytes_alg = 8 * info["nnz"] + 8 * n
    eng.close()
    return {"value": 2.0 * (info["nnz"] + n) / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "ms_warm_l2": ms_warm,
            "steps": steps,
            "config": {"workload": f"C1 imbalanced CSR {n}x{n} fp32, {info['nnz']} nnz (power-law rows + 4 dense rows)",
                       "kernel": info["kernel_name"], "l2": L2Flush.WHAT + " (ms_warm_l2: without)"},
            "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak, "kernel_ms": ms,
                         "algorithmic_bytes_per_launch": bytes_alg},
            "parity": {"rows_checked": n, "max_scaled_err": err, "tolerance": TOL, "against": "float64 oracle, every row"}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    affinity = "not requested"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if not args.no_numa_bind:
            from hispmv_b200.sharded import bind_to_gpu_numa
            affinity = bind_to_gpu_numa(local)      # before any pinned allocation
    wl = workloads()
    ctx = (world, rank, local)

    if args.workload == "c5":   # BASELINE configs[4] as the headline (development switch)
        spec = wl.c5_uniform(args.scale)
        scaling = "strong"
    else:                       # BASELINE configs[1] per GPU (weak scaling)
        spec = wl.c2_weak(world, args.scale)   # N stacked blocks of C2's shape, 100.0 M nonzeros each
        scaling = "weak"
    main = run_sparse(args, ctx, spec, args.workload, args.steps, args.warmup, True, world == 1 and not args.no_cpu,
                      clocks_for=True)

    extras = {}
    want = [w for w in args.configs.split(",") if w] if args.configs != "auto" else (
        ["c5"] if (world > 1 or args.workload == "c5") else ["c5", "c4", "gemv8192", "c1"])
    if args.workload == "c5":
        want = [w for w in want if w != "c5"]
    k = max(5, min(args.steps, 30))
    for name in want:
        try:
            if name == "c5":
                rec = run_sparse(args, ctx, wl.c5_uniform(args.scale), "c5", k, 3, False, False)
                if rec:
                    rec["scaling"] = "strong"
            elif name == "c4" and world == 1:
                rec = run_sparse(args, ctx, wl.c4_stencil(args.scale), "c4", k, 3, False, False)
            elif name == "gemv8192" and world == 1:
                rec = run_gemv(args, local, 8192, 8192, k)
            elif name == "c1" and world == 1:
                rec = run_c1(args, local, k)
            else:
                continue
        except Exception as ex:  # noqa: BLE001  -- a side workload must not take the headline down with it
            rec = {"failed": f"{type(ex).__name__}: {ex}"} if rank == 0 else None
        if rank == 0 and rec is not None:
            extras[name] = rec

    bad = []
    if rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        for key in ("config", "gb_per_s", "roofline", "kernels", "e2e", "e2e_pageable", "parity", "vs_cusparse", "phases",
                    "gpu_launches", "clocks", "cpu_baseline"):
            if key in main:
                line[key] = main[key]
        line["e2e"]["cpu_affinity_rank0"] = affinity
        line["configs"] = extras
        for name, rec in [("headline", main)] + list(extras.items()):
            p = rec.get("parity") if isinstance(rec, dict) else None
            if p and not (p["max_scaled_err"] <= TOL):
                bad.append(f"{name}: max scaled error {p['max_scaled_err']:.3g} > {TOL}")
            if isinstance(rec, dict) and rec.get("e2e") and not rec["e2e"]["bit_identical_to_device_path"]:
                bad.append(f"{name}: host-buffer path differs from the device-resident path")
        line["parity_ok"] = not bad
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit("PARITY FAILURE: " + "; ".join(bad))


def cpu_baseline(eng, idx, spec, xh):
    """The reference's MKL call on this box's host cores, on the very matrix the GPU just ran (copied back)."""
    import oracle_lib as ol
    import numpy as np
    try:
        ol.build()
        rp, ci, vv = eng.plan_csr(idx)
        threads = os.cpu_count() or 1
        if ol.have_ref():
            calls = 10
            ns = time_reference_mkl(rp, ci, vv, spec.rows, spec.cols, xh, calls, 2, threads)
            kind = "reference"
        else:
            calls, threads, kind = 2, 1, "port"
            y = np.zeros(spec.rows, np.float32)
            t0 = time.time()
            for _ in range(calls):
                ol.oracle().oracle_spmv_csr_f32(spec.rows, rp, ci, vv, xh, y, ALPHA, 0.0)
            ns = (time.time() - t0) / calls * 1e9
        return {"value": 2.0 * (ci.size + spec.rows) / ns, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": f"full matrix, {calls} mkl_sparse_s_mv calls after 2 warm-up, {ns * 1e-6:.1f} ms each"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"{type(ex).__name__}: {ex}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workloads (development only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin each rank to its GPU's local CPUs")
    ap.add_argument("--x-exchange", default="auto", choices=["auto", "multicast", "nccl"],
                    help="N>1: how x reaches every rank each step (auto = NVSwitch multicast if available, else NCCL)")
    ap.add_argument("--x-source", default="auto", choices=["auto", "distributed", "root"],
                    help="N>1: where each step's x comes from: 'distributed' = every rank holds 1/N of it (the shape of an "
                         "SpMV chain: each rank produces a block of the next x) and the blocks are all-gathered; 'root' = "
                         "rank 0 holds all of it and replicates it (round 1's exchange); 'auto' = distributed up to 64 MB "
                         "of x, root above (measured at N=8: C2 6669 vs 6035 GFLOP/s, C5 1425 vs 1578)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="the headline: c2 (default: configs[1], weak scaling) or c5 (configs[4], strong scaling)")
    ap.add_argument("--configs", default="auto",
                    help="the other BASELINE shapes reported under 'configs': auto (c5 at every N; c4, gemv8192, c1 too at "
                         "N=1), a comma list, or '' for none")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its
    # version there) are sent to stderr, the line itself goes to a private copy of the original stdout
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
