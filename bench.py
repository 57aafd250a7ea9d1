#!/usr/bin/env python
"""bench.py -- the headline measurement: SpMV GFLOP/s (and HBM GB/s against the roofline) on BASELINE.json's
power-law workload (configs[1], "C2": 10M x 10M, ~100M nnz fp32, highly imbalanced rows).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass y = alpha*A*x + beta*y0 over the whole matrix.
  N = 1   the C2 matrix on one B200.
  N > 1   weak scaling: the matrix has N x 10M rows (same generator, same 10M columns), split into
          nnz-balanced contiguous row blocks, one per rank (one process per GPU, torchrun).  x is produced on
          rank 0 and replicated every step -- one multimem.st store stream to the NVSwitch multicast address
          (hispmv_multicast_copy; NCCL broadcast when no multicast mapping exists or x is larger than 64 MB); the
          exchange of step k+1 runs on a second stream under the SpMV of step k (the reference pipelines
          consecutive vectors the same way, pyhispmv/src/fpga_handle.cpp:366-379).
`value` is device-resident whole-job throughput (CUDA events, max over ranks).  `e2e` is the same metric through
the plugin's host-buffer call (hispmv_run: x and bias from pinned host memory, y back to the host, every step; at
N > 1 every rank sends 1/N of x across PCIe and the slices meet over NVLink, then hispmv_run_xdev).
`--impl reference` times the reference's own CPU path (mkl_sparse_s_mv exactly as cpu/src/main.cpp:26-49 calls
it, compiled unmodified into oracle/_ref) on the host cores, on the same matrix.
"""
from __future__ import annotations

import argparse
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
L2_SECTOR_RATE = 276.9e9  # measured: 32-byte sector requests per second the L2 serves, chip-wide


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(spec, n_gpus):
    if spec.name.startswith("C5"):
        base = f"C5 uniform-random CSR {spec.rows}x{spec.cols} fp32, 6 + popcount(8 bits) nnz per row, seed {spec.seed}"
    else:
        base = (f"C2 power-law CSR {spec.rows}x{spec.cols} fp32, row len ~ min(1M, 0.6912/u), cols ~ Zipf(0.8), "
                f"seed {spec.seed}")
    return base + (f", {n_gpus} nnz-balanced row blocks" if n_gpus > 1 else "")


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
# reference arm: the reference's CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------
def host_matrix(spec, row_begin, row_end):
    import oracle_lib as ol
    return ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, row_begin, row_end)


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
    from hispmv_b200 import synth
    ol.build()
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    if args.workload == "c5":
        full = synth.c5_uniform(args.scale)
        # bounded sample: the first tenth of the rows (rows are independent; 100 M nnz, about 30 ms per MKL call)
        sample_rows = max(1, full.rows // 10)
    else:
        base = synth.c2_powerlaw(args.scale)
        full = synth.SynthSpec(base.name, base.kind, base.seed, base.rows * world, base.cols, base.params)
        sample_rows = base.rows   # bounded sample: the first 10 M rows = exactly the N=1 matrix (rows are hash-generated)
    spec = synth.SynthSpec(full.name, full.kind, full.seed, sample_rows, full.cols, full.params)
    threads = os.cpu_count() or 1
    t0 = time.time()
    rp, ci, vv = host_matrix(full, 0, sample_rows)
    t_gen = time.time() - t0
    x, _ = synth.reference_vectors(spec.rows, spec.cols)
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
              f"hash-generated, so this is the N=1 matrix), {args.steps} mkl_sparse_s_mv calls after {args.warmup} "
              f"warm-up, host-generated in {t_gen:.1f}s")
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
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hispmv_b200 import Engine, synth

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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if not args.no_numa_bind:
            from hispmv_b200.sharded import bind_to_gpu_numa
            affinity = bind_to_gpu_numa(local)      # before any pinned allocation

    if args.workload == "c5":   # BASELINE configs[4]: one 100M x 100M, 1B-nnz matrix split over the ranks (strong scaling)
        spec = synth.c5_uniform(args.scale)
    else:                       # BASELINE configs[1] per GPU (weak scaling)
        base = synth.c2_powerlaw(args.scale)
        spec = synth.SynthSpec(base.name, base.kind, base.seed, base.rows * world, base.cols, base.params)
    if world > 1:
        bounds, total_nnz = synth.synth_shard_bounds(spec, world)
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
    else:
        rb, re, total_nnz = 0, spec.rows, None
    dcsr = synth.DeviceCSR(spec, rb, re)
    eng = Engine(local)
    idx = eng.create_sparse_handle_csr_dev(dcsr.row_ptr, dcsr.col, dcsr.val, re - rb, spec.cols)
    local_nnz = dcsr.nnz
    dcsr.close()
    if total_nnz is None:
        total_nnz = local_nnz
    info = eng.matrix_info(idx)
    n_local = re - rb

    xh, y0h = synth.reference_vectors(spec.rows, spec.cols)
    x_host = torch.from_numpy(xh).pin_memory()
    b_host = torch.from_numpy(y0h[rb:re].copy()).pin_memory()
    y_host = torch.empty(n_local, dtype=torch.float32).pin_memory()
    x_src = x_host.cuda()                     # x as rank 0 produces it
    if world > 1:
        from hispmv_b200.sharded import XReplicator
        xrep = XReplicator(spec.cols, torch.device("cuda", local), mode=args.x_exchange)
        xbuf = [xrep.buffer(0), xrep.buffer(1)]
    else:
        xrep = None
        xbuf = [x_src, x_src]                 # no exchange at N=1: one resident x
    bias = b_host.cuda()
    y = torch.empty(n_local, device="cuda")
    comp = torch.cuda.Stream()
    comm = torch.cuda.Stream(priority=-1)   # NCCL's CTAs take freed SM slots ahead of the SpMV's queued CTAs
    ev_x = [torch.cuda.Event() for _ in range(2)]      # x buffer k is filled
    ev_done = [torch.cuda.Event() for _ in range(2)]   # SpMV reading x buffer k has finished

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def steps_device(n):
        """n pipelined steps; returns nothing, work is on comp/comm streams."""
        for k in range(n):
            cur = k & 1
            if world > 1:
                comm.wait_event(ev_done[cur])              # this rank's replica is free again (SpMV k-2 done)
                xrep.replicate(k, x_src, comm)
                ev_x[cur].record(comm)
                comp.wait_event(ev_x[cur])
            eng.run_dev(idx, xbuf[cur], bias, y, ALPHA, BETA, comp.cuda_stream)
            ev_done[cur].record(comp)

    # -------- device-resident timing -------------------------------------------------------------------
    steps_device(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(comp)
    t_host = time.perf_counter()
    steps_device(args.steps)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps   # CPU time to enqueue one step
    comm.synchronize()
    e1.record(comp)
    barrier()
    ms_total = e0.elapsed_time(e1)

    # -------- the kernel alone (roofline numerator): CUDA events on the launching stream, every step ---------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b_ in kev:
        a.record(comp)
        eng.run_dev(idx, xbuf[0], bias, y, ALPHA, BETA, comp.cuda_stream)
        b_.record(comp)
    barrier()
    k_ms = [a.elapsed_time(b_) for a, b_ in kev]
    kernel_ms = sum(k_ms) / len(k_ms)

    # -------- the x exchange alone (N > 1): NCCL broadcast of x, timed on the communication stream ----------------
    bcast_ms = 0.0
    if world > 1:
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(comm)
        for k in range(args.steps):
            xrep.replicate(k, x_src, comm)
        b1.record(comm)
        barrier()
        bcast_ms = b0.elapsed_time(b1) / args.steps

    # -------- end to end through the host-buffer plugin call ------------------------------------------------
    import ctypes as C
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
    for k in range(args.steps):
        step_e2e(k)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    # the host-buffer path must reproduce the device-resident result bit for bit (same kernel, same x)
    eng.run_dev(idx, x_src if world == 1 else xbuf[0], bias, y, ALPHA, BETA, comp.cuda_stream)
    comp.synchronize()
    e2e_exact = bool(torch.equal(y.cpu().view(torch.int32), y_host.view(torch.int32)))
    clocks = sampler.stop() if sampler else None

    # -------- reduce over ranks ----------------------------------------------------------------------------
    vals = torch.tensor([ms_total, kernel_ms, e2e_ms, bcast_ms, 0.0 if e2e_exact else 1.0], device="cuda",
                        dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms_max, e2e_ms, bcast_ms, e2e_bad = [float(v) for v in vals.tolist()]
    flops_step = 2.0 * (total_nnz + spec.rows)
    ms_step = ms_total / args.steps
    value = flops_step / (ms_step * 1e-3) / 1e9
    e2e_value = flops_step / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_alg_local = 8 * local_nnz + 4 * spec.cols + 4 * n_local      # nnz*(val+idx) + x + y
        achieved = bytes_alg_local / (kernel_ms * 1e-3) / 1e9
        traffic, sectors = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if world == 1 and args.scale == 1.0:
                traffic = tj.get(args.workload + "_" + info["kernel_name"] + "_dram_bytes_per_launch")
                sectors = tj.get(args.workload + "_" + info["kernel_name"] + "_l1_miss_sectors_per_launch")
        except Exception:
            pass
        # second roofline: the L2 slices answer ~277 G sector requests per second chip-wide, however many
        # SMs ask (tools/gather_bench.cu: 276-277 G/s with 37, 74 or 148 SMs busy; 4- or 16-byte payloads alike), and a
        # scattered x gather costs a whole 32-byte sector.  sectors = ncu's L1-miss count for this kernel on this matrix.
        gather_roofline = None
        if sectors:
            min_ms = sectors / L2_SECTOR_RATE * 1e3
            gather_roofline = {"bound": "l2_sector_requests", "sectors_per_launch": sectors,
                               "peak_sectors_per_s": L2_SECTOR_RATE, "min_ms": min_ms, "frac": min_ms / kernel_ms,
                               "peak_source": "tools/gather_bench.cu on B200 (profiles/r1_microbench_stream_gather.txt)",
                               "source": "ncu l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum (profiles/)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.workload == "c5" else "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(spec, world), "rows": spec.rows, "cols": spec.cols, "nnz": int(total_nnz),
                       "alpha": ALPHA, "beta": BETA, "kernel": info["kernel_name"], "tile_items": info["tile_items"],
                       "column_slabs": info.get("num_slabs", 0),
                       "split_rows": info["num_split_rows"],
                       "l2": f"matrix stream is {8 * local_nnz / 1e6:.0f} MB per step per GPU, larger than the 126 MB L2; "
                             f"x ({4 * spec.cols / 1e6:.0f} MB) is the only operand that can stay L2-resident",
                       "x_exchange": "none (N=1)" if world == 1 else (
                           ("one store of x from rank 0 to the NVSwitch multicast address each step "
                            f"(hispmv_multicast_copy, {'copy engine' if xrep.mc_ctas < 0 else str(xrep.mc_ctas or 32) + ' CTAs of multimem.st'}, "
                            "symmetric-memory replicas, two device barriers)"
                            if xrep.mode == "multicast" else "NCCL broadcast of x from rank 0 each step")
                           + ", double-buffered under the previous step's SpMV")},
            "gb_per_s": (8 * total_nnz + 4 * spec.cols + 4 * spec.rows) / (ms_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_alg_local,
                         "launches_per_step": int(eng.launches_per_run(idx)),
                         "note": "rank 0's row block, one launch per step; traffic = ncu dram read+write of the same "
                                 "kernel on the N=1 matrix (profiles/); the binding limit on this matrix is the L2's "
                                 "sector request rate (every scattered x gather costs a 32-byte sector), not HBM (DESIGN.md)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * spec.cols + 4 * spec.rows),
                    "d2h_bytes_per_step": int(4 * spec.rows), "ms_per_step": e2e_ms,
                    "bit_identical_to_device_path": e2e_bad == 0.0, "cpu_affinity_rank0": affinity,
                    "api": "hispmv_run (host x, bias -> host y), pinned host memory" if world == 1 else
                           "per rank: 1/N of the host x up + slices exchanged over NVLink (XReplicator.gather_from_host, "
                           f"{xrep.mode}), then hispmv_run_xdev (host bias block -> host y block), pinned host memory"},
            "gather_roofline": gather_roofline,
            "phases": {"spmv_ms_max_over_ranks": kernel_ms_max, "x_broadcast_ms": bcast_ms,
                       "host_enqueue_ms_per_step": host_enqueue_ms,
                       "spmv_only_gflops": flops_step / (kernel_ms_max * 1e-3) / 1e9,
                       "note": "value includes the per-step x broadcast (pipelined under the previous SpMV); "
                               "spmv_only is the same step with x already resident"},
            "gpu_launches": int(eng.launches_per_run(idx)) * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(eng, idx, spec, xh)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (development only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin each rank to its GPU's local CPUs")
    ap.add_argument("--x-exchange", default="auto", choices=["auto", "multicast", "nccl"],
                    help="N>1: how x reaches every rank each step (auto = NVSwitch multicast if available, else NCCL)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default, the headline: configs[1], weak scaling) or c5 (configs[4], strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
