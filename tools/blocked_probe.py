"""Development probe on one B200: the blocked (two-pass) strategy against the one-pass adaptive kernel and cuSPARSE
(tools/cusparse_ref.cu, the reference's gpu/ baseline call) on the gather-heavy workloads.

    python tools/blocked_probe.py [--workloads c2,c5s,c5shard] [--params "49152,16384,1024,32768;24576,16384,1024,32768"]
                                  [--iters 20] [--out gpurun_out/blocked_probe.json]

Workloads: c2 (BASELINE configs[1]), c5s (C5 at 1/10: 10 M x 10 M uniform), c5shard (rank 0 of 8 of C5: 12.5 M rows x
100 M columns), c5 (the whole 1 B-nnz matrix).  Every variant is checked against the float64 oracle on the first rows.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from hispmv_b200 import Engine, capi, synth  # noqa: E402
from sweep import time_runs, check_sample, peak_gbs  # noqa: E402


def cusparse_ms(d, spec_cols, x, steps=20):
    path = os.path.join(ROOT, "tools", "libcusparse_ref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.cusparse_ref_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                      C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.cusparse_ref_error.restype = C.c_char_p
    y = torch.zeros(d.rows, device="cuda")
    ms, buf = C.c_float(), C.c_int64()
    st = lib.cusparse_ref_spmv(d.row_ptr, d.col, d.val, d.rows, spec_cols, d.nnz, x.data_ptr(), y.data_ptr(), 0.85, 0.0,
                               3, steps, None, C.byref(ms), C.byref(buf))
    if st != 0:
        return {"error": lib.cusparse_ref_error().decode()}
    return {"ms": float(ms.value), "buffer_bytes": int(buf.value)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2,c5s,c5shard")
    ap.add_argument("--params", default="49152,16384,1024,32768")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/blocked_probe.json")
    ap.add_argument("--no-adaptive", action="store_true")
    args = ap.parse_args()
    peak = peak_gbs()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    res = []
    for wl in args.workloads.split(","):
        if wl == "c2":
            spec, rb, re = synth.c2_powerlaw(1.0), 0, None
        elif wl == "c5s":
            spec, rb, re = synth.c5_uniform(0.1), 0, None
        elif wl == "c5":
            spec, rb, re = synth.c5_uniform(1.0), 0, None
        elif wl == "c5shard":
            spec = synth.c5_uniform(1.0)
            bounds, _ = synth.synth_shard_bounds(spec, 8)
            rb, re = int(bounds[0]), int(bounds[1])
        else:
            raise SystemExit(f"unknown workload {wl}")
        re = spec.rows if re is None else re
        d = synth.DeviceCSR(spec, rb, re)
        rows = re - rb
        x = torch.rand(spec.cols, device="cuda") + 0.5
        b = torch.rand(rows, device="cuda")
        y = torch.empty(rows, device="cuda")
        bytes_alg = 8 * d.nnz + 4 * spec.cols + 4 * rows
        cs = cusparse_ms(d, spec.cols, x)
        print(f"## {wl}: rows={rows} cols={spec.cols} nnz={d.nnz} bytes_alg={bytes_alg/1e6:.0f} MB cusparse={cs}", flush=True)
        os.environ["HISPMV_BLOCKED_AUTO"] = "0"
        eng = Engine(0)
        t0 = time.time()
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, rows, spec.cols)
        t_plan = time.time() - t0
        variants = [] if args.no_adaptive else [("one-pass", None)]
        variants += [("blocked", p) for p in args.params.split(";")]
        sub = synth.SynthSpec(spec.name, spec.kind, spec.seed, spec.rows, spec.cols, spec.params)
        for name, p in variants:
            if p is not None:
                os.environ["HISPMV_BLOCKED"] = p
                t0 = time.time()
                eng.force_kernel(idx, capi.KERNEL_BLOCKED)
                t_plan = time.time() - t0
            info = eng.matrix_info(idx)
            med, mn = time_runs(eng, idx, x, b, y, args.iters, flush)
            warm, _ = time_runs(eng, idx, x, b, y, args.iters, None)
            err = check_sample(eng, idx, sub, x, b, y, nrows=100000) if rb == 0 else float("nan")
            ph = [0.0, 0.0]
            if p is not None:   # the two passes on their own (cold L2 each)
                stc = torch.cuda.current_stream().cuda_stream
                for which in (1, 2):
                    ts = []
                    for _ in range(args.iters):
                        flush.add_(1.0)
                        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                        e0.record()
                        eng.run_dev_phase(idx, x, b, y, 0.85, -2.06, which, stc)
                        e1.record()
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ph[which - 1] = float(np.median(ts))
            gbs = bytes_alg / (med * 1e-3) / 1e9
            extra = {}
            if p is not None:
                extra = eng.plan_blocked(idx, arrays=False)
            row = dict(workload=wl, variant=name, params=p, kernel=info["kernel_name"], ms_med=med, ms_min=mn,
                       ms_warm=warm, gbs=gbs, frac=gbs / peak, err=err, plan_s=t_plan, slabs=info["num_slabs"],
                       panels=info["num_tiles"], device_bytes=info["device_bytes"], cusparse=cs, pass1_ms=ph[0],
                       pass2_ms=ph[1], **extra)
            res.append(row)
            print(f"{wl:8s} {name:9s} {str(p):34s} {info['kernel_name']:9s} med={med:8.4f} ms (warm L2 {warm:8.4f}) "
                  f"{gbs:7.0f} GB/s frac={gbs/peak:5.3f} err={err:.1e} plan={t_plan:5.1f}s p1={ph[0]:.3f} p2={ph[1]:.3f} "
                  f"{'' if not extra else 'pieces=%d segs=%d chunks=%d staged=%d smem=%dK' % (extra['num_pieces'], extra['num_seg'], extra['num_chunks'], extra['stage_total'], extra['reduce_words'] * 4 // 1024)}",
                  flush=True)
            json.dump(res, open(args.out, "w"), indent=1)
        eng.close()
        d.close()
        del x, b, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
