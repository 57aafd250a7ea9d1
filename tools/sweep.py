"""Kernel sweep on one B200: every SpMV strategy on the synthetic configs, timed with CUDA events.

Usage (on the GPU box):  python tools/sweep.py [--scale 1.0] [--configs c2,c4,c5,c3] [--out gpurun_out/sweep.json]
Prints one line per (config, kernel) with ms, GB/s of algorithmic bytes, fraction of the measured HBM peak and
the max scaled error against the float64 oracle on a sampled row block.  Development tool, not the bench.
"""
from __future__ import annotations

import argparse
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


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_runs(eng, idx, x, b, y, iters, flush):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush[0].add_(1.0)       # cold L2 ...
            _ = flush[1].sum()       # ... and clean: the flush's dirty lines are written back before the timed kernel
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times)), float(np.min(times))


def check_sample(eng, idx, spec, x, b, y, nrows=200000, host_csr=None):
    """Float64 oracle on the first `nrows` rows (regenerated on the CPU, bit-exact generator)."""
    import oracle_lib as ol
    n = min(nrows, spec.rows)
    if host_csr is not None:
        n = spec.rows
        rp, ci, vv = host_csr
    else:
        rp, ci, vv = ol.synth_csr(spec.kind, spec.seed, spec.cols, spec.params, 0, n)
    xh, bh = x.cpu().numpy(), b[:n].cpu().numpy()
    y64, scale = ol.spmv_f64(rp, ci, vv, xh, bh, 0.85, -2.06)
    err, at = ol.max_scaled_error(y[:n].cpu().numpy(), y64, scale)
    return err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="c2,c4,c5")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--tiles", default="896,1792,2816,3584")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--adaptive", default="2048,1024,4096", help="';'-separated B,T,CH triples")
    ap.add_argument("--rowstage", default="", help="';'-separated LANES,B,T,CH quadruples ('auto' = planner default)")
    ap.add_argument("--hot", default="0", help="','-separated hot-column thresholds for the split L1 policy (0 = keep all)")
    ap.add_argument("--persist", default="", help="','-separated x-window sizes for the persistent adaptive kernel")
    ap.add_argument("--pipeline", default="", help="';'-separated B,T,CH triples for the warp-specialised pipeline kernel")
    ap.add_argument("--warptile", default="", help="';'-separated B,T,CH triples for the warp-per-tile kernel")
    ap.add_argument("--only-auto", action="store_true", help="time only what the selector picks")
    ap.add_argument("--skip", default="", help="','-separated variant names to skip (adapt,rows,merge,vector,scalar)")
    args = ap.parse_args()
    torch.cuda.init()
    peak = peak_gbs()
    results = []
    if os.environ.get("HISPMV_NOCHECK"):
        args.no_check = True
    flush = (torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda"),
             torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda"))
    for name in args.configs.split(","):
        host_csr = None
        t0 = time.time()
        eng = Engine(0)
        if name in ("g8192", "g3a", "gtall"):
            # dense overlay: 8192^2 (cpu/run_gemv.sh shape), the MLP's dense layer, general_test.py's 50000 x 10000
            nr, nc = {"g8192": (8192, 8192), "g3a": (8192, 4096), "gtall": (50000, 10000)}[name]
            a = torch.rand(nr, nc, device="cuda") - 0.5
            idx = eng.create_dense_handle_dev(a, nr, nc)
            bytes_alg = 4 * nr * nc + 4 * nc + 4 * nr
            x = torch.rand(nc, device="cuda") + 0.5
            b = torch.rand(nr, device="cuda")
            y = torch.empty(nr, device="cuda")
            med, mn = time_runs(eng, idx, x, b, y, args.iters, flush)
            ref = 0.85 * (a.double() @ x.double()) - 2.06 * b.double()
            scale = 0.85 * (a.double().abs() @ x.double().abs()) + 2.06 * b.double().abs()
            err = float(((y.double() - ref).abs() / scale).max())
            gbs = bytes_alg / (med * 1e-3) / 1e9
            results.append(dict(config=name, kernel="gemv", lanes=0, tile="", ms_med=med, ms_min=mn, gbs=gbs, frac=gbs / peak,
                                gflops=(2 * nr * nc + nr) / (med * 1e-3) / 1e9, err=err, rows=nr, cols=nc, nnz=nr * nc))
            print(f"## {name}: dense {nr}x{nc} bytes_alg={bytes_alg/1e6:.1f}MB", flush=True)
            print(f"{name:14s} gemv   lanes= 0 tile={'':26s} med={med:8.4f}ms min={mn:8.4f}ms {gbs:8.1f} GB/s "
                  f"frac={gbs/peak:5.3f} err={err:.2e}", flush=True)
            del a
            eng.close()
            continue
        if name in ("c2", "c4", "c5"):
            spec = {"c2": synth.c2_powerlaw, "c4": synth.c4_stencil, "c5": synth.c5_uniform}[name](args.scale)
            d = synth.DeviceCSR(spec)
            torch.cuda.synchronize()
            t_gen = time.time() - t0
            t0 = time.time()
            idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
            t_plan = time.time() - t0
            d.close()
        else:  # small host-built configs: C1 and the model_test MLP's sparse layers (apps/model.py:10-44)
            if name == "c1":
                r, c, v, nr, nc = synth.c1_imbalanced_coo()
            else:
                nr, nc, dens = {"c3b": (8192, 8192, 0.1), "c3c": (1024, 8192, 0.25)}[name]
                g0 = torch.Generator().manual_seed(0)
                w = torch.randn(nr, nc, generator=g0) * (torch.rand(nr, nc, generator=g0) < dens)
                nzr, nzc = torch.nonzero(w, as_tuple=True)
                r, c, v = nzr.numpy().astype(np.int32), nzc.numpy().astype(np.int32), w[nzr, nzc].numpy()
            spec = synth.SynthSpec(name.upper(), 0, 0, nr, nc, (0, 0, 0))
            t_gen = time.time() - t0
            t0 = time.time()
            idx = eng.create_sparse_handle(r, c, v, nr, nc)
            t_plan = time.time() - t0
            host_csr = eng.plan_csr(idx)
        info = eng.matrix_info(idx)
        bytes_alg = 8 * info["nnz"] + 4 * spec.cols + 4 * spec.rows
        print(f"## {spec.name}: rows={spec.rows} nnz={info['nnz']} max_row={info['max_row_nnz']} empty={info['empty_rows']} "
              f"auto={info['kernel_name']}/{info['vector_lanes']} gen={t_gen:.2f}s plan={t_plan:.2f}s "
              f"bytes_alg={bytes_alg/1e6:.1f}MB hist={info['hist'][:22]}", flush=True)
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand(spec.cols, device="cuda", generator=g) + 0.5
        b = torch.rand(spec.rows, device="cuda", generator=g)
        y = torch.empty(spec.rows, device="cuda")
        variants = []
        for h in args.hot.split(","):
            hs = "" if h in ("", "0") else "H" + h
            variants += [("adapt", capi.KERNEL_ADAPTIVE, 0, "A" + a + hs + "P0") for a in args.adaptive.split(";") if a]
            variants += [("rows", capi.KERNEL_ROWSTAGE, 0 if q == "auto" else int(q.split(",")[0]), "R" + q + hs)
                         for q in args.rowstage.split(";") if q]
        variants += [("persist", capi.KERNEL_ADAPTIVE, 0, "A" + a + "P" + w) for a in args.adaptive.split(";") if a
                     for w in args.persist.split(",") if w]
        variants += [("pipe", capi.KERNEL_ADAPTIVE, 0, "A" + a + "Q") for a in args.pipeline.split(";") if a]
        variants += [("warpt", capi.KERNEL_ADAPTIVE, 0, "W" + a) for a in args.warptile.split(";") if a]
        variants += [("auto", capi.KERNEL_AUTO, 0, "")]
        variants += [("merge", capi.KERNEL_MERGE, 0, t) for t in args.tiles.split(",") if t]
        variants += [("vector", capi.KERNEL_CSR_VECTOR, l, "") for l in (2, 4, 8, 16, 32)]
        variants += [("scalar", capi.KERNEL_CSR_SCALAR, 0, "")]
        skip = set(args.skip.split(","))
        if args.only_auto:
            variants = [("auto", capi.KERNEL_AUTO, 0, "")]
        for kname, k, lanes, tile in variants:
            if kname in skip:
                continue
            if kname in ("vector", "scalar") and info["max_row_nnz"] > 50000 and lanes != 32:
                continue  # a 1M-nnz row on one thread / a narrow sub-warp would run for seconds
            for var in ("HISPMV_ADAPTIVE", "HISPMV_ROWSTAGE", "HISPMV_HOT", "HISPMV_MERGE_TILE", "HISPMV_PERSIST",
                        "HISPMV_PIPELINE", "HISPMV_WARPTILE"):
                os.environ.pop(var, None)
            spec_s = tile
            if spec_s.startswith("W"):
                os.environ["HISPMV_WARPTILE"] = spec_s[1:]
                spec_s = ""
            if spec_s.endswith("Q"):
                spec_s = spec_s[:-1]
                os.environ["HISPMV_PIPELINE"] = "1"
            if "P" in spec_s:
                spec_s, win = spec_s.split("P")
                os.environ["HISPMV_PERSIST"] = win
            if "H" in spec_s:
                spec_s, hot = spec_s.split("H")
                os.environ["HISPMV_HOT"] = hot
            if spec_s.startswith("A"):
                os.environ["HISPMV_ADAPTIVE"] = spec_s[1:]
            elif spec_s.startswith("R"):
                if spec_s[1:] != "auto":
                    os.environ["HISPMV_ROWSTAGE"] = spec_s[1:]
            elif spec_s:
                os.environ["HISPMV_MERGE_TILE"] = spec_s
            try:
                eng.force_kernel(idx, k, lanes)
                med, mn = time_runs(eng, idx, x, b, y, args.iters, flush)
                err = float("nan") if args.no_check else check_sample(eng, idx, spec, x, b, y, host_csr=host_csr)
            except Exception as ex:  # noqa: BLE001
                print(f"{spec.name} {kname}{lanes or ''}{('/' + tile) if tile else ''}: FAILED {ex}", flush=True)
                continue
            gbs = bytes_alg / (med * 1e-3) / 1e9
            chosen = eng.matrix_info(idx)
            rec = dict(config=spec.name, kernel=kname, lanes=lanes, tile=tile, ms_med=med, ms_min=mn, gbs=gbs,
                       frac=gbs / peak, gflops=2 * (info["nnz"] + spec.rows) / (med * 1e-3) / 1e9, err=err,
                       rows=spec.rows, cols=spec.cols, nnz=info["nnz"], planned=chosen["kernel_name"],
                       planned_lanes=chosen["vector_lanes"], tile_items=chosen["tile_items"])
            results.append(rec)
            print(f"{spec.name:14s} {kname:6s} lanes={lanes:2d} tile={tile:26s} med={med:8.4f}ms min={mn:8.4f}ms "
                  f"{gbs:8.1f} GB/s frac={gbs/peak:5.3f} err={err:.2e}", flush=True)
        eng.close()
        os.environ.pop("HISPMV_MERGE_TILE", None)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
