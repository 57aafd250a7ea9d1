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
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        eng.run_dev(idx, x, b, y, 0.85, -2.06, st)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times)), float(np.min(times))


def check_sample(eng, idx, spec, x, b, y, nrows=200000):
    """Float64 oracle on the first `nrows` rows (regenerated on the CPU, bit-exact generator)."""
    import oracle_lib as ol
    n = min(nrows, spec.rows)
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
    args = ap.parse_args()
    torch.cuda.init()
    peak = peak_gbs()
    results = []
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for name in args.configs.split(","):
        spec = {"c2": synth.c2_powerlaw, "c4": synth.c4_stencil, "c5": synth.c5_uniform}[name](args.scale)
        t0 = time.time()
        d = synth.DeviceCSR(spec)
        torch.cuda.synchronize()
        t_gen = time.time() - t0
        eng = Engine(0)
        t0 = time.time()
        idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols)
        t_plan = time.time() - t0
        d.close()
        info = eng.matrix_info(idx)
        bytes_alg = 8 * info["nnz"] + 4 * spec.cols + 4 * spec.rows
        print(f"## {spec.name}: rows={spec.rows} nnz={info['nnz']} max_row={info['max_row_nnz']} empty={info['empty_rows']} "
              f"auto={info['kernel_name']}/{info['vector_lanes']} gen={t_gen:.2f}s plan={t_plan:.2f}s "
              f"bytes_alg={bytes_alg/1e6:.1f}MB hist={info['hist'][:22]}", flush=True)
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand(spec.cols, device="cuda", generator=g) + 0.5
        b = torch.rand(spec.rows, device="cuda", generator=g)
        y = torch.empty(spec.rows, device="cuda")
        variants = [("adapt", capi.KERNEL_ADAPTIVE, 0, "A" + a) for a in args.adaptive.split(";") if a]
        variants += [("merge", capi.KERNEL_MERGE, 0, t) for t in args.tiles.split(",") if t]
        variants += [("vector", capi.KERNEL_CSR_VECTOR, l, "") for l in (2, 4, 8, 16, 32)]
        variants += [("scalar", capi.KERNEL_CSR_SCALAR, 0, "")]
        for kname, k, lanes, tile in variants:
            if kname in ("vector", "scalar") and info["max_row_nnz"] > 50000 and lanes != 32:
                continue  # a 1M-nnz row on one thread / a narrow sub-warp would run for seconds
            if tile.startswith("A"):
                os.environ["HISPMV_ADAPTIVE"] = tile[1:]
            elif tile:
                os.environ["HISPMV_MERGE_TILE"] = tile
            try:
                eng.force_kernel(idx, k, lanes)
                med, mn = time_runs(eng, idx, x, b, y, args.iters, flush)
                err = float("nan") if args.no_check else check_sample(eng, idx, spec, x, b, y)
            except Exception as ex:  # noqa: BLE001
                print(f"{spec.name} {kname}{lanes or ''}{('/' + tile) if tile else ''}: FAILED {ex}", flush=True)
                continue
            gbs = bytes_alg / (med * 1e-3) / 1e9
            rec = dict(config=spec.name, kernel=kname, lanes=lanes, tile=tile, ms_med=med, ms_min=mn, gbs=gbs,
                       frac=gbs / peak, gflops=2 * (info["nnz"] + spec.rows) / (med * 1e-3) / 1e9, err=err)
            results.append(rec)
            print(f"{spec.name:14s} {kname:6s} lanes={lanes:2d} tile={tile:16s} med={med:8.4f}ms min={mn:8.4f}ms "
                  f"{gbs:8.1f} GB/s frac={gbs/peak:5.3f} err={err:.2e}", flush=True)
        eng.close()
        os.environ.pop("HISPMV_MERGE_TILE", None)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
