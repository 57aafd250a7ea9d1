"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked summaries under profiles/.

    python tools/make_profiles.py --round r1 --launches gpurun_out/launches_r1.csv \
        --full gpurun_out/bench_c2_adaptive_r1.ncu-rep:c2_adaptive [--full other.ncu-rep:tag ...]

Writes profiles/<round>_launches.txt (every launch of the bench command with its device time and share),
profiles/<round>_<tag>_ncu.txt (the metrics that matter for a bandwidth-bound kernel) and updates
profiles/traffic.json (dram read+write bytes per launch, which bench.py reports as roofline.traffic).
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import WANT  # noqa: E402


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    iname, imetric, ival, iunit = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    recs = [(r[iname], float(r[ival].replace(",", "")), r[iunit]) for r in rows[1:] if r[imetric] == "gpu__time_duration.sum"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    recs = [(n, v * scale.get(u, 1.0)) for n, v, u in recs]
    tot = sum(v for _, v in recs)
    agg = {}
    for n, v in recs:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {len(recs)} launches, {tot:.1f} us in total\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'launches':>8s} {'total_us':>12s} {'share':>7s}  kernel\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{c:8d} {v:12.1f} {100 * v / tot:6.1f}%  {n[:150]}\n")
        f.write("\n# in launch order\n")
        for n, v in recs:
            f.write(f"{v:12.1f} us  {n[:150]}\n")


def full(rep, tag, rnd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    traffic = {}
    add = tag.endswith("+")        # "tag+": the report holds the launches of ONE step -- their traffic is added up
    tag = tag.rstrip("+")
    with open(os.path.join(ROOT, "profiles", f"{rnd}_{tag}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none : {os.path.basename(rep)}\n")
        for vals in rows[2:]:
            name = vals[h.index("Kernel Name")]
            f.write(f"== {name[:160]}\n")
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    f.write(f"  {w:84s} {vals[i]:>18s} {units[i]}\n")
            def get(m):
                i = h.index(m)
                return float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            traffic[tag + "_dram_bytes_per_launch"] = (traffic.get(tag + "_dram_bytes_per_launch", 0.0) if add else 0.0) + \
                get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
            m = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum"
            if m in h:
                traffic[tag + "_l1_miss_sectors_per_launch"] = float(vals[h.index(m)].replace(",", ""))
    return traffic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r1")
    ap.add_argument("--launches", default="")
    ap.add_argument("--full", action="append", default=[])
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        launches(a.launches, os.path.join(ROOT, "profiles", f"{a.round}_launches.txt"))
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for spec in a.full:
        rep, tag = spec.rsplit(":", 1)
        tj.update(full(rep, tag, a.round))
    json.dump(tj, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
