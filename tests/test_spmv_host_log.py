"""The log of tools/spmv_host.py (our spmv-host) through the reference's own scraper patterns (builds/collect_data.py:8-23,
restated here): a maintainer's CSV collection must keep working on the GPU engine's logs.  tests/golden/spmv_host_log.txt
is a log recorded on a B200 (`python tools/spmv_host.py <c1-like>.mtx --exec_ms 200 --power_s 1`)."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))

# builds/collect_data.py:8-21 (METRIC_PATTERNS) and :23 (SAMPLE_PATTERN)
METRIC_PATTERNS = {
    "Pre-Processing Time": r"Pre-processing Time: ([\d\.]+) secs",
    "CPU Time": r"CPU TIME: ([\d\.]+) ms",
    "CPU GFLOPS": r"CPU GFLOPS: ([\d\.]+)",
    "Matrix A Length": r"Matrix A Length: (\d+)",
    "Approx. Clock Cycles": r"Approx\. Clock Cycles: (\d+)",
    "Repeat Time": r"Using Repeat Time: (\d+)",
    "Num Samples": r"Using Num samples: (\d+)",
    "Average Power": r"Average Power: ([\d\.]+) Watts",
    "Max Power": r"Max Power: ([\d\.]+) Watts",
    "Total Kernel Runtime": r"Total Kernel Runtime: ([\d\.]+)ms",
    "FPGA Time": r"FPGA TIME: ([\d\.]+)us",
    "FPGA GFLOPS": r"FPGA GFLOPS: ([\d\.]+)",
}
SAMPLE_PATTERN = r"sample: ([\d\.]+)"


def test_every_metric_the_reference_scrapes_is_in_our_log():
    content = open(os.path.join(HERE, "golden", "spmv_host_log.txt")).read()
    entry = {}
    for metric, pattern in METRIC_PATTERNS.items():
        m = re.search(pattern, content)
        assert m, f"'{metric}' is missing from the log"
        entry[metric] = float(m.group(1))
    samples = [float(s) for s in re.findall(SAMPLE_PATTERN, content)]
    assert len(samples) >= 2 and all(20.0 < w < 1500.0 for w in samples)
    assert min(samples) <= entry["Average Power"] <= entry["Max Power"] == max(samples)
    assert entry["Matrix A Length"] > 0 and entry["Repeat Time"] >= 1 and entry["Num Samples"] >= 1
    # FPGA GFLOPS line = 2 (nnz + rows) / time (common/src/spmv-host.cpp:185); rows is not in the log, nnz is a lower bound
    assert entry["FPGA GFLOPS"] >= 2.0 * entry["Matrix A Length"] / (entry["FPGA Time"] * 1e3) * 0.999
    assert "Kernel Launched" in content and "Kernel Finished" in content
    assert re.search(r"Relative Error Range:|No mismatch found|Found atmost 10 mismatches", content)
