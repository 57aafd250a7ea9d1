"""Per-kernel SASS evidence for profiles/: which Blackwell instructions each kernel of libhispmv_cuda.so uses.
    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt
UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = bulk L2 prefetch, SYNCS = mbarrier, LDGSTS = cp.async, SHFL / VOTE / REDUX =
warp collectives, ATOMS = shared-memory atomics, MULTIMEM = multimem.st through the NVSwitch multicast address."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "hispmv_b200/libhispmv_cuda.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "REDUX", "ATOMS", "ATOMG", "RED",
        "MULTIMEM", "BAR", "FADD", "FMUL", "FFMA", "UTCHMMA", "HMMA", "LDTM"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["TOTAL"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
        if "MULTIMEM" in line.upper():
            counts[cur]["MULTIMEM"] += 0
demangle = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel (sm_100a); no tensor-core instruction anywhere by design:")
print("# batch-1 SpMV / GeMV is 0.25-0.5 flop per byte (DESIGN.md 5)")
print(f"{'kernel':100s} " + " ".join(f"{w:>7s}" for w in ["TOTAL"] + WANT[:16] + WANT[19:]))
for name, d in zip(order, demangle):
    short = re.sub(r"hispmv::\(anonymous namespace\)::|hispmv::", "", d)[:100]
    c = counts[name]
    if c["TOTAL"] < 40:
        continue
    print(f"{short:100s} " + " ".join(f"{c[w]:7d}" for w in ["TOTAL"] + WANT[:16] + WANT[19:]))
