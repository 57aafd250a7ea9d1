import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
os.environ["HISPMV_PIPELINE"]="1"
from hispmv_b200 import Engine, capi, synth
from hispmv_b200.capi import lib
spec = synth.c2_powerlaw(float(sys.argv[1]) if len(sys.argv)>1 else 0.2)
d = synth.DeviceCSR(spec); eng = Engine(0)
idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols); d.close()
eng.force_kernel(idx, capi.KERNEL_ADAPTIVE)
x = torch.rand(spec.cols, device="cuda"); b = torch.rand(spec.rows, device="cuda"); y = torch.empty(spec.rows, device="cuda")
for _ in range(3):
    eng.run_dev(idx, x, b, y, 0.85, -2.06, 0)
torch.cuda.synchronize()
out = np.zeros(8*512, np.int64)
lib.hispmv_debug_pipe.argtypes=[C.c_void_p]
print("rc", lib.hispmv_debug_pipe(C.c_void_p(out.ctypes.data)))
ev = out.reshape(8,512)
t0 = ev[1,0]
names = ["prod:before_wait_empty","prod:issue","team:full_seen","team:gathered","(unused)","team:released"]
print("k  " + "  ".join(f"{n:>22s}" for n in names))
for k in range(0, 48):
    print(f"{k:2d} " + "  ".join(f"{(ev[e,k]-t0) if ev[e,k] else -1:22d}" for e in range(6)))
k=np.arange(20,200)
print("mean per-tile issue interval", np.diff(ev[1,20:200]).mean(), "full latency", (ev[2,k]-ev[1,k]).mean(), "tile in team", (ev[5,k]-ev[2,k]).mean(), "empty wait", (ev[1,k]-ev[0,k]).mean())
