import os, sys, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from hispmv_b200 import Engine, synth
from hispmv_b200.capi import lib, check
spec = synth.c2_powerlaw(1.0)
d = synth.DeviceCSR(spec)
eng = Engine(0)
idx = eng.create_sparse_handle_csr_dev(d.row_ptr, d.col, d.val, spec.rows, spec.cols); d.close()
xh, y0h = synth.reference_vectors(spec.rows, spec.cols)
x = torch.from_numpy(xh).pin_memory(); b = torch.from_numpy(y0h).pin_memory(); y = torch.empty(spec.rows).pin_memory()
eng.select_matrix(idx)
import time
for k in range(4):
    t0 = time.perf_counter()
    check(lib.hispmv_run(eng._ctx, C.c_void_p(x.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(y.data_ptr()), 0.85, -2.06), "run")
    print("call", k, (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr, flush=True)
