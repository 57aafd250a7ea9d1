import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hispmv_b200.synth import c1_imbalanced_coo, write_mtx
r, c, v, n, _ = c1_imbalanced_coo(n=16384, target_nnz=200000, dense_rows=3, dense_len=6000)
write_mtx(sys.argv[1] if len(sys.argv) > 1 else "c1_small.mtx", r, c, v, n, n)
print("wrote", r.size)
