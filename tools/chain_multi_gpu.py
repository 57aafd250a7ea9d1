"""model_test's MLP sharded over the GPUs of one box (torchrun, one process per GPU): every rank holds a row block of
each layer, y blocks are all-gathered into the next layer's replicated x (hispmv_b200.sharded / layers.DeviceChain).
Checks the result against the CPU model on every rank and prints per-layer / per-pass timings from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/chain_multi_gpu.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hispmv_b200 import Engine  # noqa: E402
from hispmv_b200.layers import DeviceChain, ThreeLayerFCModel, ThreeLayerFCModelConfig  # noqa: E402
from hispmv_b200.sharded import RowBlockComm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                 # same weights on every rank (apps/model_test.py defaults)
    model = ThreeLayerFCModel(ThreeLayerFCModelConfig()).eval()
    for p in model.parameters():
        p.requires_grad = False
    eng = Engine(local, shard=(rank, world))
    comm = RowBlockComm()
    chain = DeviceChain(eng, [model.dense, model.sparse1, model.sparse2], relu=[True, True, True], comm=comm)
    ok = True
    for trial in range(3):
        x = torch.randn(4096)
        with torch.no_grad():
            ref = model(x.view(1, -1)).numpy().reshape(-1)
        out = chain.forward(x.cuda()).cpu().numpy()
        ok = ok and bool(np.allclose(out, ref, rtol=1e-3, atol=1e-4))
    xd = torch.randn(4096, device="cuda")
    for _ in range(5):
        chain.forward(xd)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    iters = 200
    for _ in range(iters):
        chain.forward(xd)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters * 1e6
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        blocks = [(b[1] - b[0]) for b in chain.blocks]
        print(f"chain_multi_gpu: world={world} parity={'ok' if int(flag) else 'FAILED'} rows per rank {blocks} "
              f"{dt:.1f} us per forward pass (3 layers + {sum(1 for b, s in zip(chain.blocks, chain.shapes) if b[1]-b[0] != s[0])} all-gathers)")
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
