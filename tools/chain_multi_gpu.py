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
    layers = [model.dense, model.sparse1, model.sparse2]
    chain = DeviceChain(eng, layers, relu=[True, True, True], comm=comm)
    fused = DeviceChain(eng, layers, relu=[True, True, True], comm=comm, fused=True)
    ok, same = True, True
    for trial in range(20):
        x = torch.randn(4096)
        with torch.no_grad():
            ref = model(x.view(1, -1)).numpy().reshape(-1)
        out = chain.forward(x.cuda()).cpu().numpy()
        # three un-normalised randn layers: outputs are O(1e3) sums of cancelling terms, so the bar is relative to
        # the largest output (general_test.py's own check is rtol=1e-3 on such sums, apps/general_test.py:106)
        ok = ok and bool(np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max())
        out_f = fused.forward(x.cuda()).cpu().numpy()
        same = same and bool(np.array_equal(out.view(np.uint32), out_f.view(np.uint32)))   # same kernels, same bits
    xd = torch.randn(4096, device="cuda")
    times = {}
    for name, ch in (("nccl all-gather", chain), ("fused multicast store" if fused.fused else "fused(unavailable)", fused)):
        for _ in range(5):
            ch.forward(xd)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200
        e0.record()
        for _ in range(iters):
            ch.forward(xd)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = float(t)
    flag = torch.tensor([1 if ok else 0, 1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        blocks = [(b[1] - b[0]) for b in chain.blocks]
        print(f"chain_multi_gpu: world={world} parity={'ok' if int(flag[0]) else 'FAILED'} "
              f"fused==allgather bit-for-bit={'yes' if int(flag[1]) else 'NO'} (fused path active: {fused.fused}"
              f"{'' if fused.fused else ' -- ' + getattr(fused, 'fused_unavailable', '')}) rows per rank {blocks}")
        for name, us in times.items():
            print(f"  {name:24s} {us:8.1f} us per forward pass (3 layers, device time, max over ranks)")
    flag = flag.min().reshape(1)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
