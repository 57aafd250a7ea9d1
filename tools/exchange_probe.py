"""Probe of the x-exchange options on N GPUs (torchrun): NCCL broadcast vs copy-engine pulls from peer-mapped memory vs
one multicast store over NVSwitch (hispmv_multicast_copy).  Prints availability and per-exchange times for a 40 MB x."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hispmv_b200.capi import lib, check  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    src = torch.arange(n, device="cuda", dtype=torch.float32) + 1 if rank == 0 else torch.zeros(n, device="cuda")
    xs = symm_mem.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(xs, group=dist.group.WORLD)
    mc = bool(hdl.multicast_ptr)
    if rank == 0:
        print(f"world={world} n={n} multicast_support={mc} mc_ptr={hex(hdl.multicast_ptr) if mc else None}", flush=True)
    st = torch.cuda.current_stream()
    # (a) NCCL broadcast
    t_nccl = timed(lambda: dist.broadcast(src, src=0))
    # (b) copy-engine pull: every rank copies from rank 0's symmetric buffer
    if rank == 0:
        xs.copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    root_view = hdl.get_buffer(0, (n,), torch.float32)
    dst = torch.empty(n, device="cuda")

    def pull():
        hdl.barrier(channel=0)
        if rank != 0:
            dst.copy_(root_view)
    t_pull = timed(pull)
    ok_pull = bool(rank == 0 or torch.equal(dst, torch.arange(n, device="cuda", dtype=torch.float32) + 1))
    # (c) multicast store from rank 0
    t_mc, ok_mc = float("nan"), None
    if mc:
        xs.zero_()
        torch.cuda.synchronize()
        dist.barrier()

        t_by_ctas = {}
        for ctas in (-1, 16, 32, 64):
            def mcast():
                hdl.barrier(channel=1)              # everyone is done reading the previous x
                if rank == 0:
                    check(lib.hispmv_multicast_copy(C.c_void_p(hdl.multicast_ptr), C.c_void_p(src.data_ptr()), n, ctas,
                                                    C.c_void_p(st.cuda_stream)), "multicast_copy")
                hdl.barrier(channel=2)              # x has landed everywhere
            t_by_ctas[ctas] = timed(mcast)
        if rank == 0:
            print("multicast(+2 barriers) by CTAs: " + "  ".join(f"{c}: {t:.4f} ms ({4 * n / t / 1e6:.0f} GB/s)"
                                                                for c, t in t_by_ctas.items()), flush=True)
        t_mc = min(t_by_ctas.values())
        ok_mc = bool(torch.equal(xs, torch.arange(n, device="cuda", dtype=torch.float32) + 1))
    flags = torch.tensor([int(ok_pull), int(ok_mc) if ok_mc is not None else 1], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"nccl_broadcast {t_nccl:.4f} ms | ce_pull(+barrier) {t_pull:.4f} ms ok={bool(flags[0])} | "
              f"multicast(+2 barriers) {t_mc:.4f} ms ok={bool(flags[1]) if mc else None}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
