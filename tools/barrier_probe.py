"""What one step's x exchange is made of (N ranks, torchrun): the two symmetric-memory barriers on their own, the copy
on its own (multicast stores / peer stores), and the whole sequence.  Device time per step, max over ranks."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hispmv_b200.sharded import XReplicator  # noqa: E402
from hispmv_b200.capi import lib, check  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rep = XReplicator(n, torch.device("cuda", local), mode="multicast")
    x = torch.rand(n, device="cuda")
    s = torch.cuda.Stream()
    lo, hi = rep.slice_bounds()
    hdl = rep._hdl

    def timed(name, fn, iters=200):
        for _ in range(5):
            fn(0)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for k in range(iters):
            fn(k)
        e1.record(s)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:44s} {float(t) * 1e3:8.1f} us per step", flush=True)

    def barriers(k):
        with torch.cuda.stream(s):
            hdl.barrier(channel=k & 1)
            hdl.barrier(channel=2 + (k & 1))

    def one_barrier(k):
        with torch.cuda.stream(s):
            hdl.barrier(channel=k & 1)

    def mc_copy(k):
        mc = hdl.multicast_ptr + ((k & 1) * rep.npad + lo) * 4
        check(lib.hispmv_multicast_copy(C.c_void_p(mc), C.c_void_p(x.data_ptr() + lo * 4), hi - lo, rep.mc_ctas,
                                        C.c_void_p(s.cuda_stream)), "mc")

    def peer_copy(k, ctas=rep.peer_ctas):
        off = ((k & 1) * rep.npad + lo) * 4
        ptrs = (C.c_void_p * world)(*[int(hdl.buffer_ptrs[r]) + off for r in range(world)])
        check(lib.hispmv_peer_copy(ptrs, world, C.c_void_p(x.data_ptr() + lo * 4), hi - lo, ctas, C.c_void_p(s.cuda_stream)), "peer")

    def root_copy(k):
        if rank == 0:
            mc = hdl.multicast_ptr + (k & 1) * rep.npad * 4
            check(lib.hispmv_multicast_copy(C.c_void_p(mc), C.c_void_p(x.data_ptr()), n, rep.mc_ctas, C.c_void_p(s.cuda_stream)), "mc")

    if rank == 0:
        print(f"# {world} GPUs, x = {n * 4 / 1e6:.0f} MB, slice = {(hi - lo) * 4 / 1e6:.1f} MB", flush=True)
    timed("one barrier", one_barrier)
    timed("two barriers", barriers)
    timed("slices, multicast stores (no barrier)", mc_copy)
    for c in (2, 4, 8, 16):
        timed(f"slices, peer stores, {c} CTAs per peer (no barrier)", lambda k, c=c: peer_copy(k, c))
    timed("whole x from rank 0, multicast (no barrier)", root_copy)
    timed("allgather_slices (peer)", lambda k: rep.allgather_slices(k, x, s))
    rep.slice_path = "multicast"
    timed("allgather_slices (multicast)", lambda k: rep.allgather_slices(k, x, s))
    timed("replicate from rank 0", lambda k: rep.replicate(k, x, s))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
