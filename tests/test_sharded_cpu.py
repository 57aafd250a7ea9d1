"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-block bookkeeping, the padded all-gather of unequal
y blocks, and x broadcast.  The per-rank SpMV is injected from the oracle here (test infrastructure) -- on the GPU box
it is libhispmv_cuda.so (tests/test_gpu_parity.py::test_row_block_shards_reassemble, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hispmv_b200.sharded import RowBlockComm
        rng = np.random.default_rng(3)                      # same matrix on every rank
        rows, cols = 5000, 3000
        lens = np.minimum(rng.zipf(1.7, rows), 2500)
        lens[7] = 2500
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
        rp, ci, vv = ol.coo_to_csr(rows, r, c, v)
        bounds = ol.shard_bounds(rp, world)                 # nnz-balanced split points (bit-exact contract)
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
        comm = RowBlockComm()
        got = comm.set_blocks("y", rb, re, torch.device("cpu"))
        assert np.array_equal(got, bounds)
        # x lives on rank 0 only
        x = torch.from_numpy(rng.standard_normal(cols).astype(np.float32))
        if rank != 0:
            x.zero_()
        comm.broadcast_x(x, 0)
        bias = rng.standard_normal(rows).astype(np.float32)
        # local block through the oracle's fp32 CSR SpMV
        lrp = (rp[rb:re + 1] - rp[rb]).astype(np.int32)
        y_local = bias[rb:re].copy()
        ol.oracle().oracle_spmv_csr_f32(re - rb, lrp, ci[rp[rb]:rp[re]].copy(), vv[rp[rb]:rp[re]].copy(),
                                        x.numpy(), y_local, 0.85, -2.06)
        full = comm.allgather_rows("y", torch.from_numpy(y_local), torch.empty(rows))
        y_ref = bias.copy()
        ol.oracle().oracle_spmv_csr_f32(rows, rp, ci, vv, x.numpy(), y_ref, 0.85, -2.06)
        assert np.array_equal(full.numpy().view(np.uint32), y_ref.view(np.uint32))
        # a second vector with a different (empty-block) layout reuses the same communicator
        b2 = [0, 0, 11][: world + 1] if world == 2 else None
        if b2:
            comm.set_blocks("z", b2[rank], b2[rank + 1], torch.device("cpu"))
            z = comm.allgather_rows("z", torch.arange(b2[rank], b2[rank + 1], dtype=torch.float32), torch.empty(11))
            assert torch.equal(z, torch.arange(11, dtype=torch.float32))
        # x exchange: replication from the root, and the sliced host upload (every rank carries 1/world of x)
        from hispmv_b200.sharded import XReplicator
        n = 1003                                            # not a multiple of 4 * world: ragged last slice
        rep = XReplicator(n, torch.device("cpu"), mode="nccl")
        assert rep.mode == "nccl" and rep.buffer(0).numel() == n and rep.buffer(1).data_ptr() != rep.buffer(0).data_ptr()
        cuts = [rep.slice_bounds(q) for q in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(lo % 4 == 0 for lo, _ in cuts)
        xs = torch.arange(n, dtype=torch.float32) * 0.5 - 7
        rep.replicate(0, xs if rank == 0 else None, None)
        assert torch.equal(rep.buffer(0), xs)
        poisoned = xs.clone()
        lo, hi = rep.slice_bounds()
        poisoned[:lo] = float("nan")                        # a rank may only read its own slice of the host vector
        poisoned[hi:] = float("nan")
        sent = rep.gather_from_host(1, poisoned, None)
        assert sent == 4 * (hi - lo) and torch.equal(rep.buffer(1), xs)
        rep.buffer(0).zero_()                               # distributed x: every rank contributes only its own block
        rep.allgather_slices(0, poisoned, None)
        assert torch.equal(rep.buffer(0), xs)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_row_block_collectives_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_set_blocks_rejects_gaps():
    mp.spawn(_gap_worker, args=(2, _free_port()), nprocs=2, join=True)


def _gap_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hispmv_b200.sharded import RowBlockComm
        comm = RowBlockComm()
        try:
            comm.set_blocks("bad", 0 if rank == 0 else 6, 5 if rank == 0 else 9, torch.device("cpu"))
        except ValueError:
            return
        raise AssertionError("gap between row blocks was accepted")
    finally:
        dist.destroy_process_group()
