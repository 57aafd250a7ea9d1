"""Two ranks on two B200s of one box (NCCL + NVSwitch multicast), spawned from the test: the exchange kernels and the
collectives fused into the SpMV epilogue.  Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests -m gpu`);
on a one-GPU box the tests skip.  The host-side bookkeeping of the same classes runs over gloo in test_sharded_cpu.py.

Bars: the x replicas are bit-identical to the source; the row-sharded SpMV is within 1e-5 (north_star) of the float64
oracle and bit-identical between the host-buffer and the device-resident call; the chain with the all-gather fused into
the kernel (multimem.st stores) is bit-identical to the chain that calls NCCL's all-gather.
"""
import os
import socket

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs_two = pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs on one box")


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hispmv_b200 import Engine
        from hispmv_b200.capi import lib, check
        from hispmv_b200.layers import DeviceChain, ThreeLayerFCModel, ThreeLayerFCModelConfig
        from hispmv_b200.sharded import RowBlockComm, XReplicator
        dev = torch.device("cuda", rank)
        comm_stream = torch.cuda.Stream()

        # ---- x exchange: replicate from the root, and the sliced host upload, in both modes ----
        n = 1_000_003
        xs = torch.arange(n, dtype=torch.float32) * 0.25 - 1000.0
        modes = []
        for mode in ("auto", "nccl"):
            rep = XReplicator(n, dev, mode=mode)
            modes.append(rep.mode)
            src = xs.cuda() if rank == 0 else None
            rep.replicate(0, src, comm_stream)
            comm_stream.synchronize()
            assert torch.equal(rep.buffer(0).cpu(), xs), mode
            host = xs.clone().pin_memory()
            lo, hi = rep.slice_bounds()
            host[:lo] = float("nan")                      # a rank may read only its own slice of the host vector
            host[hi:] = float("nan")
            sent = rep.gather_from_host(1, host, comm_stream)
            comm_stream.synchronize()
            assert sent == 4 * (hi - lo) and torch.equal(rep.buffer(1).cpu(), xs), mode
            # distributed x: every rank contributes only its own block (multicast slices, peer stores, NCCL all-gather)
            mine = host.cuda()                            # NaN outside this rank's slice
            for path in (("multicast", "peer") if rep.mode == "multicast" else ("nccl",)):
                rep.slice_path = path
                rep.buffer(0).fill_(-1.0)
                torch.cuda.synchronize()
                dist.barrier()
                rep.allgather_slices(0, mine, comm_stream)
                comm_stream.synchronize()
                assert torch.equal(rep.buffer(0).cpu(), xs), (mode, path)
        assert modes[1] == "nccl"

        # ---- row-sharded SpMV: device-resident and host-buffer calls against the oracle ----
        rng = np.random.default_rng(11)                   # same matrix on every rank
        rows, cols = 60000, 50000
        lens = np.minimum(rng.zipf(1.6, rows), 20000)
        lens[[3, 40000]] = 20000
        r = np.repeat(np.arange(rows, dtype=np.int32), lens)
        c = rng.integers(0, cols, r.size).astype(np.int32)
        v = rng.standard_normal(r.size).astype(np.float32)
        x = rng.standard_normal(cols).astype(np.float32)
        b = rng.standard_normal(rows).astype(np.float32)
        rp, ci, vv = ol.coo_to_csr(rows, r, c, v)
        y64, scale = ol.spmv_f64(rp, ci, vv, x, b, np.float32(0.85), np.float32(-2.06))
        eng = Engine(rank, shard=(rank, world))
        idx = eng.create_sparse_handle(r, c, v, rows, cols)
        info = eng.matrix_info(idx)
        rb, re = info["row_begin"], info["row_end"]
        assert np.array_equal(ol.shard_bounds(rp, world)[rank:rank + 2], [rb, re])
        rep = XReplicator(cols, dev, mode="auto")
        rep.replicate(0, torch.from_numpy(x).cuda() if rank == 0 else None, comm_stream)
        comm_stream.synchronize()
        yd = torch.empty(re - rb, device=dev)
        bd = torch.from_numpy(b[rb:re].copy()).cuda()
        eng.run_dev(idx, rep.buffer(0), bd, yd, 0.85, -2.06, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        err, at = ol.max_scaled_error(yd.cpu().numpy(), y64[rb:re], scale[rb:re])
        assert err <= 1e-5, (err, at)
        xh = torch.from_numpy(x).pin_memory()
        bh = torch.from_numpy(b[rb:re].copy()).pin_memory()
        yh = torch.empty(re - rb).pin_memory()
        eng.select_matrix(idx)
        rep.gather_from_host(1, xh, comm_stream)
        check(lib.hispmv_run_xdev(eng._ctx, C.c_void_p(rep.buffer(1).data_ptr()), C.c_void_p(comm_stream.cuda_stream),
                                  C.c_void_p(bh.data_ptr()), C.c_void_p(yh.data_ptr()), 0.85, -2.06), "run_xdev")
        assert torch.equal(yh.view(torch.int32), yd.cpu().view(torch.int32))
        eng.close()

        # ---- chained layers: all-gather fused into the kernel vs NCCL all-gather ----
        torch.manual_seed(0)
        model = ThreeLayerFCModel(ThreeLayerFCModelConfig(input_size=512, dense_size=1536, sparse_size1=2048,
                                                          sparse_size2=300, density1=0.1, density2=0.25)).eval()
        eng = Engine(rank, shard=(rank, world))
        comm = RowBlockComm()
        layers = [model.dense, model.sparse1, model.sparse2]
        plain = DeviceChain(eng, layers, relu=[True, True, True], comm=comm)
        fused = DeviceChain(eng, layers, relu=[True, True, True], comm=comm, fused=True)
        if modes[0] == "multicast":
            assert fused.fused, getattr(fused, "fused_unavailable", "")
        for trial in range(10):
            xin = torch.randn(512)
            with torch.no_grad():
                ref = model(xin.view(1, -1)).numpy().reshape(-1)
            a = plain.forward(xin.cuda()).cpu().numpy()
            f = fused.forward(xin.cuda()).cpu().numpy()
            assert np.abs(a - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
            assert np.array_equal(a.view(np.uint32), f.view(np.uint32))
        eng.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{modes[0]} fused={fused.fused}")
    finally:
        dist.destroy_process_group()


@needs_two
def test_exchange_sharded_spmv_and_fused_chain_on_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    notes = [(tmp_path / f"ok{r}").read_text() for r in range(world)]
    assert len(notes) == world
    print("x exchange mode / fused chain:", notes)
