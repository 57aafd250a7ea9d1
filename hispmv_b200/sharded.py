"""One process per GPU: nnz-balanced row blocks, x replicated by broadcast, y blocks all-gathered.

The reference is single-device (one xrt::device, pyhispmv/src/fpga_handle.cpp:55); this module is the multi-GPU
design the north star adds.  Rows are independent, so the SpMV itself needs no collective: rank r owns the
contiguous row block [bounds[r], bounds[r+1]) of every matrix (split points = lower_bound(row_ptr, k*nnz/G), rows are
never split across GPUs; dense matrices use equal row blocks) and a full copy of x.  Two exchange steps exist:
  * x replication   x produced on one rank (or held by every rank in host memory) -> all ranks: XReplicator, one
                    multimem.st store stream to the NVSwitch multicast address (hispmv_multicast_copy), NCCL
                    broadcast / all-gather as the fallback and for x above 64 MB
  * allgather_rows  chained layers: every rank's y block -> the next layer's replicated x.  Blocks have unequal
                    row counts, so they travel padded to the largest block and are compacted by one gather.
                    layers.DeviceChain(fused=True) removes this step: the kernel stores y through the multicast
                    address of the next layer's x (hispmv_run_dev_mc).
torch.distributed supplies the process group (NCCL on GPUs; the same code runs over gloo on CPU tensors, which is how
tests/test_sharded_cpu.py covers the bookkeeping).  All arithmetic stays in libhispmv_cuda.so.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


class RowBlockComm:
    """Collective plumbing for row-block sharded vectors."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._plans: Dict[Hashable, dict] = {}

    def set_blocks(self, key: Hashable, row_begin: int, row_end: int, device) -> np.ndarray:
        """Register this rank's block of vector `key`; returns the bounds of all ranks (world+1 entries)."""
        mine = torch.tensor([row_begin, row_end], dtype=torch.int64, device=device)
        allb = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allb, mine, group=self.group)
        pairs = torch.stack(allb).cpu().numpy()
        bounds = np.concatenate([pairs[:, 0], pairs[-1:, 1]])
        if not (np.all(pairs[1:, 0] == pairs[:-1, 1]) and bounds[0] == 0):
            raise ValueError(f"row blocks do not tile the vector: {pairs.tolist()}")
        counts = np.diff(bounds)
        pad = int(counts.max()) if counts.size else 0
        # position of row i inside the padded (world x pad) receive buffer
        index = np.concatenate([r * pad + np.arange(counts[r]) for r in range(self.world)]) if pad else np.zeros(0)
        self._plans[key] = {
            "bounds": bounds, "pad": pad,
            "index": torch.from_numpy(index.astype(np.int64)).to(device),
            "send": torch.zeros(max(pad, 1), dtype=torch.float32, device=device),
            "recv": torch.zeros(max(pad, 1) * self.world, dtype=torch.float32, device=device),
        }
        return bounds

    def bounds(self, key: Hashable) -> np.ndarray:
        return self._plans[key]["bounds"]

    def allgather_rows(self, idx_layer: Hashable, y_local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[bounds[r]:bounds[r+1]] = rank r's y_local, on every rank."""
        p = self._plans[idx_layer]
        n = y_local.numel()
        p["send"][:n].copy_(y_local)
        chunks = list(p["recv"].view(self.world, -1).unbind(0))
        dist.all_gather(chunks, p["send"], group=self.group)
        torch.index_select(p["recv"], 0, p["index"], out=out)
        return out

    def broadcast_x(self, x: torch.Tensor, src: int = 0) -> torch.Tensor:
        dist.broadcast(x, src=src, group=self.group)
        return x


def bind_to_gpu_numa(device_index: int) -> str:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index`, so that pinned host buffers
    allocated afterwards (first touch) sit on the GPU's NUMA node and host <-> device copies do not cross sockets.
    One process per GPU makes this the natural place.  Returns a short description; never raises."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"unchanged ({len(allowed)} cpus, no narrower GPU-local set)"
        os.sched_setaffinity(0, cpus)
        return f"bound to {len(cpus)} GPU-local cpus ({min(cpus)}-{max(cpus)})"
    except Exception as ex:  # noqa: BLE001
        return f"unchanged ({type(ex).__name__}: {ex})"


class XReplicator:
    """x produced on one rank -> a replica on every rank, double-buffered so that the exchange of step k+1 runs under
    the SpMV of step k.

    mode "multicast": the replicas live in symmetric memory (torch.distributed._symmetric_memory: allocation and
    rendezvous only) and the root stores x ONCE to the NVSwitch multicast address with hispmv_multicast_copy
    (multimem.st): the switch delivers it to every GPU, no SM on the receivers runs anything, and the root's NVLink
    egress carries x once instead of once per peer.  Two small device-side barriers per exchange order it against the
    readers.  mode "nccl": dist.broadcast.  "auto": multicast for x up to AUTO_MULTICAST_BYTES, NCCL above or when
    the multicast mapping is unavailable.

    Measured on 8 B200 (profiles/r1_exchange_probe.txt): one root reaches ~330-400 GB/s through the multicast address
    (copy engine or 32+ CTAs alike; 560 GB/s with one peer), NCCL's broadcast 400-630 GB/s.  For C2's 40 MB x the
    multicast store still gives the faster step (4233 vs 3992 GFLOP/s: it takes no SM and no receiver-side kernel away
    from the SpMV it overlaps); for C5's 400 MB x the transfer itself dominates and NCCL wins (0.64 vs 1.0 ms)."""

    AUTO_MULTICAST_BYTES = 64 << 20

    def __init__(self, n: int, device, group=None, mode: str = "auto", root: int = 0, mc_ctas: Optional[int] = None,
                 distributed: bool = False):
        import os
        self.n, self.device, self.group, self.root = n, device, group, root
        self.rank = dist.get_rank(group)
        self.npad = (n + 3) & ~3
        self.mode = "nccl"
        self._hdl = None
        # who issues the multicast stores: N > 0 = N CTAs of multimem.st, -1 = a copy engine.  Inside the pipelined
        # step 16 CTAs are the best measured (N=2, C2: 1086 GFLOP/s; 32 CTAs 1082, 8 CTAs 1074, copy engine 984 --
        # its transfer is as fast in isolation but overlaps the SpMV worse; profiles/r1_exchange_probe.txt)
        self.mc_ctas = int(os.environ.get("HISPMV_MC_CTAS", "16")) if mc_ctas is None else mc_ctas
        # allgather_slices: "peer" = every rank stores its block into every replica with plain stores (hispmv_peer_copy),
        # "multicast" = one multimem.st stream per rank
        # Measured at N=8 on C2 (profiles/r2_n8_exchange_sweep.txt): multicast slices from 4 CTAs per rank 0.2646 ms per
        # step, from 16 CTAs 0.2658, peer stores 0.31-0.40, one root 0.2916.
        self.slice_path = os.environ.get("HISPMV_SLICE_PATH", "multicast")
        self.slice_ctas = int(os.environ.get("HISPMV_SLICE_CTAS", "0"))     # 0: by slice size, see allgather_slices
        self.peer_ctas = int(os.environ.get("HISPMV_PEER_CTAS", "0")) or max(2, 16 // max(1, dist.get_world_size(group)))
        # the size rule is about ONE root pushing the whole vector; blocks of a distributed x go through multicast at any size
        if mode == "auto" and 4 * n > self.AUTO_MULTICAST_BYTES and not distributed:
            mode = "nccl"
        if mode in ("auto", "multicast"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(2 * self.npad, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(buf, group=group if group is not None else dist.group.WORLD)
                if not hdl.multicast_ptr:
                    raise RuntimeError("no multicast mapping (NVLS unavailable)")
                self._hdl, self._buf, self.mode = hdl, buf, "multicast"
            except Exception as ex:  # noqa: BLE001
                if mode == "multicast":
                    raise
                self._why = f"{type(ex).__name__}: {ex}"
        if self.mode == "nccl":
            self._buf = torch.empty(2 * self.npad, dtype=torch.float32, device=device)
        self._buf.zero_()

    def buffer(self, k: int) -> torch.Tensor:
        """This rank's replica k (k & 1): what the SpMV of step k reads."""
        o = (k & 1) * self.npad
        return self._buf[o:o + self.n]

    @staticmethod
    def _on(stream):
        import contextlib
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def replicate(self, k: int, src_on_root: Optional[torch.Tensor], stream: Optional[torch.cuda.Stream]) -> None:
        """Enqueue on `stream` the exchange that fills replica k on every rank from `src_on_root` (root only).
        stream=None: current stream / CPU tensors over gloo (tests)."""
        with self._on(stream):
            if self.mode == "nccl":
                b = self.buffer(k)
                if self.rank == self.root and src_on_root is not None and src_on_root.data_ptr() != b.data_ptr():
                    b.copy_(src_on_root)
                dist.broadcast(b, src=self.root, group=self.group)
                return
            import ctypes as C
            from .capi import lib, check
            cur = k & 1
            self._hdl.barrier(channel=cur)          # every rank is done reading replica `cur`
            if self.rank == self.root:
                mc = self._hdl.multicast_ptr + cur * self.npad * 4
                check(lib.hispmv_multicast_copy(C.c_void_p(mc), C.c_void_p(src_on_root.data_ptr()), self.n, self.mc_ctas,
                                                C.c_void_p(stream.cuda_stream)), "multicast_copy")
            self._hdl.barrier(channel=2 + cur)      # the stores have landed everywhere

    def slice_cta_count(self) -> int:
        """CTAs of multimem.st per rank for allgather_slices: one per ~1.25 MB of the slice, 4..16 (few CTAs, short
        kernel: at N=8 four CTAs per rank beat sixteen)."""
        lo, hi = self.slice_bounds()
        return self.slice_ctas or max(4, min(16, -(-(hi - lo) * 4 // (1280 * 1024))))

    def allgather_slices(self, k: int, x_local: torch.Tensor, stream: Optional[torch.cuda.Stream]) -> None:
        """Enqueue on `stream` the exchange that fills replica k on every rank when x is DISTRIBUTED: rank r holds (at
        least) elements slice_bounds(r) of x in `x_local` (a full-length device vector of which only that slice is
        read) -- the shape of an SpMV chain, where every rank produces a block of the next x.  Every rank stores its
        slice once to the multicast address (multimem.st): per-rank egress is n/N, every GPU's ingress n (N-1)/N, and
        no single root's link carries the whole vector (one root tops out at ~330-400 GB/s through the switch at N=8;
        eight roots of 1/8 each finish in about a third of the time).  mode "nccl": all_gather_into_tensor."""
        lo, hi = self.slice_bounds()
        with self._on(stream):
            if self.mode == "nccl":
                world = dist.get_world_size(self.group)
                per = self.slice_bounds(0)[1]
                if not hasattr(self, "_mine"):
                    self._gath = torch.empty(world * max(per, 4), dtype=torch.float32, device=self.device)
                    self._mine = torch.zeros(max(per, 4), dtype=torch.float32, device=self.device)
                if hi > lo:
                    self._mine[:hi - lo].copy_(x_local[lo:hi])
                dist.all_gather_into_tensor(self._gath, self._mine, group=self.group)
                self.buffer(k).copy_(self._gath[:self.n])
                return
            import ctypes as C
            from .capi import lib, check
            cur = k & 1
            self._hdl.barrier(channel=cur)          # every rank is done reading replica `cur`
            if hi > lo:
                off = (cur * self.npad + lo) * 4
                if self.slice_path == "peer":       # unicast stores into every replica (this rank's included)
                    world = dist.get_world_size(self.group)
                    ptrs = (C.c_void_p * world)(*[int(self._hdl.buffer_ptrs[r]) + off for r in range(world)])
                    check(lib.hispmv_peer_copy(ptrs, world, C.c_void_p(x_local.data_ptr() + lo * 4), hi - lo,
                                               self.peer_ctas, C.c_void_p(stream.cuda_stream)), "peer_copy")
                else:
                    ctas = self.slice_cta_count()
                    check(lib.hispmv_multicast_copy(C.c_void_p(self._hdl.multicast_ptr + off),
                                                    C.c_void_p(x_local.data_ptr() + lo * 4), hi - lo, ctas,
                                                    C.c_void_p(stream.cuda_stream)), "multicast_copy")
            self._hdl.barrier(channel=2 + cur)      # every slice has landed everywhere

    # ---- host-resident x: every rank holds the same x in (pinned) host memory --------------------------------
    def slice_bounds(self, rank: Optional[int] = None):
        """[lo, hi) of x that `rank` carries across PCIe: equal 16-byte-aligned slices."""
        world = dist.get_world_size(self.group)
        per = (((self.n + world - 1) // world) + 3) & ~3
        r = self.rank if rank is None else rank
        return min(self.n, r * per), min(self.n, (r + 1) * per)

    def gather_from_host(self, k: int, x_host: torch.Tensor, stream: Optional[torch.cuda.Stream]) -> int:
        """Enqueue on `stream`: this rank copies ONLY its slice of the (pinned) host x to the GPU and the slices meet
        in replica k of every rank over NVLink -- one multimem.st store per rank to the multicast address (mode
        "multicast") or an NCCL all-gather.  x crosses PCIe once per node instead of once per GPU.  Returns the bytes
        this rank sent host -> device."""
        lo, hi = self.slice_bounds()
        cur = k & 1
        world = dist.get_world_size(self.group)
        if not hasattr(self, "_stage"):
            per = self.slice_bounds(0)[1]
            self._stage = torch.empty(max(per, 4), dtype=torch.float32, device=self.device)
            if self.mode == "nccl":
                self._gath = torch.empty(world * max(per, 4), dtype=torch.float32, device=self.device)
        with self._on(stream):
            if hi > lo:
                self._stage[:hi - lo].copy_(x_host[lo:hi], non_blocking=True)
            if self.mode == "nccl":
                dist.all_gather_into_tensor(self._gath, self._stage, group=self.group)
                self.buffer(k).copy_(self._gath[:self.n])
                return (hi - lo) * 4
            import ctypes as C
            from .capi import lib, check
            self._hdl.barrier(channel=cur)          # every rank is done reading replica `cur`
            if hi > lo:
                mc = self._hdl.multicast_ptr + (cur * self.npad + lo) * 4
                check(lib.hispmv_multicast_copy(C.c_void_p(mc), C.c_void_p(self._stage.data_ptr()), hi - lo,
                                                self.mc_ctas, C.c_void_p(stream.cuda_stream)), "multicast_copy")
            self._hdl.barrier(channel=2 + cur)      # every slice has landed everywhere
        return (hi - lo) * 4


class ShardedEngine:
    """An Engine that keeps row block `rank` of `world` of every matrix, plus the collectives around it."""

    def __init__(self, local_device: int, group=None, **engine_kwargs):
        from .engine import Engine
        self.comm = RowBlockComm(group)
        self.engine = Engine(local_device, shard=(self.comm.rank, self.comm.world), **engine_kwargs)
        self.device = torch.device("cuda", local_device)
        self._y: Dict[int, torch.Tensor] = {}

    def add(self, idx: int) -> int:
        """Register matrix idx (already created on self.engine) for gathers of its y."""
        info = self.engine.matrix_info(idx)
        self.comm.set_blocks(idx, info["row_begin"], info["row_end"], self.device)
        self._y[idx] = torch.empty(info["row_end"] - info["row_begin"], device=self.device)
        return idx

    def spmv(self, idx: int, x: torch.Tensor, bias_local: Optional[torch.Tensor], alpha: float = 1.0,
             beta: float = 0.0, x_root: Optional[int] = None, gather: bool = False, stream: int = 0):
        """y_block = alpha * A_block x + beta * bias_block.  x_root: rank whose x is broadcast first (None: x is
        already replicated).  gather=True returns the full y on every rank, else this rank's block."""
        if x_root is not None:
            self.comm.broadcast_x(x, x_root)
        y = self._y[idx]
        self.engine.run_dev(idx, x, bias_local, y, alpha, beta, stream or torch.cuda.current_stream().cuda_stream)
        if not gather:
            return y
        rows = int(self.comm.bounds(idx)[-1])
        return self.comm.allgather_rows(idx, y, torch.empty(rows, device=self.device))

    def close(self):
        self.engine.close()


def reassemble(blocks: Sequence[np.ndarray]) -> np.ndarray:
    """Host-side concatenation of per-rank y blocks (tests / debugging)."""
    return np.concatenate(list(blocks)) if len(blocks) else np.zeros(0, np.float32)
