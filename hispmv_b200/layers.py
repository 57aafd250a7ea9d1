"""The DNN-layer callers of the hot path: equivalents of the reference's apps/ helpers.

  LayerManager.process_weights / replace_layers   apps/fpga_layer_manager.py:15-52, 54-80
  FpgaLinear                                      apps/fpga_layer_manager.py:58-67  (numpy in, fpga.linear, numpy out)
  SparseLinear, ThreeLayerFCModel(Config)         apps/model.py:10-44, 47-80        (the model_test MLP)
  DeviceChain                                     SURVEY 8(f2): the same layers with activations kept in HBM,
                                                  bias + ReLU fused into the kernel epilogue, optional CUDA graph

`fpga` below is anything with the reference's FpgaHandle surface: the compiled `pyhispmv.FpgaHandle` or
`hispmv_b200.Engine`.  torch is used for modules, device memory and streams only; every matrix-vector product
goes through the C-ABI.  The CPU reference model (`ThreeLayerFCModel.forward`) exists for the same reason it
exists in the reference -- to be compared against -- and uses torch's CPU kernels, not this package.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn


class SparseLinear(nn.Module):
    """apps/model.py:10-44.  weight = randn * (rand < density) as a COO tensor, zero bias.  The reference's forward
    calls sparse_dot_mkl (absent here); torch.sparse.mm on the CPU is the stand-in for the comparison model."""

    def __init__(self, input_size: int, output_size: int, density: float):
        super().__init__()
        self.input_size, self.output_size, self.density = input_size, output_size, density
        dense_weight = torch.randn((output_size, input_size))
        mask = torch.rand_like(dense_weight) < density
        self.weight = (dense_weight * mask).to_sparse()
        self.bias = nn.Parameter(torch.zeros(output_size))

    def forward(self, x):
        return torch.sparse.mm(self.weight, x.T).T + self.bias


class ThreeLayerFCModelConfig:
    """apps/model.py:47-54."""

    def __init__(self, input_size=4096, dense_size=8192, sparse_size1=8192, sparse_size2=1024, density1=0.1,
                 density2=0.25):
        self.input_size, self.dense_size = input_size, dense_size
        self.sparse_size1, self.sparse_size2 = sparse_size1, sparse_size2
        self.density1, self.density2 = density1, density2


class ThreeLayerFCModel(nn.Module):
    """apps/model.py:57-80 without the rp_time timing loops: dense -> ReLU -> sparse -> ReLU -> sparse -> ReLU."""

    def __init__(self, config: Optional[ThreeLayerFCModelConfig] = None):
        super().__init__()
        self.config = config or ThreeLayerFCModelConfig()
        c = self.config
        self.dense = nn.Linear(c.input_size, c.dense_size)
        self.sparse1 = SparseLinear(c.dense_size, c.sparse_size1, c.density1)
        self.sparse2 = SparseLinear(c.sparse_size1, c.sparse_size2, c.density2)
        self.activation = nn.ReLU()

    def forward(self, x):
        x = self.activation(self.dense(x))
        x = self.activation(self.sparse1(x))
        return self.activation(self.sparse2(x))


def compare_model_outputs(ref_output: torch.Tensor, cmp_output: torch.Tensor) -> None:
    """The report model_test prints after the two forward passes (apps/model.py:82-142): absolute and relative error
    (|ref - cmp| / (ref + 1e-8), signed denominator as there) histograms over ten shared bins from 0 to the larger of the
    two maxima, then the maxima with the values at their positions."""
    ref, cmp_ = ref_output.flatten(), cmp_output.flatten()
    abs_err = torch.abs(ref - cmp_)
    rel_err = torch.abs((ref - cmp_) / (ref + 1e-8))
    a, r = abs_err.cpu().numpy(), rel_err.cpu().numpy()
    bins = np.linspace(0, max(np.max(a), np.max(r)), 11)
    for title, err in (("Absolute", a), ("Relative", r)):
        hist, edges = np.histogram(err, bins=bins)
        print(f"\n{title} Error Histogram:")
        for i in range(len(hist)):
            print(f"Range: ({edges[i]:.4f}, {edges[i + 1]:.4f}), Count: {hist[i]}")
    ia, ir = torch.argmax(abs_err), torch.argmax(rel_err)
    print("\n")
    print(f"Max Absolute Error: {torch.max(abs_err).item()}")
    print(f"Max Relative Error: {torch.max(rel_err).item()}")
    print(f"Values at Max Absolute Error (index {ia.item()}):")
    print(f"Ref Output: {ref[ia].item()}, Actual Output: {cmp_[ia].item()}")
    print(f"Values at Max Relative Error (index {ir.item()}):")
    print(f"Ref Output: {ref[ir].item()}, Actual Output: {cmp_[ir].item()}")


def layer_arrays(layer) -> Tuple[str, tuple, np.ndarray]:
    """What process_weights extracts from a layer (apps/fpga_layer_manager.py:15-52):
    ("sparse", (rows_i32, cols_i32, vals_f32, out, in), bias) or ("dense", (flat_f32, out, in), bias).
    nn.Linear / Conv1D weights go dense when more than half of the entries are nonzero, COO otherwise."""
    if not hasattr(layer, "weight"):
        raise ValueError("Layer must have a weight attribute.")
    if isinstance(layer, SparseLinear) or (isinstance(layer.weight, torch.Tensor) and layer.weight.is_sparse):
        w = layer.weight.coalesce() if not layer.weight.is_coalesced() else layer.weight
        idx, val = w.indices(), w.values()
        out_f, in_f = w.shape
        kind, arrays = "sparse", (idx[0].detach().numpy().astype(np.int32), idx[1].detach().numpy().astype(np.int32),
                                  val.detach().numpy().astype(np.float32), out_f, in_f)
    else:
        conv1d = type(layer).__name__ == "Conv1D"  # transformers.Conv1D stores (in, out)
        weight = (layer.weight.t() if conv1d else layer.weight).detach().numpy()
        out_f, in_f = weight.shape
        if np.count_nonzero(weight) / weight.size > 0.5:
            kind, arrays = "dense", (np.ascontiguousarray(weight, np.float32).reshape(-1), out_f, in_f)
        else:
            r, c = np.nonzero(weight)
            kind, arrays = "sparse", (r.astype(np.int32), c.astype(np.int32), weight[r, c].astype(np.float32), out_f, in_f)
    bias = getattr(layer, "bias", None)
    bias = bias.detach().numpy().astype(np.float32) if bias is not None else np.zeros(out_f, np.float32)
    return kind, arrays, bias


class FpgaLinear(nn.Module):
    """apps/fpga_layer_manager.py:58-67: forward = fpga.linear(idx, x.view(-1).numpy(), bias), reshaped."""

    def __init__(self, fpga, matrix_idx: int, bias: np.ndarray):
        super().__init__()
        self.fpga, self.matrix_idx, self.bias_npy = fpga, matrix_idx, bias

    def forward(self, x):
        y = self.fpga.linear(self.matrix_idx, x.reshape(-1).numpy(), self.bias_npy)
        return torch.from_numpy(np.asarray(y)).view(*x.shape[:-1], self.bias_npy.shape[0])


class LayerManager:
    """FpgaLayerManager (apps/fpga_layer_manager.py:8-80)."""

    def process_weights(self, layer, fpga) -> Tuple[int, np.ndarray]:
        kind, arrays, bias = layer_arrays(layer)
        idx = fpga.create_sparse_handle(*arrays) if kind == "sparse" else fpga.create_dense_handle(*arrays)
        if idx == -1:
            raise RuntimeError("FPGA memory is full.")  # the reference's message, fpga_layer_manager.py:49-50
        return idx, bias

    def replace_layers(self, model: nn.Module, fpga) -> nn.Module:
        """A copy of `model` whose Linear / Conv1D / SparseLinear children run through fpga.linear."""
        import copy
        new_model = copy.copy(model)
        new_model._modules = dict(model._modules)
        targets = [(n, m) for n, m in model.named_modules()
                   if isinstance(m, (nn.Linear, SparseLinear)) or type(m).__name__ == "Conv1D"]
        for name, module in targets:
            parts = name.split(".")
            parent = new_model
            for p in parts[:-1]:
                child = copy.copy(getattr(parent, p))
                child._modules = dict(child._modules)
                setattr(parent, p, child)
                parent = child
            idx, bias = self.process_weights(module, fpga)
            setattr(parent, parts[-1], FpgaLinear(fpga, idx, bias))
        fpga.load_matrices()
        return new_model


FpgaLayerManager = LayerManager  # the reference's class name


class DeviceChain:
    """Chained layers that never leave HBM: y_k = relu(A_k x_k + b_k) with bias and ReLU fused into the kernel
    epilogue (hispmv_linear_dev).  With `shards` (a hispmv_b200.sharded.RowBlocks per layer) every rank holds a row
    block of each layer and the blocks of y are all-gathered into the next layer's replicated x.  `graph=True`
    captures the launches of one forward pass in a CUDA graph and replays it.

    `fused=True` (sharded chains on NVSwitch boxes): the activations live in symmetric memory and every layer's kernel
    stores its y block straight to the multicast address of the next layer's x (Engine.run_dev_mc, multimem.st), so
    the all-gather disappears into the producing kernel; one device-side barrier per layer orders it against the
    readers.  Falls back to the NCCL all-gather when no multicast mapping can be made."""

    def __init__(self, engine, layers: Sequence[nn.Module], relu: Sequence[bool], comm=None, graph: bool = False,
                 fused: bool = False):
        self.engine, self.comm = engine, comm
        self.idx: List[int] = []
        self.bias: List[torch.Tensor] = []
        self.relu = list(relu)
        self.shapes: List[Tuple[int, int]] = []
        self.blocks = []
        mgr = LayerManager()
        for layer in layers:
            idx, bias = mgr.process_weights(layer, engine)
            info = engine.matrix_info(idx)
            rb, re = info["row_begin"], info["row_end"]
            self.idx.append(idx)
            self.bias.append(torch.from_numpy(bias[rb:re].copy()).cuda(engine.device_id))
            self.shapes.append((info["rows"], info["cols"]))
            self.blocks.append((rb, re))
            if comm is not None:
                comm.set_blocks(("chain", id(self), len(self.blocks) - 1), rb, re, torch.device("cuda", engine.device_id))
        engine.load_matrices()
        dev = torch.device("cuda", engine.device_id)
        sizes = [c for _, c in self.shapes] + [self.shapes[-1][0]]
        self.fused, self._hdl, self._mc = False, None, []
        if fused and comm is not None:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                import torch.distributed as dist
                offs = [0]
                for n in sizes:
                    offs.append(offs[-1] + ((n + 3) & ~3))          # 16-byte aligned activations in one allocation
                buf = symm_mem.empty(offs[-1], dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, group=comm.group if comm.group is not None else dist.group.WORLD)
                if not hdl.multicast_ptr:
                    raise RuntimeError("no multicast mapping (NVLS unavailable)")
                buf.zero_()
                self._x = [buf[offs[k]:offs[k] + n] for k, n in enumerate(sizes)]
                self._mc = [hdl.multicast_ptr + 4 * o for o in offs[:-1]]   # multicast view of activation k
                self._hdl, self._symm, self.fused = hdl, buf, True
            except Exception as ex:  # noqa: BLE001
                self.fused_unavailable = f"{type(ex).__name__}: {ex}"
        if not self.fused:
            self._x = [torch.empty(n, device=dev) for n in sizes]
        self._y_local = [torch.empty(re - rb, device=dev) for rb, re in self.blocks]
        self._graph = None
        self._want_graph = graph and comm is None

    def _launch(self, stream: int):
        for k, idx in enumerate(self.idx):
            rb, re = self.blocks[k]
            whole = (re - rb) == self.shapes[k][0]
            if self.fused:
                # y block -> every rank's copy of the next activation, by the kernel's own stores
                self.engine.run_dev_mc(idx, self._x[k], self.bias[k], self._mc[k + 1] + 4 * rb, 1.0, 1.0,
                                       relu=self.relu[k], stream=stream)
                self._hdl.barrier(channel=k)
                continue
            y = self._x[k + 1] if whole else self._y_local[k]
            self.engine.linear_dev(idx, self._x[k], self.bias[k], y, relu=self.relu[k], stream=stream)
            if not whole:
                self.comm.allgather_rows(("chain", id(self), k), y, self._x[k + 1])

    def forward_batch(self, x: torch.Tensor) -> torch.Tensor:
        """x: (batch, in_features) CUDA tensor -> (batch, out_features).  Every layer takes the whole batch in passes of
        up to eight vectors over its matrix (Engine.run_dev_batch: col/val or the dense rows are read once per pass),
        bias and ReLU fused, activations stay in HBM.  Single-GPU chains only (the sharded hand-over is per vector)."""
        if self.comm is not None:
            raise NotImplementedError("forward_batch is for unsharded chains")
        if x.dim() != 2 or x.shape[1] != self.shapes[0][1]:
            raise ValueError(f"expected (batch, {self.shapes[0][1]})")
        stream = torch.cuda.current_stream().cuda_stream
        act = x.contiguous().float()
        for k, idx in enumerate(self.idx):
            out = torch.empty((act.shape[0], self.shapes[k][0]), device=act.device, dtype=torch.float32)
            self.engine.run_dev_batch(idx, act, self.bias[k], out, 1.0, 1.0, relu=self.relu[k], stream=stream)
            act = out
        return act

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (in_features,) CUDA tensor.  Returns the (out_features,) result, replicated on every rank."""
        self._x[0].copy_(x.reshape(-1))
        stream = torch.cuda.current_stream().cuda_stream
        if self._want_graph:
            if self._graph is None:
                self._launch(stream)  # warm-up outside capture (lazy attribute setup)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(torch.cuda.current_stream().cuda_stream)
                self._graph = g
            self._graph.replay()
        else:
            self._launch(stream)
        return self._x[-1]
