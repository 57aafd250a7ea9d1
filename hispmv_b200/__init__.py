"""hispmv_b200 -- B200 (sm_100a) SpMV / GeMV engine behind HiSpMV's plugin surface.

Layout:
  csrc/            hand-written CUDA (spmv.cu, gemv.cu, partition.cu, synth.cu), the C-ABI (capi.cu) and the
                   pybind11 shim (pyhispmv_bindings.cpp)
  capi.py          ctypes view of include/hispmv.h
  engine.py        host-side mirror of the reference's FpgaHandle, plus device-tensor entry points
  sharded.py       one-process-per-GPU row-block sharding over torch.distributed (NCCL)
  layers.py        FpgaLayerManager / FpgaLinear equivalents for the apps/ MLP path
  synth.py         BASELINE.json's synthetic workloads
There is no CPU implementation in this package and no fallback path.
"""
from . import capi  # noqa: F401  (fails loudly when libhispmv_cuda.so is not built)
from .engine import Engine, shard_bounds  # noqa: F401

__all__ = ["Engine", "shard_bounds", "capi"]
