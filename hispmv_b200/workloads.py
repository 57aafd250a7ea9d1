"""BASELINE.json's synthetic workloads as plain data: shapes, generator parameters and the reference's closed-form
test vectors.  No compute and no import of the CUDA library, so bench.py's reference arm (and anything else that
only needs the numbers) can load this file on its own:

    importlib.util.spec_from_file_location("workloads", ".../hispmv_b200/workloads.py")

The generators themselves live in csrc/synth.cu (device) and oracle/oracle.c (restatement); synth.py wraps the
device side.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

SYNTH_POWERLAW, SYNTH_UNIFORM, SYNTH_STENCIL27 = 1, 2, 3   # include/hispmv_synth.h


@dataclass(frozen=True)
class SynthSpec:
    name: str
    kind: int
    seed: int
    rows: int
    cols: int
    params: Tuple[int, int, int]

    def params_array(self) -> np.ndarray:
        return np.asarray(self.params, dtype=np.int64)


# P(len >= L) = (K / 2^32) / L.  K/2^32 = 0.6912 gives 10.0 nnz/row on average (100.0M nnz in 10M rows) when clipped at 1M (alpha = 2 tail);
# 5 rows in 10M sit at the 1M clip, 31% of the rows are empty.
_K_C2 = int(round(0.6912 * 2 ** 32))


def c2_powerlaw(scale: float = 1.0) -> SynthSpec:
    """C2: 10M x 10M, ~100M nnz, power-law rows (clip 1M) and power-law columns (gamma 5 ~ Zipf s=0.8)."""
    n = max(1024, int(10_000_000 * scale))
    clip = max(64, int(1_000_000 * min(1.0, scale * 4)))
    return SynthSpec("C2_powerlaw", SYNTH_POWERLAW, 1, n, n, (_K_C2, clip, 5))


def c2_weak(world: int, scale: float = 1.0) -> SynthSpec:
    """The weak-scaling workload at N GPUs: N stacked blocks of C2's shape.  Block k has C2's row-length sequence
    (params[2] carries the period, include/hispmv_synth.h) and its own entries, so every block -- and with nnz-balanced
    row blocks every GPU -- holds exactly C2's 100.0 M nonzeros.  world = 1 is C2 itself."""
    base = c2_powerlaw(scale)
    if world <= 1:
        return base
    k, clip, gamma = base.params
    return SynthSpec(base.name, base.kind, base.seed, base.rows * world, base.cols, (k, clip, gamma | (base.rows << 8)))


def c4_stencil(scale: float = 1.0) -> SynthSpec:
    """C4: 27-point stencil on a 272^3 grid (20.1M rows, ~540M nnz), banded / FEM-like."""
    g = max(4, int(round(272 * scale ** (1.0 / 3.0))))
    n = g * g * g
    return SynthSpec("C4_stencil27", SYNTH_STENCIL27, 2, n, n, (g, g, g))


def c5_uniform(scale: float = 1.0) -> SynthSpec:
    """C5: 100M x 100M, ~1B nnz, 6 + popcount(8 random bits) nnz per row (mean 10), uniform columns."""
    n = max(1024, int(100_000_000 * scale))
    return SynthSpec("C5_uniform", SYNTH_UNIFORM, 3, n, n, (6, 0xFF, 0))


def reference_vectors(rows: int, cols: int):
    """The reference's closed-form test vectors (cpu/src/main.cpp:173-178): x_j=(j+1)/(j+2), y0_i=-2(i+1)/(i+2)."""
    j = np.arange(cols, dtype=np.int64)
    i = np.arange(rows, dtype=np.int64)
    x = (j + 1).astype(np.float32) / (j + 2).astype(np.float32)          # float(j+1)/float(j+2)
    y0 = (np.float32(-2.0) * (i + 1).astype(np.float32)) / (i + 2).astype(np.float32)  # -2.0f*(i+1)/float(i+2)
    return x.astype(np.float32), y0.astype(np.float32)


REF_ALPHA, REF_BETA = np.float32(0.85), np.float32(-2.06)  # cpu/src/main.cpp:147-148
