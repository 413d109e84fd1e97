"""Seeded synthetic inputs shared by the GPU path, the oracle and bench.py.

A self-contained counter-based PRNG (splitmix64 finaliser over ``seed, stream, counter``) so the
CPU and GPU paths see bit-identical buffers on any machine -- no dependence on numpy's or Rust's
generators.  "Protein-like" = non-negative blobs, the only kind of data on which the reference's
feature graph is non-degenerate (SURVEY top-of-file facts).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _u64(seed: int, stream: int, start: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = _mix(np.array([np.uint64(seed) * _GOLD + np.uint64(stream)], dtype=np.uint64))[0]
        ctr = np.arange(start, start + count, dtype=np.uint64)
        return _mix(base + (ctr + np.uint64(1)) * _GOLD)


def uniform01(seed: int, stream: int, start: int, count: int) -> np.ndarray:
    """U[0,1) doubles with 53 random bits."""
    return (_u64(seed, stream, start, count) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal01(seed: int, stream: int, start: int, count: int) -> np.ndarray:
    """N(0,1) by Box-Muller on two independent streams."""
    u1 = uniform01(seed, 2 * stream + 100, start, count)
    u2 = uniform01(seed, 2 * stream + 101, start, count)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def protein_like(n: int, f: int, seed: int = 42, n_blobs: int = 64, noise: float = 0.05,
                 out: np.ndarray | None = None, chunk_rows: int = 65536, row0: int = 0) -> np.ndarray:
    """x = clip(c[g] + noise * N(0,1), 0, inf), c ~ U(0,1)^F per blob, g ~ U{0..n_blobs-1}.

    Rows are a pure function of (seed, global row index): ``row0`` selects the window
    [row0, row0+n) of the virtual dataset, so a row-sharded job generates its shard locally."""
    centres = uniform01(seed, 1, 0, n_blobs * f).reshape(n_blobs, f)
    if out is None:
        out = np.empty((n, f), dtype=np.float64)
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        g = (uniform01(seed, 2, row0 + r0, r1 - r0) * n_blobs).astype(np.int64)
        z = normal01(seed, 3, (row0 + r0) * f, (r1 - r0) * f).reshape(r1 - r0, f)
        np.maximum(centres[g] + noise * z, 0.0, out=out[r0:r1])
    return out


def rows_at(indices, f: int, seed: int = 42, n_blobs: int = 64, noise: float = 0.05) -> np.ndarray:
    """The rows of the virtual dataset at arbitrary global indices (same values protein_like gives)."""
    indices = np.asarray(indices, dtype=np.int64)
    centres = uniform01(seed, 1, 0, n_blobs * f).reshape(n_blobs, f)
    out = np.empty((len(indices), f), dtype=np.float64)
    for t, i in enumerate(indices):
        g = int(uniform01(seed, 2, int(i), 1)[0] * n_blobs)
        z = normal01(seed, 3, int(i) * f, f)
        np.maximum(centres[g] + noise * z, 0.0, out=out[t])
    return out


def query_indices(n: int, nq: int, seed: int = 43) -> np.ndarray:
    return (uniform01(seed, 7, 0, nq) * n).astype(np.int64)


def queries_from_items(items: np.ndarray, nq: int, seed: int = 43, scale: float = 1.02):
    """queries = random items x 1.02 (examples/01_compare_cosine.rs:86-90); returns (queries, item ids)."""
    n = items.shape[0]
    idx = query_indices(n, nq, seed)
    return np.ascontiguousarray(items[idx] * scale), idx
