"""Host-side clustering heuristic: what stays on the host in front of the B200 path.

``ClusteringHeuristic::compute_optimal_k`` (src/clustering.rs:36-72) mixes three things:
  * the Two-NN distance scan (:118-145)           -> on the GPU (asb_twonn_distances, K1);
  * closed-form bounds (step1_bounds, :75-98)      -> restated here exactly;
  * a Calinski-Harabasz search over smartcore KMeans and a pilot-k-means radius
    (:167-492), seeded through rand::StdRng        -> third-party, NOT reproducible outside Rust.
In a Rust deployment the shim keeps calling the reference's own ``compute_optimal_k`` and hands
``(k_opt, radius)`` to the C ABI.  Without Rust, :func:`compute_optimal_k_standin` provides the
documented stand-in of SURVEY 8d: K = step1_bounds' k_max, radius = 1.5 x p90 of the squared
distance from <=1000 seeded sample rows to the nearest of K seeded pilot rows.  It only produces
INPUTS of the parity path (both the CUDA path and the oracle receive the same K and radius).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

CLUSTERING_SEED = 128  # src/clustering.rs:30


def intrinsic_dim_from_distances(n: int, f: int, d1: np.ndarray, d2: np.ndarray) -> int:
    """src/clustering.rs:108-110,136-163."""
    if n < 10:
        return min(f, 2)
    ok = d1 > 1e-12 if n - 1 >= 2 else np.zeros_like(d1, dtype=bool)
    if not ok.any():
        return min(f, 3)
    ratios = d2[ok] / d1[ok]
    mean_ratio = float(np.sum(ratios) / len(ratios))
    idv = 1.0 / math.log(mean_ratio) if mean_ratio > 1.001 else float(f)
    r = math.floor(abs(idv) + 0.5) * (1 if idv >= 0 else -1)  # f64::round: half away from zero
    return int(min(max(r, 1), f))


def step1_bounds(n: int, f: int, id_est: int) -> Tuple[int, int]:
    """src/clustering.rs:85-97."""
    k_min = max(int(math.ceil(math.sqrt(n / 10.0))), 2)
    cands = [f, n // 10, 5 * id_est, int(math.pow(float(n), 0.5))]
    k_max = min(max(min(cands), k_min + 1), n // 2)
    return k_min, k_max


def _take_rows(rows, idx: np.ndarray) -> np.ndarray:
    if hasattr(rows, "is_cuda"):
        import torch
        return rows[torch.as_tensor(idx, device=rows.device)].cpu().numpy()
    return np.asarray(rows)[idx]


def sample_indices(n: int, size: int, seed: int) -> np.ndarray:
    """Stand-in for StdRng::seed_from_u64(seed) + shuffle (src/clustering.rs:54-57,113-116)."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    return rs.permutation(n)[: min(size, n)].astype(np.int64)


def pilot_radius(rows, k: int, seed: int) -> float:
    """1.5 x p90 of d^2(sample row, nearest pilot row) -- stand-in for compute_threshold_from_pilot
    (src/clustering.rs:384-492; the 1.5 x p90 rule is the reference's, the pilot centres are not)."""
    n = int(rows.shape[0])
    samp = _take_rows(rows, sample_indices(n, 1000, seed + 7))
    pil = _take_rows(rows, sample_indices(n, k, seed + 11))
    d2 = (samp * samp).sum(1)[:, None] + (pil * pil).sum(1)[None, :] - 2.0 * samp @ pil.T
    d2 = np.maximum(d2, 0.0)
    d2[d2 < 1e-18] = np.inf  # a sample row that is itself a pilot row
    nearest = d2.min(1)
    nearest = nearest[np.isfinite(nearest)]
    if len(nearest) == 0:
        return 1.0
    return float(1.5 * np.percentile(nearest, 90))


def compute_optimal_k_standin(ctx, rows, seed: Optional[int] = None) -> Tuple[int, float, int]:
    """(k_opt, radius, id_est) -- see the module docstring."""
    base_seed = CLUSTERING_SEED if seed is None else int(seed)
    n, f = int(rows.shape[0]), int(rows.shape[1])
    if n < 10:
        id_est = min(f, 2)
    else:
        si = sample_indices(n, 500, base_seed + 1)
        d1, d2 = ctx.twonn_distances(rows, si)
        id_est = intrinsic_dim_from_distances(n, f, d1, d2)
    _, k_max = step1_bounds(n, f, id_est)
    k_opt = max(k_max, 1)
    radius = pilot_radius(rows, k_opt, base_seed)
    return k_opt, radius, id_est
