"""Row-sharded multi-GPU driver for the path: one process per GPU, items and their
lambdas sharded by row, centroids + feature Laplacian + queries replicated.

Exchanges (torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests):
  * clustering is order dependent, so the K x F centroid state travels down the ranks
    (rank g resumes the walk from rank g-1's state: <= 6.1 MB point-to-point), then the final
    state is broadcast -- identical to one walk over the concatenated rows;
  * lambda statistics: all_reduce(min, max, sum) of 3 doubles (src/eigenmaps.rs:372-382);
  * search: all_gather of the per-shard top-k (Q x k x 16 B per rank) + k-way merge by
    (score desc, global index asc) = the order of the reference's stable sort (src/core.rs:785).
No collective touches the N x F items.

The per-shard arithmetic is injected (``compute``): :class:`GpuCompute` drives the C ABI; the CPU
gloo tests inject an oracle-backed object to exercise exactly this host logic without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Block sharding: rank g owns items [g*ceil(n/P), min(n, (g+1)*ceil(n/P)))."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class GpuCompute:
    """Per-shard stages on the local B200 through the C ABI (device tensors in, device tensors out)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def cluster_resume(self, rows, maxk, radius, cent, sizes, x):
        return self.ctx.cluster_incremental_resume(rows, maxk, radius, cent, sizes, x)

    def laplacian(self, cent_valid, gp):
        return self.ctx.build_feature_laplacian(cent_valid, gp)

    def taumode(self, rows, csr, taumode):
        lam, n2, stats = self.ctx.compute_taumode(rows, csr, taumode, want_norms=True)
        return lam, n2, stats

    def query_lambdas(self, queries, csr, taumode):
        return self.ctx.prepare_query_lambdas(queries, csr, taumode)

    def search(self, rows, lam, n2, queries, lq, k, alpha, offset):
        return self.ctx.search_lambda_aware_batch(rows, lam, queries, lq, k, alpha, norms2=n2, index_offset=offset)

    def merge(self, scores, idx, parts, nq, k):
        return self.ctx.topk_merge(scores, idx, parts, nq, k)


@dataclass
class ShardedIndex:
    rows: object            # local shard (numpy or torch tensor)
    offset: int             # global index of the first local row
    n_global: int
    lambdas: object
    norms2: object
    centroids: np.ndarray   # replicated, x * f
    csr: tuple              # replicated feature Laplacian
    lambda_stats: Tuple[float, float, float]   # global (min, max, mean)
    assignments: object
    taumode: object
    timings: dict


def _to_host(a) -> np.ndarray:
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def build_sharded(compute, dist, rows_local, offset: int, n_global: int, gp, taumode, max_clusters: int,
                  radius: float, comm_device="cpu", rank: Optional[int] = None, world: Optional[int] = None):
    """Stages 1-3 of ArrowSpaceBuilder::build (src/builder.rs:249-455) over a row-sharded dataset."""
    import torch

    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    f = int(rows_local.shape[1])
    cent = torch.zeros((max_clusters, f), dtype=torch.float64, device=comm_device)
    sizes = torch.zeros(max_clusters, dtype=torch.int64, device=comm_device)   # u64 bit pattern
    xt = torch.zeros(1, dtype=torch.int64, device=comm_device)
    # ---- stage 1: order-preserving pipeline (rank g resumes from rank g-1)
    if world > 1 and rank > 0:
        dist.recv(xt, src=rank - 1)
        dist.recv(cent, src=rank - 1)
        dist.recv(sizes, src=rank - 1)
    use_np = comm_device == "cpu" or str(comm_device) == "cpu"
    cbuf = cent.numpy() if use_np else cent
    sbuf = sizes.numpy().view(np.uint64) if use_np else sizes
    x_new, assign = compute.cluster_resume(rows_local, max_clusters, radius, cbuf, sbuf, int(xt.item()))
    xt.fill_(int(x_new))
    if world > 1:
        if rank < world - 1:
            dist.send(xt, dst=rank + 1)
            dist.send(cent, dst=rank + 1)
            dist.send(sizes, dst=rank + 1)
        dist.broadcast(xt, src=world - 1)
        dist.broadcast(cent, src=world - 1)
        dist.broadcast(sizes, src=world - 1)
    x = int(xt.item())
    cent_valid = _to_host(cent)[:x].copy()
    # ---- stage 2: replicated feature Laplacian (identical input -> identical CSR on every rank)
    csr = compute.laplacian(cent_valid, gp)
    # ---- stage 3: local taumode + global lambda statistics
    lam, n2, stats = compute.taumode(rows_local, csr, taumode)
    st = torch.tensor([float(stats[0]), -float(stats[1]), float(stats[2])], dtype=torch.float64, device=comm_device)
    if world > 1:
        mn = st[0:1].clone()
        mx = st[1:2].clone()
        sm = st[2:3].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MIN)   # min of the negated maxima
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        st = torch.cat([mn, mx, sm])
    lam_stats = (float(st[0]), -float(st[1]), float(st[2]) / float(n_global))
    return ShardedIndex(rows_local, offset, n_global, lam, n2, cent_valid, csr, lam_stats, assign, taumode, {})


def search_sharded(compute, dist, index: ShardedIndex, queries, k: int, alpha: float, comm_device="cpu",
                   world: Optional[int] = None):
    """Batched EigenMaps::search (src/eigenmaps.rs:410-455) over the shards; every rank returns the
    merged (idx, score, count)."""
    import torch

    world = dist.get_world_size() if world is None else world
    nq = int(queries.shape[0])
    lq = compute.query_lambdas(queries, index.csr, index.taumode)
    idx, score, count = compute.search(index.rows, index.lambdas, index.norms2, queries, lq, k, alpha, index.offset)
    if world == 1:
        return idx, score, count
    ti = torch.as_tensor(_to_host(idx) if comm_device == "cpu" else idx).to(comm_device).contiguous().clone()
    ts = torch.as_tensor(_to_host(score) if comm_device == "cpu" else score).to(comm_device).contiguous().clone()
    tc = torch.as_tensor(_to_host(count) if comm_device == "cpu" else count).to(comm_device).contiguous()
    # mark unused tail slots (local shard smaller than k)
    col = torch.arange(ti.shape[1], device=ti.device)[None, :]
    ti[col >= tc[:, None]] = -1
    gi = torch.empty((world,) + tuple(ti.shape), dtype=ti.dtype, device=comm_device)
    gs = torch.empty((world,) + tuple(ts.shape), dtype=ts.dtype, device=comm_device)
    dist.all_gather_into_tensor(gi, ti) if hasattr(dist, "all_gather_into_tensor") and comm_device != "cpu" else \
        dist.all_gather(list(gi.unbind(0)), ti)
    dist.all_gather_into_tensor(gs, ts) if hasattr(dist, "all_gather_into_tensor") and comm_device != "cpu" else \
        dist.all_gather(list(gs.unbind(0)), ts)
    if comm_device == "cpu":
        ms, mi, mc = compute.merge(gs.numpy(), gi.numpy(), world, nq, k)
    else:
        ms, mi, mc = compute.merge(gs, gi, world, nq, k)
    return mi, ms, mc
