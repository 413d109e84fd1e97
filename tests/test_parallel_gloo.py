"""world_size-2 gloo tests (CPU) of the row-sharded host logic: the centroid-state pipeline,
the lambda statistics all-reduce and the top-k all-gather + merge.  The per-shard arithmetic is
the ORACLE here (allowed in tests), so what is exercised is exactly arrowspace_b200.parallel."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


class OracleCompute:
    """Drop-in for parallel.GpuCompute backed by the CPU oracle."""

    def __init__(self):
        from oracle_binding import Oracle
        self.o = Oracle()

    def cluster_resume(self, rows, maxk, radius, cent, sizes, x):
        # the oracle has no resume entry: replay is not needed, emulate by continuing the same walk
        import ctypes as C
        rows = np.ascontiguousarray(rows)
        n, f = rows.shape
        asg = np.full(n, -1, dtype=np.int64)
        kc = int(x)
        for r in range(n):
            row = rows[r]
            if kc == 0:
                cent[0] = row; sizes[0] = 1; asg[r] = 0; kc = 1
                continue
            b, d2 = self.o.nearest_centroid(row, cent[:kc])
            if kc < maxk and d2 > radius * 0.5:
                cent[kc] = row; sizes[kc] = 1; asg[r] = kc; kc += 1
            elif d2 <= radius:
                k_new = float(sizes[b]) + 1.0
                cent[b] += (row - cent[b]) / k_new
                sizes[b] += 1; asg[r] = b
            elif d2 <= radius * 1.5:
                sizes[b] += 1; asg[r] = b
        return kc, asg

    def laplacian(self, cent, gp):
        return self.o.feature_laplacian(cent, eps=gp.eps, k=gp.k, topk=gp.topk, p=gp.p, sigma=gp.sigma)

    def taumode(self, rows, csr, tm):
        lam = self.o.compute_taumode(rows, csr, tm.mode, tm.value)
        return lam, (rows * rows).sum(1), np.array([lam.min(), lam.max(), lam.sum()])

    def query_lambdas(self, queries, csr, tm):
        return self.o.compute_taumode(queries, csr, tm.mode, tm.value)

    def search(self, rows, lam, n2, queries, lq, k, alpha, offset):
        idx, sc, cnt = self.o.search_lambda_aware_batch(rows, lam, queries, lq, k, alpha)
        idx = np.where(idx >= 0, idx + offset, idx)
        return idx, sc, cnt

    def merge(self, scores, idx, parts, nq, k):
        out_s = np.full((nq, k), -np.inf)
        out_i = np.full((nq, k), -1, dtype=np.int64)
        out_c = np.zeros(nq, dtype=np.int64)
        for q in range(nq):
            cand = [(-scores[p, q, e], idx[p, q, e]) for p in range(parts) for e in range(k) if idx[p, q, e] >= 0]
            cand.sort()
            cand = cand[:k]
            out_c[q] = len(cand)
            for r, (ns, i) in enumerate(cand):
                out_s[q, r] = -ns
                out_i[q, r] = i
        return out_s, out_i, out_c


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, f, maxk, radius, nq, k, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import arrowspace_b200 as asb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = asb.parallel.shard_bounds(n, rank, world)
        rows = asb.synth.protein_like(hi - lo, f, seed=42, row0=lo)
        gp = asb.GraphParams(0.5, 12, 4, 2.0, 0.25)
        comp = OracleCompute()
        index = asb.parallel.build_sharded(comp, dist, rows, lo, n, gp, asb.TauMode.Median, maxk, radius)
        full = asb.synth.protein_like(n, f, seed=42)
        queries, _ = asb.synth.queries_from_items(full, nq, seed=43)
        idx, score, count = asb.parallel.search_sharded(comp, dist, index, queries, k, 0.7)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx, score=score, count=count,
                 lam=index.lambdas, cent=index.centroids, asg=index.assignments, stats=np.array(index.lambda_stats),
                 indptr=index.csr[0], indices=index.csr[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(1500, 2), (1001, 3)])
def test_sharded_equals_single_process(tmp_path, oracle, asb, n, world):
    import torch.multiprocessing as mp
    f, maxk, nq, k = 32, 24, 9, 10
    radius = 1.5 * f * 0.0025 * 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, f, maxk, radius, nq, k, str(tmp_path)), nprocs=world, join=True)
    # single-process oracle on the concatenated rows
    x = asb.synth.protein_like(n, f, seed=42)
    cent, asg, sizes = oracle.cluster_incremental(x, maxk, radius)
    csr = oracle.feature_laplacian(cent, eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    lam = oracle.compute_taumode(x, csr, 1)
    queries, _ = asb.synth.queries_from_items(x, nq, seed=43)
    lq = oracle.compute_taumode(queries, csr, 1)
    widx, wscore, wcount = oracle.search_lambda_aware_batch(x, lam, queries, lq, k, 0.7)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for r, p in enumerate(parts):
        lo, hi = asb.parallel.shard_bounds(n, r, world)
        assert np.array_equal(p["cent"].view(np.uint64), cent.view(np.uint64))      # pipeline == one walk
        assert np.array_equal(p["asg"], asg[lo:hi])
        assert np.array_equal(p["indptr"], csr[0]) and np.array_equal(p["indices"], csr[1])
        assert np.array_equal(p["lam"], lam[lo:hi])
        assert np.allclose(p["stats"], [lam.min(), lam.max(), lam.mean()], rtol=1e-12)  # all-reduce
        assert np.array_equal(p["idx"], widx) and np.array_equal(p["count"], wcount)    # all-gather + merge
        assert np.allclose(p["score"], wscore, rtol=0, atol=1e-15)


def test_shard_bounds(asb):
    sb = asb.parallel.shard_bounds
    assert [sb(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [sb(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sb(1_000_000, 7, 8) == (875_000, 1_000_000)
