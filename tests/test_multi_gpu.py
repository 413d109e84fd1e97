"""2-GPU NCCL test of the row-sharded path (skipped when fewer than 2 GPUs are visible): the
sharded build + search must equal the single-GPU result on the concatenated rows bit for bit
(clustering pipeline hand-off, replicated Laplacian, lambda all-reduce, top-k all-gather + merge)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, f, maxk, radius, nq, k, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import arrowspace_b200 as asb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = asb.Context(rank, stream=torch.cuda.current_stream().cuda_stream or None)
        comp = asb.parallel.GpuCompute(ctx)
        lo, hi = asb.parallel.shard_bounds(n, rank, world)
        rows = torch.from_numpy(asb.synth.protein_like(hi - lo, f, seed=42, row0=lo)).to(dev)
        gp = asb.GraphParams(0.5, 12, 4, 2.0, 0.25)
        index = asb.parallel.build_sharded(comp, dist, rows, lo, n, gp, asb.TauMode.Median, maxk, radius,
                                           comm_device=dev, rank=rank, world=world)
        q = asb.synth.rows_at(asb.synth.query_indices(n, nq, 43), f, 42) * 1.02
        idx, score, count = asb.parallel.search_sharded(comp, dist, index, torch.from_numpy(q).to(dev), k, 0.7,
                                                        comm_device=dev, world=world)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx.cpu().numpy(), score=score.cpu().numpy(),
                 count=count.cpu().numpy(), lam=index.lambdas.cpu().numpy(), cent=index.centroids,
                 asg=index.assignments.cpu().numpy(), stats=np.array(index.lambda_stats),
                 indptr=index.csr[0], indices=index.csr[1], data=index.csr[2])
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_equals_single_gpu(tmp_path, asb):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n, f, maxk, nq, k, world = 60_001, 128, 100, 257, 10, 2
    radius = 1.5 * f * 0.0025 * 2
    mp.spawn(_worker, args=(world, _free_port(), n, f, maxk, radius, nq, k, str(tmp_path)), nprocs=world, join=True)
    ctx = asb.Context(0)
    x = asb.synth.protein_like(n, f, seed=42)
    b = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_synthesis(asb.TauMode.Median)
         .with_seed(42).with_inline_sampling(None).with_dims_reduction(False, None).with_cluster_params(maxk, radius))
    aspace, gl = b.build(x)
    q = asb.synth.rows_at(asb.synth.query_indices(n, nq, 43), f, 42) * 1.02
    widx, wscore, wcount, _ = aspace.search_batch(q, k, 0.7)
    for r in range(world):
        p = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = asb.parallel.shard_bounds(n, r, world)
        assert np.array_equal(p["cent"].view(np.uint64), gl.init_data.view(np.uint64))
        assert np.array_equal(p["asg"], aspace.cluster_assignments[lo:hi])
        assert np.array_equal(p["indptr"], gl.indptr) and np.array_equal(p["indices"], gl.indices)
        assert np.array_equal(p["data"], gl.data)
        assert np.array_equal(p["lam"], aspace.lambdas[lo:hi])
        lam = aspace.lambdas
        assert np.allclose(p["stats"], [lam.min(), lam.max(), lam.mean()], rtol=1e-12)
        assert np.array_equal(p["idx"], widx) and np.array_equal(p["count"], wcount)
        assert np.allclose(p["score"], wscore, rtol=0, atol=1e-14)


# ---- the same through the C ABI's own NCCL entry points (asb_comm_*, asb_*_sharded): no torch.distributed collective
# on the data path -- torch only ships the 128-byte ncclUniqueId between the processes
def _worker_abi(rank, world, port, n, f, maxk, radius, nq, k, out_dir, oracle_check):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import arrowspace_b200 as asb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)        # CPU rendezvous only
    try:
        ctx = asb.Context(rank)
        box = [asb.host.Comm.make_unique_id(ctx) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = asb.host.Comm(ctx, world, rank, box[0])
        lo, hi = asb.parallel.shard_bounds(n, rank, world)
        rows = torch.from_numpy(asb.synth.protein_like(hi - lo, f, seed=42, row0=lo)).to(dev)
        sample = asb.heuristics.sample_indices(n, 300, 129)
        d1, d2 = ctx.twonn_distances_sharded(comm, rows, lo, sample)
        gp = asb.GraphParams(0.5, 12, 4, 2.0, 0.25)
        bp = asb.host.BuildParamsC(gp.to_c(), asb.TauMode.Median.mode, asb.TauMode.Median.value, maxk, radius, 0)
        index = asb.host.ShardedIndex(ctx, comm, rows, lo, n, bp)
        spec, fb = ctx.kernel_ms("cluster_shard_speculative"), ctx.kernel_ms("cluster_shard_fallback")
        q = asb.synth.rows_at(asb.synth.query_indices(n, nq, 43), f, 42) * 1.02
        idx, score, count, lq = index.search(torch.from_numpy(q).to(dev), k, 0.7)
        torch.cuda.synchronize()
        info = index.info()
        ip, ii, dd = index.laplacian()
        np.savez(os.path.join(out_dir, f"abi_rank{rank}.npz"), idx=idx.cpu().numpy(), score=score.cpu().numpy(),
                 count=count.cpu().numpy(), lam=index.lambdas(), cent=index.centroids(), asg=index.assignments(),
                 sizes=index.cluster_sizes(), stats=np.array([info.lambda_min, info.lambda_max, info.lambda_sum]),
                 indptr=ip, indices=ii, data=dd, d1=d1, d2=d2, spec=np.array([spec, fb]))
        index.close()
        comm.close()
    finally:
        dist.destroy_process_group()


def _check_abi_outputs(asb, tmp_path, world, n, f, maxk, radius, nq, k, expect_speculative):
    ctx = asb.Context(0)
    x = asb.synth.protein_like(n, f, seed=42)
    b = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_synthesis(asb.TauMode.Median)
         .with_seed(42).with_inline_sampling(None).with_dims_reduction(False, None).with_cluster_params(maxk, radius))
    aspace, gl = b.build(x)
    q = asb.synth.rows_at(asb.synth.query_indices(n, nq, 43), f, 42) * 1.02
    widx, wscore, wcount, _ = aspace.search_batch(q, k, 0.7)
    wd1, wd2 = ctx.twonn_distances(x, asb.heuristics.sample_indices(n, 300, 129))
    lam = aspace.lambdas
    spec_seen = 0
    for r in range(world):
        p = np.load(tmp_path / f"abi_rank{r}.npz")
        lo, hi = asb.parallel.shard_bounds(n, r, world)
        assert np.array_equal(p["cent"].view(np.uint64), gl.init_data.view(np.uint64)), "centroids must be bit-identical"
        assert np.array_equal(p["sizes"], np.asarray(aspace.cluster_sizes).astype(np.uint64))
        assert np.array_equal(p["asg"], aspace.cluster_assignments[lo:hi])
        assert np.array_equal(p["indptr"], gl.indptr) and np.array_equal(p["indices"], gl.indices)
        assert np.array_equal(p["data"], gl.data)
        assert np.array_equal(p["lam"], lam[lo:hi])
        assert np.allclose(p["stats"], [lam.min(), lam.max(), lam.sum()], rtol=1e-12)
        assert np.array_equal(p["idx"], widx) and np.array_equal(p["count"], wcount)
        assert np.allclose(p["score"], wscore, rtol=0, atol=1e-14)
        assert np.allclose(p["d1"], wd1, rtol=1e-9) and np.allclose(p["d2"], wd2, rtol=1e-9)    # K1 cross-shard merge
        spec_seen += int(p["spec"][0])
    if expect_speculative:
        assert spec_seen >= 1, "no shard was proven against the common snapshot (every shard fell back)"


@pytest.mark.parametrize("n,f,maxk,rscale,expect_spec", [
    (600_000, 384, 384, None, True),     # the bench's shape and radius rule: pieces of shard 1 proven against the snapshot
    (600_000, 128, 100, 1.0, False),     # two centroids share some blobs: rows on their bisector, pieces fall back
    (60_001, 128, 100, 1.0, False),      # snapshot = whole shard 0
    (40_000, 64, 500, 0.2, False)])      # never saturates: no speculation at all
def test_two_gpu_c_abi_sharded_equals_single_gpu(tmp_path, asb, n, f, maxk, rscale, expect_spec):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    nq, k, world = 257, 10, 2
    if rscale is None:
        radius = asb.heuristics.pilot_radius(asb.synth.protein_like(50_000, f, seed=42), maxk, asb.heuristics.CLUSTERING_SEED)
    else:
        radius = rscale * 1.5 * f * 0.0025 * 2
    mp.spawn(_worker_abi, args=(world, _free_port(), n, f, maxk, radius, nq, k, str(tmp_path), False), nprocs=world, join=True)
    _check_abi_outputs(asb, tmp_path, world, n, f, maxk, radius, nq, k, expect_spec)
