"""CPU pin of the certified parallel replay of the clustering walk (csrc/cluster_replay.cu restated in numpy by
tests/replay_proto.py): whatever is proven must be the oracle's walk bit for bit, and settled data must be provable."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def _walk_with_replay(asb, oracle, x, kmax, radius, prefix, chunk):
    import replay_proto as rp
    C, a0, cnt = oracle.cluster_incremental(x[:prefix], kmax, radius)
    cnt = cnt.astype(np.float64)
    asg, proven, lo, n = [a0], [], prefix, len(x)
    while lo < n:
        hi = min(lo + chunk, n)
        ok, a, C2, cnt2, info = rp.single_sweep_chunk(x[lo:hi], C, cnt, radius, kmax)
        proven.append(ok)
        if ok:
            C, cnt = C2, cnt2
            asg.append(a)
        else:   # the sequential kernel's job: walk the chunk from the same state (here: the oracle up to hi)
            C, full, cntn = oracle.cluster_incremental(x[:hi], kmax, radius)
            cnt = cntn.astype(np.float64)
            asg.append(full[lo:hi])
        lo = hi
    return C, np.concatenate(asg), cnt, proven


def test_replay_is_the_walk_on_settled_data(asb, oracle):
    n, f = 28_000, 384
    x = asb.synth.protein_like(n, f, seed=42)
    kmax = 384
    radius = asb.heuristics.pilot_radius(x, kmax, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = oracle.cluster_incremental(x, kmax, radius)
    assert len(cent) == kmax                                    # saturated: no creations after the first rows
    C, a, cnt, proven = _walk_with_replay(asb, oracle, x, kmax, radius, 12_000, 8_000)
    assert all(proven) and len(proven) == 2
    assert np.array_equal(C.view(np.uint64), cent.view(np.uint64))
    assert np.array_equal(a, asg) and np.array_equal(cnt.astype(np.uint64), sizes)


def test_replay_refuses_what_it_cannot_prove(asb, oracle):
    rng = np.random.default_rng(5)
    x = rng.random((12_000, 64))                                # uniform noise: nothing settles
    cent, asg, sizes = oracle.cluster_incremental(x, 100, 9.0)
    C, a, cnt, proven = _walk_with_replay(asb, oracle, x, 100, 9.0, 4_000, 4_000)
    assert not any(proven)
    assert np.array_equal(C.view(np.uint64), cent.view(np.uint64)) and np.array_equal(a, asg)
    # a blob no centroid was opened for (K = 200 covers 63 of the 64 blobs): its rows are farther than sqrt(1.5 radius)
    # from everything, near-tied between many centroids -- and dropped whoever is nearest, which is provable
    x = asb.synth.protein_like(20_000, 384, seed=42)
    radius = asb.heuristics.pilot_radius(x, 200, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = oracle.cluster_incremental(x, 200, radius)
    assert (asg < 0).sum() > 100
    C, a, cnt, proven = _walk_with_replay(asb, oracle, x, 200, radius, 12_000, 8_000)
    assert all(proven)
    assert np.array_equal(C.view(np.uint64), cent.view(np.uint64)) and np.array_equal(a, asg)
    # unsaturated walk (K < max): a chunk is only provable while no row opens a new centroid
    x = asb.synth.protein_like(9_000, 128, seed=42)
    _, kmax = asb.heuristics.step1_bounds(9_000, 128, 128)
    radius = asb.heuristics.pilot_radius(x, kmax, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = oracle.cluster_incremental(x, kmax, radius)
    assert len(cent) < kmax
    C, a, cnt, proven = _walk_with_replay(asb, oracle, x, kmax, radius, 3_000, 3_000)
    assert np.array_equal(C.view(np.uint64), cent.view(np.uint64)) and np.array_equal(a, asg)
    assert np.array_equal(cnt.astype(np.uint64), sizes)
