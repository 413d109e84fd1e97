"""K2 with certified parallel replay (option ``cluster_replay``, csrc/cluster_replay.cu) against the oracle's walk.

Whatever the replay proves or fails to prove, the outputs must be the walk's: centroids bit-identical, assignments
and sizes identical (src/clustering.rs:547-928, deterministic branch).  The algorithm is pinned on the CPU by
``tests/replay_proto.py`` (numpy restatement, bit-identical to the oracle).  The option is on by default since round 2
(first B200 run: profiles/r02_replay_first_run.json); the fixture still sets it explicitly and restores the defaults."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.fixture()
def rctx(ctx):
    ctx.set_option("cluster_replay", 1)
    try:
        yield ctx
    finally:
        for key, val in (("cluster_replay", 1), ("cluster_replay_prefix", 2048), ("cluster_replay_chunk", 1024),
                         ("cluster_replay_chunk_max", 262144), ("cluster_replay_generic_chain", 0),
                         ("cluster_replay_tf32", 0), ("twonn_prefilter", 1)):
            ctx.set_option(key, val)


def _same_walk(got, want):
    cent, asg, sizes = got
    wcent, wasg, wsizes = want
    assert np.asarray(cent).shape == wcent.shape
    assert np.array_equal(np.ascontiguousarray(cent).view(np.uint64), wcent.view(np.uint64)), "centroids differ"
    assert np.array_equal(np.asarray(asg), wasg), "assignments differ"
    assert np.array_equal(np.asarray(sizes).astype(np.uint64), wsizes), "sizes differ"


@pytest.mark.parametrize("n,f,prefix,chunk", [(60_000, 384, 16_384, 32_768), (40_000, 128, 4_096, 8_192),
                                              (30_000, 25, 2_048, 4_096), (20_001, 384, 1_024, 1_000),
                                              (12_000, 200, 2_048, 1_024), (9_000, 520, 2_048, 1_024)])
def test_replay_reproduces_the_walk(rctx, asb, oracle, n, f, prefix, chunk):
    x = asb.synth.protein_like(n, f, seed=42)
    _, kmax = asb.heuristics.step1_bounds(n, f, f)
    radius = asb.heuristics.pilot_radius(x[: min(n, 50_000)], kmax, asb.heuristics.CLUSTERING_SEED)
    want = oracle.cluster_incremental(x, kmax, radius)
    rctx.set_option("cluster_replay_prefix", prefix)
    rctx.set_option("cluster_replay_chunk", chunk)
    got = rctx.cluster_incremental(x, kmax, radius)
    _same_walk(got, want)
    assert rctx.kernel_ms("cluster_replay_chunks") >= 1
    rctx.set_option("cluster_replay_generic_chain", 1)      # the memory-resident chain kernel (any f) gives the same bits
    _same_walk(rctx.cluster_incremental(x, kmax, radius), want)


def test_replay_proves_the_settled_chunks_of_the_bench_data(rctx, asb, oracle):
    """C3-shaped data (64 blobs, K = 384 saturates in the first rows): every chunk after the prefix is proven."""
    n, f = 120_000, 384
    x = asb.synth.protein_like(n, f, seed=42)
    kmax, radius = 384, asb.heuristics.pilot_radius(x[:50_000], 384, asb.heuristics.CLUSTERING_SEED)
    want = oracle.cluster_incremental(x, kmax, radius)
    got = rctx.cluster_incremental(x, kmax, radius)
    _same_walk(got, want)
    tried, ok = rctx.kernel_ms("cluster_replay_chunks"), rctx.kernel_ms("cluster_replay_chunks_ok")
    assert tried == ok >= 7 and rctx.kernel_ms("cluster_replay_rows") == n - 2_048      # 1 k, 2 k, ... 64 k, the rest


@pytest.mark.parametrize("case", ["dup_at_1", "dup_at_5", "nan_row_3", "few_clusters", "all_creators", "tight"])
def test_growth_run_edge_cases(rctx, asb, oracle, case):
    """The creator run at the start of a fresh walk (growth_pairs_kernel): rows that each open a centroid are settled by
    one parallel pass over their pairwise distances; it must stop exactly where the walk stops opening centroids."""
    n, f, maxk = 3_000, 48, 60
    x = asb.synth.protein_like(n, f, seed=11)
    radius = 1.5 * f * 0.0025 * 2
    expect = None
    if case == "dup_at_1":
        x[1] = x[0]
        expect = 0              # a run of one row is not worth the hand-off
    elif case == "dup_at_5":
        x[5] = x[2]
        expect = 5
    elif case == "nan_row_3":
        x[3, 7] = np.nan        # NaN distances are never the nearest: the row opens a centroid, and so do its successors
    elif case == "few_clusters":
        maxk = 9
    elif case == "all_creators":
        radius = 1e-6
        maxk = 2_500            # the run is capped at 2048 rows
        expect = 2_048
    elif case == "tight":
        radius = 1e3            # everything falls into the first centroid
        expect = 0
    want = oracle.cluster_incremental(x, maxk, radius)
    rctx.set_option("cluster_replay_prefix", 512)
    got = rctx.cluster_incremental(x, maxk, radius)
    g = rctx.kernel_ms("cluster_growth_rows")
    if case != "nan_row_3":
        _same_walk(got, want)
    else:
        assert np.array_equal(np.asarray(got[1]), want[1]) and np.array_equal(np.asarray(got[2]).astype(np.uint64), want[2])
        m = ~np.isnan(want[0])
        assert np.array_equal(np.isnan(np.asarray(got[0])), np.isnan(want[0])) and np.array_equal(np.asarray(got[0])[m], want[0][m])
    if expect is not None:
        assert g == expect, (g, expect)
    # the run is a prefix of rows that are their own centroid
    assert np.array_equal(want[1][: int(g)], np.arange(int(g)))
    rctx.set_option("cluster_growth_run", 0)
    try:
        got0 = rctx.cluster_incremental(x, maxk, radius)
    finally:
        rctx.set_option("cluster_growth_run", 1)
    assert np.array_equal(np.asarray(got0[1]), np.asarray(got[1]))


def test_replay_gives_up_on_unsettled_data(rctx, asb, oracle):
    """Uniform noise never settles (every row is about as far from every centroid): chunks fail certification, the
    sequential kernel walks them -- twice as far after every failure in a row -- and the outputs are still the walk's."""
    rng = np.random.default_rng(5)
    n, f = 50_000, 64
    x = rng.random((n, f))
    kmax, radius = 100, 9.0          # |x - y|^2 ~ 10.7 for uniform rows: updates, soft assignments and drops all occur
    want = oracle.cluster_incremental(x, kmax, radius)
    rctx.set_option("cluster_replay_prefix", 8_192)
    rctx.set_option("cluster_replay_chunk", 8_192)
    got = rctx.cluster_incremental(x, kmax, radius)
    _same_walk(got, want)
    assert rctx.kernel_ms("cluster_replay_chunks_ok") == 0 and 1 <= rctx.kernel_ms("cluster_replay_chunks") <= 4


def test_replay_through_the_builder_and_resume(rctx, asb, oracle):
    """ArrowSpaceBuilder.build with the option set == without it; the resume entry (multi-GPU hand-off) too."""
    n, f, maxk = 70_000, 128, 100
    x = asb.synth.protein_like(n, f, seed=42)
    radius = 1.5 * f * 0.0025 * 2

    def build():
        return (asb.ArrowSpaceBuilder.new(rctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_seed(42)
                .with_inline_sampling(None).with_dims_reduction(False, None).with_cluster_params(maxk, radius).build(x))
    a1, g1 = build()
    rctx.set_option("cluster_replay", 0)
    a0, g0 = build()
    assert np.array_equal(g1.init_data.view(np.uint64), g0.init_data.view(np.uint64))
    assert np.array_equal(a1.cluster_assignments, a0.cluster_assignments)
    assert np.array_equal(a1.lambdas, a0.lambdas)
    rctx.set_option("cluster_replay", 1)
    want = oracle.cluster_incremental(x, maxk, radius)
    cent = np.zeros((maxk, f))
    sizes = np.zeros(maxk, dtype=np.uint64)
    k, parts = 0, []
    for a, b in [(0, 30_000), (30_000, n)]:
        k, part = rctx.cluster_incremental_resume(np.ascontiguousarray(x[a:b]), maxk, radius, cent, sizes, k)
        parts.append(part)
    _same_walk((cent[:k], np.concatenate(parts), sizes[:k]), want)


def test_replay_with_the_tf32_ranking(rctx, asb, oracle):
    """cluster_replay_tf32: nearest / runner-up of the snapshot from the certified TF32 ranking with direct-form
    distances (search_pf.cuh, PF_L2) instead of the FP64 tensor kernel -- same bits."""
    n, f = 60_000, 384
    x = asb.synth.protein_like(n, f, seed=42)
    kmax, radius = 384, asb.heuristics.pilot_radius(x[:50_000], 384, asb.heuristics.CLUSTERING_SEED)
    want = oracle.cluster_incremental(x, kmax, radius)
    rctx.set_option("cluster_replay_tf32", 1)
    _same_walk(rctx.cluster_incremental(x, kmax, radius), want)
    assert rctx.kernel_ms("cluster_replay_chunks_ok") >= 1 and rctx.kernel_ms("l2_pf_flags") == 0


@pytest.mark.parametrize("n,f", [(20_000, 128), (5_001, 384), (3_000, 25)])
def test_twonn_with_the_tf32_ranking(rctx, asb, oracle, n, f):
    """twonn_prefilter: the Two-NN scan (src/clustering.rs:118-145) ranked by the certified TF32 score, the two nearest
    distances evaluated in the reference's direct form -- bit-identical to the oracle (the FP64 kernel path is held to
    1e-9)."""
    x = asb.synth.protein_like(n, f, seed=42)
    x[7] = x[3]                                              # an exact duplicate: d1 = 0 for both
    sample = asb.heuristics.sample_indices(n, 500, 129)
    sample[:2] = (3, 7)
    wd1, wd2 = oracle.twonn_distances(x, sample)
    rctx.set_option("twonn_prefilter", 1)
    d1, d2 = rctx.twonn_distances(x, sample)
    assert rctx.kernel_ms("twonn_pf_used") == 1.0
    assert np.array_equal(np.asarray(d1), wd1) and np.array_equal(np.asarray(d2), wd2)
    assert d1[0] == 0.0 and d1[1] == 0.0
