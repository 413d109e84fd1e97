"""K8 with the certified tensor-core prefilter (option ``search_prefilter``, csrc/search_pf.cuh) against the oracle.

The prefilter only chooses WHICH pairs are scored exactly; the exact scores are computed in the reference's own
arithmetic (sequential, separately rounded sums), so ids must be identical and scores bit-identical to the oracle
(``aso_search_lambda_aware``), ties included.  Inputs the error bound does not cover must come out of the exact
kernel unchanged (fallback), with the reference's error behaviour.

The option is on by default (validated on B200: gpurun_out/pf_validate2.log, profiles/r01_prefilter.md); the
fixture sets it explicitly and ``search_prefilter = 0`` selects the exact FP64 kernel."""
import numpy as np
import pytest

from oracle_binding import TAU_MEDIAN

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[(1, 1), (1, 0), (0, 0)], ids=["tcgen05_bf16x3", "tcgen05_3xtf32", "mma_sync"])
def pctx(ctx, request):
    """The instantiations of the prefilter tile: search_umma = 1 (default: tcgen05.mma + TMA + TMEM, csrc/search_umma.cuh)
    with BF16x3 planes on kind::f16 (default) or 3xTF32 planes on kind::tf32, and search_umma = 0 (mma.sync + cp.async,
    csrc/search_pf.cuh) -- same certificate, same candidate lists, same finishing kernel."""
    ctx.set_option("search_prefilter", 1)
    ctx.set_option("search_umma", request.param[0])
    ctx.set_option("search_umma_bf16", request.param[1])
    try:
        yield ctx
    finally:
        ctx.set_option("search_prefilter", 1)
        ctx.set_option("search_umma", 1)
        ctx.set_option("search_umma_bf16", 1)


def _case(asb, oracle, n, f, nq, seed=42):
    x = asb.synth.protein_like(n, f, seed=seed)
    cent, _, _ = oracle.cluster_incremental(x[:2000], 50, 1.5 * f * 0.0025 * 2)
    csr = oracle.feature_laplacian(cent, eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    lam = oracle.compute_taumode(x, csr, TAU_MEDIAN)
    queries, _ = asb.synth.queries_from_items(x, nq, seed=seed + 1)
    lq = oracle.compute_taumode(queries, csr, TAU_MEDIAN)
    return x, lam, queries, lq


def _bit_equal(got, want, k):
    idx, score, count = (np.asarray(a) for a in got)
    widx, wscore, wcount = want
    assert np.array_equal(count, wcount)
    assert np.array_equal(idx[:, :k], widx[:, :k])
    assert np.array_equal(score[:, :k].view(np.uint64), np.ascontiguousarray(wscore[:, :k]).view(np.uint64))


@pytest.mark.parametrize("n,f,nq,k,alpha", [
    (10_000, 128, 100, 10, 0.7),      # C1 shape
    (5_001, 384, 33, 10, 0.7),        # ragged item / query tiles
    (3_000, 25, 7, 1, 0.7),           # f not a multiple of 32 (zero-padded chunk), k = 1
    (20_000, 100, 300, 32, 0.7),      # largest k of the prefilter path, three query tiles
    (8_192, 64, 128, 10, 1.0),        # pure cosine (the lambda term vanishes)
    (8_192, 64, 128, 10, 0.0),        # pure lambda proximity: massive near-ties in s~ -> many candidates
    (40_000, 384, 64, 10, 0.9),       # several slabs per query tile
])
def test_prefilter_matches_oracle_bitwise(pctx, asb, oracle, n, f, nq, k, alpha):
    x, lam, queries, lq = _case(asb, oracle, n, f, nq)
    want = oracle.search_lambda_aware_batch(x, lam, queries, lq, k, alpha)
    got = pctx.search_lambda_aware_batch(x, lam, queries, lq, k, alpha)
    used = pctx.kernel_ms("search_pf_used")
    flags = pctx.kernel_ms("search_pf_flags")
    if used == 1.0:
        assert pctx.kernel_ms("search_pf_umma") == pctx.get_option("search_umma", 1.0)   # the requested tile ran
        _bit_equal(got, want, k)
        assert pctx.kernel_ms("search_pf_rescored") >= nq * min(k, n)
    else:   # overflow -> exact kernel: the usual parity bar
        assert flags != 0
        assert np.array_equal(np.asarray(got[0]), want[0])
        assert np.allclose(np.asarray(got[1]), want[1], rtol=0, atol=1e-12)
    if alpha not in (0.0,):
        assert used == 1.0, f"prefilter fell back (flags={flags}) on a case it should cover"


def test_prefilter_candidate_volume_is_small(pctx, asb, oracle):
    """The point of the scheme: ~k ln(N/k) candidates per query see the exact arithmetic, not N."""
    n, f, nq, k = 60_000, 128, 256, 10
    x, lam, queries, lq = _case(asb, oracle, n, f, nq, seed=7)
    pctx.search_lambda_aware_batch(x, lam, queries, lq, k, 0.7)
    assert pctx.kernel_ms("search_pf_used") == 1.0
    emitted = pctx.kernel_ms("search_pf_candidates") / nq
    rescored = pctx.kernel_ms("search_pf_rescored") / nq
    assert k <= rescored <= emitted <= pctx.kernel_ms("search_pf_cap")
    assert rescored < 0.02 * n


def test_prefilter_falls_back_where_the_bound_does_not_hold(pctx, asb, oracle, golden):
    db = golden["proteins"]
    q = np.ascontiguousarray(db[[2, 5]])
    lq = np.array([0.5, 0.4])
    # every row 200 times: all scores tie in blocks of 200 -> ties must resolve to the lower index
    dup = np.ascontiguousarray(np.vstack([db[:8]] * 200))
    lam = np.full(len(dup), 0.5)
    for k in (1, 7, 32):
        want = oracle.search_lambda_aware_batch(dup, lam, q, lq, k, 0.7)
        got = pctx.search_lambda_aware_batch(dup, lam, q, lq, k, 0.7)
        assert np.array_equal(np.asarray(got[0])[:, :k], want[0][:, :k])
    big = np.ascontiguousarray(np.vstack([db] * 20))               # 1280 rows: above the prefilter's minimum
    lam = np.full(len(big), 0.3)
    bad = big.copy()
    bad[7, 7] = np.inf                                             # reference: NaN score -> panic (core.rs:785)
    with pytest.raises(asb.ArrowSpaceError) as ei:
        pctx.search_lambda_aware_batch(bad, lam, q, lq, 3, 0.7)
    assert ei.value.status == 9
    assert pctx.kernel_ms("search_pf_used") == 0.0
    with pytest.raises(asb.ArrowSpaceError) as ei:                 # core.rs:773-776
        pctx.search_lambda_aware_batch(big, lam, q, np.array([0.5, 0.0]), 3, 0.7)
    assert ei.value.status == 5
    bad[7, 7] = np.nan                                             # NaN item -> cos 0; NaN lambda -> lam term 0
    lam_nan = lam.copy()
    lam_nan[9] = np.nan
    want = oracle.search_lambda_aware_batch(bad, lam_nan, q, lq, 16, 0.7)
    got = pctx.search_lambda_aware_batch(bad, lam_nan, q, lq, 16, 0.7)
    assert np.array_equal(np.asarray(got[0])[:, :16], want[0][:, :16])
    zero = big.copy()
    zero[4] = 0.0                                                  # zero vector: cosine 0 on both paths, no fallback
    huge = big * 1e150                                             # |x|^2 overflows the certified range -> fallback
    for data in (zero, huge):
        want = oracle.search_lambda_aware_batch(data, lam, q, lq, 10, 0.7)
        got = pctx.search_lambda_aware_batch(data, lam, q, lq, 10, 0.7)
        assert np.array_equal(np.asarray(got[0])[:, :10], want[0][:, :10])
        assert np.allclose(np.asarray(got[1])[:, :10], want[1][:, :10], rtol=0, atol=1e-12)
    assert pctx.kernel_ms("search_pf_used") == 0.0            # the last call (huge) fell back


def test_prefilter_through_the_index_handle(pctx, asb, oracle):
    """ArrowSpaceBuilder.build + search_batch with the option set: same answers as without it."""
    x = asb.synth.protein_like(12_000, 128, seed=42)
    b = (asb.ArrowSpaceBuilder.new(pctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_synthesis(asb.TauMode.Median)
         .with_seed(42).with_inline_sampling(None).with_dims_reduction(False, None)
         .with_cluster_params(64, 1.5 * 128 * 0.0025 * 2))
    aspace, gl = b.build(x)
    queries, _ = asb.synth.queries_from_items(x, 200, seed=43)
    idx, score, count, lq = aspace.search_batch(queries, 10, 0.7)
    assert pctx.kernel_ms("search_pf_used") == 1.0
    pctx.set_option("search_prefilter", 0)
    idx0, score0, count0, _ = aspace.search_batch(queries, 10, 0.7)
    assert pctx.kernel_ms("search_pf_used") == 0.0 and pctx.kernel_ms("search_kernel") > 0   # the exact kernel ran
    assert np.array_equal(np.asarray(idx), np.asarray(idx0))
    assert np.allclose(np.asarray(score), np.asarray(score0), rtol=0, atol=1e-12)
    want = oracle.search_lambda_aware_batch(x, np.asarray(aspace.lambdas), queries, np.asarray(lq), 10, 0.7)
    assert np.array_equal(np.asarray(idx), want[0])

def test_prefilter_sends_only_the_overflowing_queries_to_the_exact_kernel(pctx, asb, oracle):
    """A blob of near-identical items puts thousands of scores inside the band of the queries drawn from it: their
    survivor lists overflow (PF_MAXSEL) and they alone take the exact FP64 kernel + reference-order rescoring, while the
    other queries keep the prefilter's answer.  Either way the ids and scores are the oracle's."""
    n, f, nq, k = 24_000, 64, 96, 10
    x, lam, queries, lq = _case(asb, oracle, n, f, nq, seed=11)
    rng = np.random.RandomState(5)
    x[:6_000] = x[0] * (1.0 + 1e-7 * rng.randn(6_000, 1)) + 1e-7 * rng.rand(6_000, f)
    x = np.ascontiguousarray(np.abs(x))
    csr = oracle.feature_laplacian(oracle.cluster_incremental(x[6_000:8_000], 50, 1.5 * f * 0.0025 * 2)[0],
                                   eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    lam = oracle.compute_taumode(x, csr, TAU_MEDIAN)
    queries[:16] = x[:16] * 1.01
    lq = oracle.compute_taumode(queries, csr, TAU_MEDIAN)
    want = oracle.search_lambda_aware_batch(x, lam, queries, lq, k, 0.9)
    got = pctx.search_lambda_aware_batch(x, lam, queries, lq, k, 0.9)
    assert pctx.kernel_ms("search_pf_used") == 1.0
    novf = pctx.kernel_ms("search_pf_overflow_queries")
    assert 16 <= novf < nq // 2, novf
    assert np.allclose(np.asarray(got[1]), want[1], rtol=0, atol=1e-12)
    same = np.asarray(got[0]) == want[0]
    gap_ok = np.abs(np.asarray(got[1]) - want[1]) < 1e-12          # ids may differ only inside exact score ties
    assert (same | gap_ok).all()
    assert np.array_equal(np.asarray(got[0])[16:], want[0][16:])   # the untouched queries: bit-identical as ever
    assert np.array_equal(np.asarray(got[1])[16:].view(np.uint64), want[1][16:].view(np.uint64))
