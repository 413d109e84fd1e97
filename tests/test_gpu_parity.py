"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): CSR indptr/indices and top-k id lists
bit-exact (score gaps < 1e-9 excepted), cluster outputs bit-exact, lambda within 1e-9 relative.
All tests here need a B200: ``pytest -m gpu``."""
import numpy as np
import pytest

from oracle_binding import TAU_FIXED, TAU_MEAN, TAU_MEDIAN, TAU_PERCENTILE

pytestmark = pytest.mark.gpu

LAMBDA_RTOL = 1e-9      # north_star: lambda-tau within 1e-9 relative in f64
SCORE_GAP = 1e-9        # top-k ids may differ only where the score gap is below this


def _tm(asb, mode, value=0.0):
    return asb.TauMode(mode, value)


def _assert_lambda_close(got, want, rtol=LAMBDA_RTOL):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w)
    ok = ~nan_w
    err = np.abs(got[ok] - want[ok])
    tol = rtol * np.maximum(np.abs(want[ok]), 1e-300) + 1e-15
    bad = err > tol
    assert not bad.any(), f"max rel err {np.max(err / np.maximum(np.abs(want[ok]), 1e-300))}"


def _graph(oracle, cent, **kw):
    p = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    p.update(kw)
    return oracle.feature_laplacian(cent, **p)


# =============================================================================== taumode (K5-K7)
@pytest.mark.parametrize("mode,value", [(TAU_MEDIAN, 0.0), (TAU_MEAN, 0.0), (TAU_PERCENTILE, 0.25),
                                        (TAU_PERCENTILE, 0.9), (TAU_FIXED, 0.3), (TAU_FIXED, -1.0)])
def test_taumode_proteins(ctx, asb, oracle, golden, mode, value):
    db = golden["proteins"]
    csr = _graph(oracle, db[:20])
    want = oracle.compute_taumode(db, csr, mode, value)
    lam, n2, stats = ctx.compute_taumode(db, csr, _tm(asb, mode, value), want_norms=True)
    _assert_lambda_close(lam, want)
    assert np.allclose(n2, (db * db).sum(1), rtol=1e-14)
    assert np.isclose(stats[0], want.min(), rtol=1e-9) and np.isclose(stats[1], want.max(), rtol=1e-9)
    assert np.isclose(stats[2], want.sum(), rtol=1e-9)
    assert want.std() > 0 and np.all((want >= 0) & (want <= 1))   # tests/test_taumode.rs:162-315


def test_taumode_quora_odd_and_even_lengths(ctx, asb, oracle, golden):
    q = golden["quora"]                                  # 15 x 384, signed values -> many taus hit the floor
    csr = _graph(oracle, np.abs(q), eps=0.9, topk=6)
    for mode, value in [(TAU_MEDIAN, 0.0), (TAU_MEAN, 0.0), (TAU_PERCENTILE, 0.5)]:
        for data in (q, np.abs(q)):
            want = oracle.compute_taumode(data, csr, mode, value)
            lam, _, _ = ctx.compute_taumode(data, csr, _tm(asb, mode, value))
            _assert_lambda_close(lam, want)
    # odd feature count (F=383): median takes the single middle element
    qo = np.ascontiguousarray(np.abs(q[:, :383]))
    csr_o = _graph(oracle, qo, eps=0.9, topk=6)
    want = oracle.compute_taumode(qo, csr_o, TAU_MEDIAN)
    lam, _, _ = ctx.compute_taumode(qo, csr_o, _tm(asb, TAU_MEDIAN))
    _assert_lambda_close(lam, want)


def test_taumode_tau_selection_exact(ctx, asb, oracle):
    """With L = I the synthetic lambda is tau/(1+tau): isolates select_tau (taumode.rs:87-127),
    including duplicates, negative values (floor) and every rank of Percentile."""
    rng = np.random.RandomState(3)
    f = 37
    eye = (np.arange(f + 1, dtype=np.int64), np.arange(f, dtype=np.int64), np.ones(f))
    rows = [rng.rand(f), np.round(rng.rand(f) * 4) / 4, np.full(f, 0.5), -rng.rand(f), rng.randn(f),
            np.concatenate([np.zeros(20), rng.rand(17)]), rng.rand(f) * 1e-12, rng.rand(f) * 1e6]
    x = np.ascontiguousarray(np.vstack(rows))
    for mode, value in [(TAU_MEDIAN, 0)] + [(TAU_PERCENTILE, p) for p in np.linspace(0, 1, 13)] + [(TAU_MEAN, 0)]:
        want = oracle.compute_taumode(x, eye, mode, float(value))
        lam, _, _ = ctx.compute_taumode(x, eye, _tm(asb, mode, float(value)))
        assert np.allclose(lam, want, rtol=1e-14, atol=0), (mode, value)
    xe = np.ascontiguousarray(x[:, :36])                 # even length -> 0.5*(a+b)
    eye_e = (np.arange(37, dtype=np.int64), np.arange(36, dtype=np.int64), np.ones(36))
    want = oracle.compute_taumode(xe, eye_e, TAU_MEDIAN)
    lam, _, _ = ctx.compute_taumode(xe, eye_e, _tm(asb, TAU_MEDIAN))
    assert np.allclose(lam, want, rtol=1e-14, atol=0)


def test_taumode_nonfinite_items(ctx, asb, oracle, golden):
    db = golden["proteins"].copy()
    csr = _graph(oracle, db[:20])
    db[5, 3] = np.nan
    db[9, 0] = np.inf
    want = oracle.compute_taumode(db, csr, TAU_MEDIAN)
    lam, _, _ = ctx.compute_taumode(db, csr, _tm(asb, TAU_MEDIAN))
    _assert_lambda_close(lam, want)


@pytest.mark.parametrize("n,f", [(10_000, 128), (4_097, 384), (777, 1024), (300, 2000)])
def test_taumode_synthetic(ctx, asb, oracle, n, f):
    x = asb.synth.protein_like(n, f, seed=42)
    cent, _, _ = oracle.cluster_incremental(x[:2000], min(64, n // 10), 1.5 * f * 0.0025 * 2)
    csr = _graph(oracle, cent)
    assert csr[0][-1] > f, "graph must be non-degenerate"
    want = oracle.compute_taumode(x, csr, TAU_MEDIAN)
    lam, _, _ = ctx.compute_taumode(x, csr, _tm(asb, TAU_MEDIAN))
    _assert_lambda_close(lam, want)
    assert np.all(np.isfinite(want)) and want.std() > 0


@pytest.mark.parametrize("variant", ["regs_ipp2", "regs_ipp1", "smem_v3"])
@pytest.mark.parametrize("n,f", [(2_001, 24), (1_500, 100), (3_333, 384), (901, 500), (700, 770), (333, 1000), (100, 1100)])
def test_taumode_symmetric_kernel_variants(ctx, asb, oracle, n, f, variant):
    """The three one-warp-per-item kernels (item in registers with one / two items per pass, item in shared memory
    only) against the oracle: every register width NPL, widths that are not multiples of 32, an odd item count (the
    last pass is half empty), non-finite values, Median / Percentile / Mean."""
    x = asb.synth.protein_like(n, f, seed=7)
    cent, _, _ = oracle.cluster_incremental(x[:1500], 40, 1.5 * f * 0.0025 * 2)
    csr = _graph(oracle, cent)
    x[3, f // 2] = np.nan
    x[n - 1, 0] = -np.inf
    x[n // 2, f - 1] = np.inf
    ctx.set_option("taumode_regs", 0 if variant == "smem_v3" else 1)
    ctx.set_option("taumode_ipp", 1 if variant == "regs_ipp1" else 2)
    try:
        for mode, value in [(TAU_MEDIAN, 0.0), (TAU_PERCENTILE, 0.9), (TAU_MEAN, 0.0)]:
            want = oracle.compute_taumode(x, csr, mode, value)
            lam, n2, _ = ctx.compute_taumode(x, csr, _tm(asb, mode, value), want_norms=True)
            _assert_lambda_close(lam, want)
            fin = np.isfinite(x).all(axis=1)
            assert np.allclose(n2[fin], (x[fin] ** 2).sum(axis=1), rtol=1e-13)
    finally:
        ctx.set_option("taumode_regs", 1)
        ctx.set_option("taumode_ipp", 1)


def test_taumode_generic_and_symmetric_kernels_agree(ctx, asb, oracle, golden):
    """Symmetric graphs take the edge-once kernel, anything else the generic CSR kernel; both must
    match the oracle (non-symmetric, positive off-diagonal and diagonal-free matrices included)."""
    x = asb.synth.protein_like(3_000, 128, seed=42)
    cent, _, _ = oracle.cluster_incremental(x[:2000], 50, 1.5 * 128 * 0.0025 * 2)
    csr = _graph(oracle, cent)
    want = oracle.compute_taumode(x, csr, TAU_MEDIAN)
    lam_sym, _, _ = ctx.compute_taumode(x, csr, _tm(asb, TAU_MEDIAN))
    ctx.set_option("taumode_generic", 1)
    try:
        lam_gen, _, _ = ctx.compute_taumode(x, csr, _tm(asb, TAU_MEDIAN))
    finally:
        ctx.set_option("taumode_generic", 0)
    _assert_lambda_close(lam_sym, want)
    _assert_lambda_close(lam_gen, want)
    rng = np.random.RandomState(5)
    f = 24
    db = golden["proteins"]
    for kind in ("nonsym", "mixed_sign_sym", "no_diag"):
        dense = np.zeros((f, f))
        for _ in range(60):
            i, j = rng.randint(0, f, 2)
            if i == j:
                continue
            v = -rng.rand() if kind != "mixed_sign_sym" else rng.randn()
            dense[i, j] = v
            if kind != "nonsym":
                dense[j, i] = v
        if kind != "no_diag":
            dense[np.arange(f), np.arange(f)] = rng.rand(f) + 0.5
        ip, ii, dd = [0], [], []
        for r in range(f):
            for c in range(f):
                if dense[r, c] != 0.0:
                    ii.append(c)
                    dd.append(dense[r, c])
            ip.append(len(ii))
        g = (np.array(ip, dtype=np.int64), np.array(ii, dtype=np.int64), np.array(dd))
        for mode, value in [(TAU_MEDIAN, 0.0), (TAU_FIXED, 0.4)]:
            want = oracle.compute_taumode(db, g, mode, value)
            lam, _, _ = ctx.compute_taumode(db, g, _tm(asb, mode, value))
            _assert_lambda_close(lam, want, rtol=1e-9)


def test_taumode_heavy_duplicates_and_extremes(ctx, asb, oracle):
    """Selection corner cases: long runs of equal values around the median, huge / tiny magnitudes."""
    rng = np.random.RandomState(11)
    f = 96
    eye = (np.arange(f + 1, dtype=np.int64), np.arange(f, dtype=np.int64), np.ones(f))
    rows = [np.concatenate([np.zeros(60), rng.rand(36)]), np.concatenate([np.full(50, 0.25), rng.rand(46)]),
            np.repeat(rng.rand(12), 8), np.full(f, 7.0), rng.rand(f) * 1e-300, rng.rand(f) * 1e300,
            np.concatenate([rng.rand(40), np.full(16, 0.5), rng.rand(40)]), np.sort(rng.rand(f)), -np.sort(rng.rand(f)),
            np.exp(rng.randn(f) * 20)]
    x = np.ascontiguousarray(np.vstack(rows))
    for mode, value in [(TAU_MEDIAN, 0)] + [(TAU_PERCENTILE, p) for p in (0.0, 0.1, 0.33, 0.5, 0.52, 0.9, 1.0)]:
        want = oracle.compute_taumode(x, eye, mode, float(value))
        lam, _, _ = ctx.compute_taumode(x, eye, _tm(asb, mode, float(value)))
        assert np.allclose(lam, want, rtol=1e-14, atol=0, equal_nan=True), (mode, value, lam, want)


def test_taumode_near_constant_rows_return_the_well_conditioned_value(ctx, asb, oracle):
    """Near-constant rows (x = c + tiny ripple): the reference sums x_i L_ij x_j (src/taumode.rs:565-588), which cancels
    catastrophically -- num is the small difference of O(c^2) terms and keeps only ~16 - 2 log10(c / ripple) digits; the
    kernel sums w_ij (x_i - x_j)^2 over the undirected edges (taumode_sym.cuh), which does not cancel.  The two agree
    to 1e-9 wherever the reference's own form is accurate to 1e-9, and where it is not the kernel returns the
    WELL-CONDITIONED value: it matches an exact (fraction arithmetic) evaluation, the oracle does not."""
    from fractions import Fraction
    rng = np.random.default_rng(3)
    f = 48
    cent = asb.synth.protein_like(40, f, seed=9)
    csr = _graph(oracle, cent)
    ip, ii, dd = csr
    rows = []
    for ripple in (1e-2, 1e-5, 1e-7):
        rows.append(3.0 + ripple * rng.standard_normal(f))
    x = np.ascontiguousarray(np.vstack(rows))
    # tau = 1 removes the (1 - tau) G term, which otherwise dominates lambda on such rows and hides the Rayleigh
    # quotient's error (with tau < 1 both forms agree to ~1e-13 even here: lambda >= (1 - tau) / #edges)
    tau = 1.0
    want = oracle.compute_taumode(x, csr, TAU_FIXED, tau)
    got, _, _ = ctx.compute_taumode(x, csr, _tm(asb, TAU_FIXED, tau))

    def exact_lambda(row, tau):
        xs = [Fraction(float(v)) for v in row]
        num, edge = Fraction(0), Fraction(0)
        shares = []
        for r in range(f):
            for e in range(ip[r], ip[r + 1]):
                c, lij = int(ii[e]), Fraction(float(dd[e]))
                num += xs[r] * lij * xs[c]
                if c != r and lij < 0:
                    t = -lij * (xs[r] - xs[c]) ** 2
                    edge += t
                    shares.append(t)
        den = sum(v * v for v in xs)
        energy = num / den if den > Fraction(1, 10 ** 12) else Fraction(0)
        g = sum((t / edge) ** 2 for t in shares) if edge > 0 else Fraction(0)
        g = min(max(g, Fraction(0)), Fraction(1))
        t = Fraction(tau)
        return float(t * (energy / (energy + t)) + (1 - t) * g)

    exact = np.array([exact_lambda(r, tau) for r in x])
    rel_gpu = np.abs(np.asarray(got) - exact) / np.abs(exact)
    rel_ref = np.abs(want - exact) / np.abs(exact)
    assert rel_gpu.max() < 1e-9, rel_gpu                 # the kernel is accurate on every row
    assert rel_ref[0] < 1e-9                              # mild ripple: the reference's form is fine, both agree
    assert abs(got[0] - want[0]) <= 1e-9 * abs(want[0])
    assert rel_ref[2] > 1e-4                              # ripple 1e-7 on a base of 3: the reference's own sum has lost it
    # same rows, tau = 0.3: the dispersion term carries lambda and the two forms agree to the parity bar
    want3 = oracle.compute_taumode(x, csr, TAU_FIXED, 0.3)
    got3, _, _ = ctx.compute_taumode(x, csr, _tm(asb, TAU_FIXED, 0.3))
    _assert_lambda_close(got3, want3)


def test_prepare_query_item(ctx, asb, oracle, golden):
    db = golden["proteins"]
    csr = _graph(oracle, db[:20])
    aspace = asb.ArrowSpace(db, asb.TauMode.Median, ctx)
    gl = asb.GraphLaplacian(*csr, nnodes=64, graph_params=None)
    q = db[3] * 1.02
    want = oracle.prepare_query_item(q, csr, TAU_MEDIAN)
    got = aspace.prepare_query_item(q, gl)
    assert abs(got - want) <= LAMBDA_RTOL * abs(want)
    assert abs(aspace.prepare_query_item(q, gl) - got) <= 1e-10 * abs(got)   # tests/test_querying_proj.rs:161-169
    bad = q.copy()
    bad[2] = np.nan
    with pytest.raises(asb.ArrowSpaceError) as ei:       # tests/test_querying_proj.rs:261-291
        aspace.prepare_query_item(bad, gl)
    assert ei.value.status == 4
    with pytest.raises(asb.ArrowSpaceError) as ei:       # wrong dimension, core.rs:510-516
        aspace.prepare_query_item(q[:10], gl)
    assert ei.value.status == 12


# ================================================================================== search (K8)
def _assert_topk_equal(idx, score, count, widx, wscore, wcount):
    idx, score, count = np.asarray(idx), np.asarray(score), np.asarray(count)
    assert np.array_equal(count, wcount)
    for q in range(idx.shape[0]):
        c = int(count[q])
        assert np.allclose(score[q, :c], wscore[q, :c], rtol=0, atol=1e-12), q
        if not np.array_equal(idx[q, :c], widx[q, :c]):
            for r in np.nonzero(idx[q, :c] != widx[q, :c])[0]:
                # allowed only inside a group of scores closer than SCORE_GAP
                near = np.abs(wscore[q, :c] - wscore[q, r]) < SCORE_GAP
                assert near.sum() > 1 and idx[q, r] in widx[q, :c][near], (q, r)


def test_search_paper_answer(ctx, asb, oracle, golden):
    """paper.md:123-133 / examples/01_compare_cosine.rs:1 through the GPU path."""
    db = golden["proteins"]
    aspace = asb.ArrowSpace(db, asb.TauMode.Median, ctx)
    aspace.lambdas = np.full(64, 0.25)
    res = aspace.search_lambda_aware(asb.ArrowItem.new(db[3] * 1.02, 0.3), 3, 1.0)
    assert [i for i, _ in res] == [3, 6, 0]
    for (_, s), want in zip(res, (1.000000, 0.999573, 0.999325)):
        assert abs(s - want) < 5e-7


@pytest.mark.parametrize("alpha", [1.0, 0.9, 0.7, 0.0])
@pytest.mark.parametrize("k", [1, 3, 10, 33, 64])
def test_search_proteins_parity(ctx, asb, oracle, golden, alpha, k):
    db = golden["proteins"]
    csr = _graph(oracle, db[:20])
    lam = oracle.compute_taumode(db, csr, TAU_MEDIAN)
    queries = np.ascontiguousarray(db[[3, 10, 40, 63]] * 1.02)
    lq = np.array([oracle.prepare_query_item(q, csr, TAU_MEDIAN) for q in queries])
    want = oracle.search_lambda_aware_batch(db, lam, queries, lq, k, alpha)
    got = ctx.search_lambda_aware_batch(db, lam, queries, lq, k, alpha)
    _assert_topk_equal(*got, *want)


@pytest.mark.parametrize("n,f,nq", [(10_000, 128, 100), (5_001, 384, 33), (3_000, 25, 7), (200, 770, 130)])
def test_search_synthetic_parity(ctx, asb, oracle, n, f, nq):
    x = asb.synth.protein_like(n, f, seed=42)
    cent, _, _ = oracle.cluster_incremental(x[:2000], 50, 1.5 * f * 0.0025 * 2)
    csr = _graph(oracle, cent)
    lam = oracle.compute_taumode(x, csr, TAU_MEDIAN)
    queries, _ = asb.synth.queries_from_items(x, nq, seed=43)
    lq = oracle.compute_taumode(queries, csr, TAU_MEDIAN)
    want = oracle.search_lambda_aware_batch(x, lam, queries, lq, 10, 0.7)
    got = ctx.search_lambda_aware_batch(x, lam, queries, lq, 10, 0.7)
    _assert_topk_equal(*got, *want)
    idx, score, _ = got
    assert np.all(np.diff(np.asarray(score), axis=1) <= 0)       # descending, test_querying_proj.rs:369-399
    assert np.all((np.asarray(idx) >= 0) & (np.asarray(idx) < n))


def test_search_ties_and_edges(ctx, asb, oracle, golden):
    db = golden["proteins"]
    dup = np.ascontiguousarray(np.vstack([db[:8]] * 40))          # 320 rows, every row 40 times
    lam = np.full(320, 0.5)
    q = np.ascontiguousarray(db[[2, 5]])
    lq = np.array([0.5, 0.4])
    for k in (1, 7, 41, 64):
        want = oracle.search_lambda_aware_batch(dup, lam, q, lq, k, 0.7)
        got = ctx.search_lambda_aware_batch(dup, lam, q, lq, k, 0.7)
        assert np.array_equal(np.asarray(got[0])[:, :k], want[0][:, :k])   # exact ties -> lower index first
    # k >= N returns all N (core.rs:786); k == 0 returns nothing
    idx, score, count = ctx.search_lambda_aware_batch(db[:5], np.full(5, 0.3), q, lq, 64, 0.7)
    assert count.tolist() == [5, 5]
    want = oracle.search_lambda_aware_batch(db[:5], np.full(5, 0.3), q, lq, 64, 0.7)
    assert np.array_equal(np.asarray(idx)[:, :5], want[0][:, :5])
    _, _, count = ctx.search_lambda_aware_batch(db, np.full(64, 0.3), q, lq, 0, 0.7)
    assert count.tolist() == [0, 0]
    with pytest.raises(asb.ArrowSpaceError) as ei:                # core.rs:773-776
        ctx.search_lambda_aware_batch(db, np.full(64, 0.3), q, np.array([0.5, 0.0]), 3, 0.7)
    assert ei.value.status == 5
    bad = db.copy()
    bad[7, 7] = np.inf                                            # inf item: cos = inf/inf = NaN
    with pytest.raises(asb.ArrowSpaceError) as ei:                # core.rs:785 partial_cmp().unwrap() on NaN
        ctx.search_lambda_aware_batch(bad, np.full(64, 0.3), q, lq, 3, 0.7)
    assert ei.value.status == 9
    bad[7, 7] = np.nan                                            # NaN item: norm NaN -> `denom > 0` false -> cos 0
    lam_nan = np.full(64, 0.3)
    lam_nan[9] = np.nan                                           # NaN lambda: f64::min(NaN, 1) = 1 -> lam term 0
    want = oracle.search_lambda_aware_batch(bad, lam_nan, q, lq, 64, 0.7)
    got = ctx.search_lambda_aware_batch(bad, lam_nan, q, lq, 64, 0.7)
    _assert_topk_equal(*got, *want)
    zero = db.copy()
    zero[4] = 0.0                                                 # zero vector -> cosine 0 (core.rs:231-236)
    want = oracle.search_lambda_aware_batch(zero, np.full(64, 0.3), q, lq, 64, 0.7)
    got = ctx.search_lambda_aware_batch(zero, np.full(64, 0.3), q, lq, 64, 0.7)
    _assert_topk_equal(*got, *want)


def test_topk_merge_matches_single_shard(ctx, asb, oracle):
    x = asb.synth.protein_like(4_000, 64, seed=5)
    lam = np.linspace(0.1, 0.9, 4_000)
    queries, _ = asb.synth.queries_from_items(x, 20, seed=6)
    lq = np.full(20, 0.5)
    k = 10
    full = ctx.search_lambda_aware_batch(x, lam, queries, lq, k, 0.7)
    parts_s, parts_i = [], []
    bounds = [0, 1000, 1001, 2500, 4000]                          # ragged shards incl. a 1-row shard
    for a, b in zip(bounds[:-1], bounds[1:]):
        i, s, c = ctx.search_lambda_aware_batch(np.ascontiguousarray(x[a:b]), lam[a:b], queries, lq, k, 0.7,
                                                index_offset=a)
        i, s = np.asarray(i).copy(), np.asarray(s).copy()
        for q in range(20):
            i[q, int(c[q]):] = -1
        parts_s.append(s)
        parts_i.append(i)
    ms, mi, mc = ctx.topk_merge(np.stack(parts_s), np.stack(parts_i), len(parts_s), 20, k)
    # shards below the prefilter's minimum size are scored by the exact DMMA kernel, the full set by the rescoring
    # pass (reference summation order): same ids, scores equal to the last few bits
    assert np.array_equal(mi, np.asarray(full[0])) and np.allclose(ms, np.asarray(full[1]), rtol=0, atol=1e-14)
    assert mc.tolist() == [k] * 20


# ============================================================================== Laplacian (K3+K4)
def _assert_csr_equal(got, want):
    gip, gii, gdd = got
    wip, wii, wdd = want
    assert np.array_equal(gip, wip), "indptr must be bit-exact"
    assert np.array_equal(gii, wii), "indices must be bit-exact"
    assert np.allclose(gdd, wdd, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("kw", [dict(), dict(eps=0.2, topk=3), dict(eps=1.0, topk=7, sigma=None),
                                dict(self_included=True), dict(rectified=True), dict(p=1.5, sigma=0.1),
                                dict(eps=1e-3, topk=4, sigma=None)])
def test_laplacian_proteins(ctx, asb, oracle, golden, kw):
    cent = golden["proteins"][:20]
    p = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    p.update(kw)
    want = oracle.feature_laplacian(cent, **p)
    gp = asb.GraphParams(p["eps"], p["k"], p["topk"], p["p"], p["sigma"], self_included=p.get("self_included", False),
                         rectified=p.get("rectified", False))
    got = ctx.build_feature_laplacian(cent, gp)
    _assert_csr_equal(got, want)


@pytest.mark.parametrize("x,f,topk,k", [(100, 128, 4, 12), (316, 384, 4, 12), (200, 1024, 12, 20), (40, 33, 3, 6)])
def test_laplacian_synthetic(ctx, asb, oracle, x, f, topk, k):
    rows = asb.synth.protein_like(4000, f, seed=11)
    cent, _, _ = oracle.cluster_incremental(rows, x, 1.5 * f * 0.0025 * 2)
    want = oracle.feature_laplacian(cent, eps=0.5, k=k, topk=topk, p=2.0, sigma=0.25)
    got = ctx.build_feature_laplacian(cent, asb.GraphParams(0.5, k, topk, 2.0, 0.25))
    _assert_csr_equal(got, want)
    ip, ii, dd = got
    dense = np.zeros((f, f))
    for r in range(f):
        dense[r, ii[ip[r]:ip[r + 1]]] = dd[ip[r]:ip[r + 1]]
    assert np.array_equal(dense, dense.T) and np.all(np.abs(dense.sum(1)) < 1e-12)   # test_laplacian.rs:51-152


def test_laplacian_inline_sparsification(ctx, asb, oracle):
    """mean degree > 10 -> keep the top half by w*sqrt(deg_i*deg_j) (laplacian.rs:229-280)."""
    rows = asb.synth.protein_like(3000, 256, seed=12)
    cent, _, _ = oracle.cluster_incremental(rows, 64, 1.5 * 256 * 0.0025 * 2)
    want = oracle.feature_laplacian(cent, eps=1.0, k=20, topk=14, p=2.0, sigma=0.25)
    got = ctx.build_feature_laplacian(cent, asb.GraphParams(1.0, 20, 14, 2.0, 0.25))
    _assert_csr_equal(got, want)
    assert want[0][-1] < 256 * (1 + 2 * 15) * 0.8            # sparsification really happened


def test_laplacian_errors(ctx, asb, golden):
    gp = asb.GraphParams(0.5, 6, 3, 2.0, None)
    with pytest.raises(asb.ArrowSpaceError) as ei:               # laplacian.rs:129-134
        ctx.build_feature_laplacian(golden["proteins"][:1], gp)
    assert ei.value.status == 6
    with pytest.raises(asb.ArrowSpaceError) as ei:               # graph.rs:185-193
        ctx.build_feature_laplacian(golden["quora"], asb.GraphParams(1e-9, 6, 3, 2.0, None, sparsity_check=True))
    assert ei.value.status == 7
    z = golden["proteins"][:10].copy()
    z[:, 5] = 0.0
    with pytest.raises(asb.ArrowSpaceError) as ei:               # zero-magnitude feature column
        ctx.build_feature_laplacian(z, gp)
    assert ei.value.status == 10


@pytest.mark.parametrize("normalise", [1, 2])
@pytest.mark.parametrize("x,f,eps", [(20, 24, 1.5), (100, 128, 1.2), (316, 384, 1.0)])
def test_laplacian_with_normalisation(ctx, asb, oracle, golden, normalise, x, f, eps):
    """normalise = true (src/laplacian.rs:146-151): column standardisation of the F x X matrix before the kNN.  Centred
    columns make cosine distances spread over [0, 2], so eps is wider than in the other tests.  normalise = 1 / 2 selects
    the variance convention (population / sample); smartcore's StandardScaler itself is unpinned."""
    cent = golden["proteins"][:x] if (x, f) == (20, 24) else asb.synth.protein_like(x, f, seed=12)
    gp = dict(eps=eps, k=12, topk=4, p=2.0, sigma=0.5)
    want = oracle.feature_laplacian(cent, normalise=normalise, **gp)
    got = ctx.build_feature_laplacian(cent, asb.GraphParams(eps, 12, 4, 2.0, 0.5, normalise=normalise))
    _assert_csr_equal(got, want)
    plain = oracle.feature_laplacian(cent, **gp)
    assert want[0][-1] != plain[0][-1] or not np.array_equal(want[1], plain[1])      # it really changes the graph
    d = np.zeros((f, f))
    for r in range(f):
        d[r, want[1][want[0][r]:want[0][r + 1]]] = want[2][want[0][r]:want[0][r + 1]]
    assert np.allclose(d.sum(1), 0.0, atol=1e-12) and np.allclose(d, d.T)             # still a Laplacian


# ================================================================================ clustering (K2)
def _assert_cluster_equal(got, want):
    gc, ga, gs = got
    wc, wa, ws = want
    assert gc.shape == wc.shape, (gc.shape, wc.shape)
    assert np.array_equal(ga, wa), "assignments must be identical"
    assert np.array_equal(gs, ws), "cluster sizes must be identical"
    assert np.array_equal(gc.view(np.uint64), wc.view(np.uint64)), "centroids must be bit-identical"


@pytest.mark.parametrize("n,f,maxk,rscale", [(5_000, 64, 40, 1.0), (20_000, 128, 100, 1.0), (3_000, 384, 316, 0.6),
                                             (2_000, 33, 7, 1.0), (4_000, 130, 1001, 0.2), (500, 24, 3, 3.0)])
@pytest.mark.parametrize("replay", [1, 0])
def test_cluster_parity(ctx, asb, oracle, n, f, maxk, rscale, replay):
    """replay = 1: the default path (sequential prefix + certified parallel replay, csrc/cluster_replay.cu);
    replay = 0: the sequential kernel walks every row."""
    x = asb.synth.protein_like(n, f, seed=21)
    radius = rscale * 1.5 * f * 0.0025 * 2
    want = oracle.cluster_incremental(x, maxk, radius)
    ctx.set_option("cluster_replay", replay)
    try:
        got = ctx.cluster_incremental(x, maxk, radius)
    finally:
        ctx.set_option("cluster_replay", 1)
    _assert_cluster_equal(got, want)
    assert (want[1] >= 0).sum() > 0


@pytest.fixture()
def seqctx(ctx):
    """The sequential clustering kernels on their own (the tests below select and inspect kernel variants)."""
    ctx.set_option("cluster_replay", 0)
    try:
        yield ctx
    finally:
        ctx.set_option("cluster_replay", 1)


@pytest.mark.parametrize("n,f,maxk,rscale", [(60_000, 64, 100, 1.0), (30_000, 128, 316, 0.8), (8_000, 770, 64, 1.0),
                                             (5_000, 33, 40, 0.7), (12_000, 384, 449, 1.0), (6_000, 768, 1001, 1.0)])
def test_cluster_parity_long_runs_all_variants(seqctx, asb, oracle, n, f, maxk, rscale):
    """Long walks exercise the blocked kernel's interval certification (counts grow, displacements
    shrink, blocks get longer); the row-wise kernel must give the same bits."""
    x = asb.synth.protein_like(n, f, seed=77)
    radius = rscale * 1.5 * f * 0.0025 * 2
    want = oracle.cluster_incremental(x, maxk, radius)
    got = seqctx.cluster_incremental(x, maxk, radius)
    _assert_cluster_equal(got, want)
    blocks = seqctx.kernel_ms("cluster_blocks")
    assert 0 < blocks <= n
    # -2: pipelined FP32-prefilter kernel (default), -1: FP32-prefilter kernel, 0/1: FP64 blocked kernel
    # (16 / 8 rows per barrier), 2: row-wise kernel.
    # A variant that does not fit shared memory for this shape falls through to the next one.
    seen = {seqctx.kernel_ms("cluster_variant")}
    for first in (-1, 0, 1, 2):
        seqctx.set_option("cluster_first_variant", first)
        try:
            got = seqctx.cluster_incremental(x, maxk, radius)
            used = seqctx.kernel_ms("cluster_variant")
        finally:
            seqctx.set_option("cluster_first_variant", -9)
        assert used >= first
        seen.add(used)
        _assert_cluster_equal(got, want)
    assert 2.0 in seen


@pytest.mark.parametrize("n,f,maxk,ring", [(9_000, 384, 700, 6), (9_000, 384, 896, 4), (5_000, 128, 1500, 8)])
def test_cluster_pipelined_with_short_row_ring(seqctx, asb, oracle, n, f, maxk, ring):
    """Large centroid counts (the 4- and 8-GPU runs use K = 634 / 896) leave less shared memory for the row ring of
    the pipelined kernel: 6 or 4 groups instead of 8 (less prefetch, same bits)."""
    x = asb.synth.protein_like(n, f, seed=13, n_blobs=256)
    radius = 0.6 * 1.5 * f * 0.0025 * 2
    want = oracle.cluster_incremental(x, maxk, radius)
    got = seqctx.cluster_incremental(x, maxk, radius)
    assert seqctx.kernel_ms("cluster_variant") == -2.0 and seqctx.kernel_ms("cluster_ring_groups") == ring
    _assert_cluster_equal(got, want)
    assert want[0].shape[0] > 300          # enough centroids to fill several tiles per CTA


@pytest.mark.parametrize("n,f,maxk", [(6_000, 384, 100), (4_000, 132, 64), (3_000, 33, 20)])
def test_cluster_tensor_tile_error_model(seqctx, asb, oracle, n, f, maxk):
    """The certified decisions of the pipelined kernel rest on an error bound for the 3xTF32 tensor-core
    distances (exact products, truncating accumulation -- an assumption about the hardware).
    `cluster_check_tile` recomputes every distance of the walk in FP64 from the same FP32 operands and
    reports the worst |error| / bound: it has to stay well inside the bound, and the walk stays bit-exact."""
    x = asb.synth.protein_like(n, f, seed=5)
    radius = 1.5 * f * 0.0025 * 2
    want = oracle.cluster_incremental(x, maxk, radius)
    seqctx.set_option("cluster_check_tile", 1)
    try:
        got = seqctx.cluster_incremental(x, maxk, radius)
        assert seqctx.kernel_ms("cluster_variant") == -2.0
        worst = seqctx.kernel_ms("cluster_phase47") * 1e-12
    finally:
        seqctx.set_option("cluster_check_tile", 0)
        seqctx.set_option("cluster_phase_times", 0)
    _assert_cluster_equal(got, want)
    assert 0.0 < worst < 0.25, worst


def test_cluster_resume_equals_single_walk(ctx, asb, oracle):
    """Shard 0, then shard 1 resumed from shard 0's state == one walk (the multi-GPU hand-off)."""
    x = asb.synth.protein_like(9_000, 96, seed=78)
    radius = 1.5 * 96 * 0.0025 * 2
    maxk = 80
    want = oracle.cluster_incremental(x, maxk, radius)
    cent = np.zeros((maxk, 96))
    sizes = np.zeros(maxk, dtype=np.uint64)
    k, asg = 0, []
    for a, b in [(0, 1), (1, 4000), (4000, 4001), (4001, 9000)]:
        k, part = ctx.cluster_incremental_resume(np.ascontiguousarray(x[a:b]), maxk, radius, cent, sizes, k)
        asg.append(part)
    _assert_cluster_equal((cent[:k], np.concatenate(asg), sizes[:k]), want)


def test_cluster_exact_path_and_ties(seqctx, asb, oracle):
    """Integer-valued data makes many distances tie exactly -> the certified fast path must hand
    those rows to the reference-arithmetic path; forcing that path for ALL rows gives the same
    result."""
    rng = np.random.RandomState(9)
    x = np.ascontiguousarray(rng.randint(0, 3, size=(3000, 16)).astype(np.float64))
    for radius in (2.0, 4.0, 7.0):
        want = oracle.cluster_incremental(x, 24, radius)
        got = seqctx.cluster_incremental(x, 24, radius)
        _assert_cluster_equal(got, want)
    assert seqctx.kernel_ms("cluster_exact_rows") > 0
    y = asb.synth.protein_like(3_000, 64, seed=22)
    want = oracle.cluster_incremental(y, 30, 0.5)
    seqctx.set_option("cluster_force_exact", 1)
    try:
        got = seqctx.cluster_incremental(y, 30, 0.5)
        assert seqctx.kernel_ms("cluster_exact_rows") == 3_000 - 1
    finally:
        seqctx.set_option("cluster_force_exact", 0)
    _assert_cluster_equal(got, want)


def test_cluster_nan_rows_and_duplicates(ctx, asb, oracle):
    x = asb.synth.protein_like(1_000, 32, seed=23)
    x[10] = x[3]                      # exact duplicate of an earlier row
    x[500, 4] = np.nan                # NaN row: never nearest to anything (clustering.rs:922)
    x[700] = x[3]
    want = oracle.cluster_incremental(x, 20, 0.3)
    got = ctx.cluster_incremental(x, 20, 0.3)
    gc, ga, gs = got
    wc, wa, ws = want
    assert np.array_equal(ga, wa) and np.array_equal(gs, ws)
    assert np.array_equal(np.isnan(gc), np.isnan(wc))
    m = ~np.isnan(wc)
    assert np.array_equal(gc[m], wc[m])


# ==================================================================================== Two-NN (K1)
@pytest.mark.parametrize("prefilter", [1, 0], ids=["tcgen05_prefilter", "fp64_kernel"])
@pytest.mark.parametrize("n,f,s", [(5_000, 64, 100), (2_000, 384, 500), (300, 25, 300), (70_000, 128, 500)])
def test_twonn_parity(ctx, asb, oracle, n, f, s, prefilter):
    """prefilter = 1 (default): certified TF32 ranking + direct-form distances, bit-identical to the oracle;
    0: the FP64 tensor kernel with exact rescoring of 4 candidates, distances to 1e-9."""
    x = asb.synth.protein_like(n, f, seed=31)
    x[17] = x[5]                      # exact duplicate -> d1 == 0 for both
    si = asb.heuristics.sample_indices(n, s, 129)
    si[0], si[1] = 5, 17
    w1, w2 = oracle.twonn_distances(x, si)
    ctx.set_option("twonn_prefilter", prefilter)
    try:
        g1, g2 = ctx.twonn_distances(x, si)
        used = ctx.kernel_ms("twonn_pf_used")
    finally:
        ctx.set_option("twonn_prefilter", 1)
    assert used == float(prefilter)
    assert g1[0] == 0.0 and g1[1] == 0.0
    assert np.allclose(g1, w1, rtol=1e-9, atol=0) and np.allclose(g2, w2, rtol=1e-9, atol=0)
    if prefilter:
        assert np.array_equal(np.asarray(g1), w1) and np.array_equal(np.asarray(g2), w2)
    assert asb.heuristics.intrinsic_dim_from_distances(n, f, g1, g2) == oracle.intrinsic_dim(n, f, w1, w2)


# =============================================================== whole build + stage equivalence
def _oracle_build(oracle, x, maxk, radius, gp, mode=TAU_MEDIAN, value=0.0):
    cent, asg, sizes = oracle.cluster_incremental(x, maxk, radius)
    csr = oracle.feature_laplacian(cent, **gp)
    lam = oracle.compute_taumode(x, csr, mode, value)
    return cent, asg, sizes, csr, lam


@pytest.mark.parametrize("n,f,maxk", [(10_000, 128, 100), (3_000, 384, 150)])
def test_build_matches_oracle_and_stages(ctx, asb, oracle, n, f, maxk):
    """ArrowSpaceBuilder::build vs the oracle pipeline, and the reference's own stage-equivalence
    contract (tests/test_eigenmaps.rs:117-329): monolithic build == the four EigenMaps stages."""
    x = asb.synth.protein_like(n, f, seed=42)
    radius = 1.5 * f * 0.0025 * 2
    gp = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    cent, asg, sizes, csr, lam = _oracle_build(oracle, x, maxk, radius, gp)

    def builder():
        return (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25)
                .with_synthesis(asb.TauMode.Median).with_seed(42).with_inline_sampling(None)
                .with_dims_reduction(False, None).with_cluster_params(maxk, radius))

    aspace, gl = builder().build(x)
    _assert_cluster_equal((gl.init_data, aspace.cluster_assignments, aspace.cluster_sizes), (cent, asg, sizes))
    _assert_csr_equal(gl.csr, csr)
    _assert_lambda_close(aspace.lambdas, lam)
    assert gl.shape() == (f, f) and gl.nnodes == n               # tests/test_builder.rs:272,339
    info = aspace.index_info()
    assert info.n_clusters == cent.shape[0] and info.nnz == csr[0][-1]
    assert np.isclose(info.lambda_min, lam.min(), rtol=1e-9) and np.isclose(info.lambda_sum, lam.sum(), rtol=1e-9)

    b = builder()
    out = asb.ArrowSpace.start_clustering(b, x)
    gl2 = out.aspace.eigenmaps(b, out.centroids, n)
    out.aspace.compute_taumode(gl2)
    assert np.array_equal(out.aspace.cluster_assignments, aspace.cluster_assignments)
    assert np.array_equal(out.aspace.cluster_sizes, aspace.cluster_sizes)
    _assert_csr_equal(gl2.csr, gl.csr)
    assert np.allclose(out.aspace.lambdas, aspace.lambdas, rtol=1e-12, atol=0)

    queries, src = asb.synth.queries_from_items(x, 50, seed=43)
    lq_want = oracle.compute_taumode(queries, csr, TAU_MEDIAN)
    want = oracle.search_lambda_aware_batch(x, lam, queries, lq_want, 10, 0.7)
    idx, score, count, lq = aspace.search_batch(queries, 10, 0.7)           # EigenMaps::search, batched
    _assert_lambda_close(lq, lq_want)
    _assert_topk_equal(idx, score, count, *want)
    one = out.aspace.search(queries[0], gl2, 10, 0.7)                       # EigenMaps::search, single
    assert [i for i, _ in one] == want[0][0].tolist()
    res = aspace.search_lambda_aware(asb.ArrowItem.new(queries[1], float(lq_want[1])), 10, 0.7)
    assert [i for i, _ in res] == want[0][1].tolist()
    # alpha = 1: the source item is the best hit (query = item x 1.02 has cosine 1)
    idx1, _, _, _ = aspace.search_batch(queries, 1, 1.0)
    assert (np.asarray(idx1)[:, 0] == src).mean() > 0.9


def test_spectral_signals_build(ctx, asb, oracle):
    """with_spectral(true) (src/builder.rs:157-162, src/graph.rs:211-231): signals = the Laplacian
    construction run on dense(L)^T; ITEM lambdas come from the signals graph (src/taumode.rs:195-200), QUERY lambdas
    keep using the feature Laplacian (src/core.rs:548).  One native call and the staged mirror must both match
    the oracle composition."""
    n, f, maxk = 4_000, 128, 80
    x = asb.synth.protein_like(n, f, seed=42)
    radius = 1.5 * f * 0.0025 * 2
    gp = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    cent, asg, sizes = oracle.cluster_incremental(x, maxk, radius)
    csr = oracle.feature_laplacian(cent, **gp)
    sig = oracle.spectral_signals(csr, **gp)
    lam = oracle.compute_taumode(x, sig, TAU_MEDIAN, 0.0)
    lam_plain = oracle.compute_taumode(x, csr, TAU_MEDIAN, 0.0)
    assert not np.allclose(lam, lam_plain)          # the two graphs really give different lambdas

    def builder():
        return (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25)
                .with_synthesis(asb.TauMode.Median).with_seed(42).with_inline_sampling(None)
                .with_dims_reduction(False, None).with_cluster_params(maxk, radius).with_spectral(True))

    aspace, gl = builder().build(x)
    _assert_csr_equal(gl.csr, csr)
    _assert_csr_equal(aspace.signals, sig)
    _assert_lambda_close(aspace.lambdas, lam)
    assert aspace.index_info().nnz_signals == sig[0][-1]

    b = builder()
    out = asb.ArrowSpace.start_clustering(b, x)
    gl2 = out.aspace.eigenmaps(b, out.centroids, n)
    out.aspace.compute_taumode(gl2)
    _assert_csr_equal(out.aspace.signals, sig)
    assert np.allclose(out.aspace.lambdas, aspace.lambdas, rtol=1e-12, atol=0)

    queries, _ = asb.synth.queries_from_items(x, 20, seed=43)
    lq_want = oracle.compute_taumode(queries, csr, TAU_MEDIAN)       # queries: feature Laplacian, not signals
    want = oracle.search_lambda_aware_batch(x, lam, queries, lq_want, 10, 0.7)
    idx, score, count, lq = aspace.search_batch(queries, 10, 0.7)
    _assert_lambda_close(lq, lq_want)
    _assert_topk_equal(idx, score, count, *want)

    plain, _ = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_seed(42)
                .with_inline_sampling(None).with_cluster_params(maxk, radius).build(x))
    assert plain.signals is None and plain.index_info().nnz_signals == 0


@pytest.mark.parametrize("n,f,k", [(20_000, 128, 10), (5_000, 33, 60), (300, 384, 7), (5, 16, 10)])
def test_search_energy_parity(ctx, asb, oracle, n, f, k):
    """EnergyMaps::search_energy (src/energymaps.rs:368-407, :838-895): (index, -energy) best
    first.  Queries include exact copies of items (distance 0, where the GEMM-form pre-ranking is weakest)."""
    x = asb.synth.protein_like(n, f, seed=9)
    rng = np.random.default_rng(3)
    lam = rng.uniform(0.05, 0.9, n)
    nq = 12
    queries = x[rng.integers(0, n, nq)] * rng.uniform(0.97, 1.03, (nq, 1))
    queries[0] = x[min(3, n - 1)]                       # exact duplicate of an item
    lq = rng.uniform(0.05, 0.9, nq)
    lq[0] = lam[min(3, n - 1)]
    idx, score, count = ctx.search_energy_batch(x, lam, queries, lq, k, 1.0, 0.5)
    for qi in range(nq):
        want = oracle.search_energy(x, lam, queries[qi], float(lq[qi]), k, 1.0, 0.5)
        assert int(count[qi]) == len(want) == min(k, n)
        wi = np.array([i for i, _ in want])
        ws = np.array([s for _, s in want])
        gi, gs = np.asarray(idx[qi, :len(want)]), np.asarray(score[qi, :len(want)])
        assert np.allclose(gs, ws, rtol=0, atol=1e-12)
        diff = gi != wi
        if diff.any():                                  # ids may only differ inside score gaps < 1e-9
            for r in np.nonzero(diff)[0]:
                assert abs(ws[r] - ws[max(r - 1, 0)]) < 1e-9 or abs(ws[r] - ws[min(r + 1, len(ws) - 1)]) < 1e-9
    assert idx[0, 0] == min(3, n - 1) and score[0, 0] == 0.0
    # other weights, and the ArrowSpace mirror on a built index
    idx2, score2, _ = ctx.search_energy_batch(x, lam, queries[:2], lq[:2], min(k, 5), 0.25, 2.0)
    want2 = oracle.search_energy(x, lam, queries[1], float(lq[1]), min(k, 5), 0.25, 2.0)
    assert [int(i) for i in idx2[1, :len(want2)]] == [i for i, _ in want2]


def test_search_energy_on_built_space(ctx, asb, oracle):
    x = asb.synth.protein_like(3_000, 64, seed=42)
    aspace, gl = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_seed(42)
                  .with_inline_sampling(None).with_cluster_params(40, 1.5 * 64 * 0.0025 * 2).build(x))
    q = x[17] * 1.01
    lq = oracle.compute_taumode(q.reshape(1, -1), gl.csr, TAU_MEDIAN)[0]
    want = oracle.search_energy(x, aspace.lambdas, q, float(lq), 8, 1.0, 0.5)
    got = aspace.search_energy(q, gl, 8, 1.0, 0.5)
    assert [i for i, _ in got] == [i for i, _ in want]
    assert np.allclose([s for _, s in got], [s for _, s in want], rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,f,r", [(3_001, 384, 91), (17, 24, 32), (1_000, 130, 300), (5, 7, 1)])
def test_project_matrix_bit_exact(ctx, asb, oracle, n, f, r):
    """JL projection with a materialised matrix (src/reduction.rs:143-199): same operations in the
    same order as the reference -> bit-identical, from host and from device buffers."""
    import torch
    rng = np.random.default_rng(11)
    x = asb.synth.protein_like(n, f, seed=3)
    g = rng.normal(size=(f, r))
    want = oracle.project_matrix(x, g)
    got = ctx.project_matrix(x, g)
    assert got.shape == (n, r)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    got_d = ctx.project_matrix(torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda())
    assert np.array_equal(got_d.cpu().numpy().view(np.uint64), want.view(np.uint64))
    proj = asb.host.ImplicitProjection(g, ctx)
    assert np.array_equal(proj.project(x[0]), want[0])
    assert np.array_equal(np.asarray(asb.host.project_matrix(x[:3], proj)), want[:3])


def test_builder_defaults_give_degenerate_graph(ctx, asb):
    """Literal builder defaults (eps=1e-3) on generic data: empty graph -> lambda == 0 -> the search
    panics in the reference (core.rs:773-776); here ASB_ERR_ZERO_LAMBDA."""
    x = asb.synth.protein_like(2_000, 64, seed=42)
    b = (asb.ArrowSpaceBuilder.new(ctx).with_seed(1).with_inline_sampling(None).with_cluster_params(32, 0.5))
    aspace, gl = b.build(x)
    assert gl.nnz() == 64 and np.all(aspace.lambdas == 0.0)
    with pytest.raises(asb.ArrowSpaceError) as ei:
        aspace.search(x[0] * 1.02, gl, 3, 0.7)
    assert ei.value.status == 5
    with pytest.raises(asb.ArrowSpaceError) as ei:               # core.rs:416-420
        asb.ArrowSpaceBuilder.new(ctx).with_inline_sampling(None).with_cluster_params(2, 1.0).build(x[:1])
    assert ei.value.status == 11


def test_device_resident_inputs(ctx, asb, oracle):
    """torch CUDA tensors are used in place (zero copy) and give the same answers as host arrays."""
    import torch
    x = asb.synth.protein_like(6_000, 128, seed=42)
    cent, _, _ = oracle.cluster_incremental(x[:2000], 50, 1.5 * 128 * 0.0025 * 2)
    csr = _graph(oracle, cent)
    xd = torch.from_numpy(x).cuda()
    lam_h, n2_h, _ = ctx.compute_taumode(x, csr, asb.TauMode.Median, want_norms=True)
    lam_d, n2_d, _ = ctx.compute_taumode(xd, csr, asb.TauMode.Median, want_norms=True)
    assert lam_d.is_cuda and np.array_equal(lam_d.cpu().numpy(), lam_h)
    queries, _ = asb.synth.queries_from_items(x, 64, seed=43)
    qd = torch.from_numpy(queries).cuda()
    lq_d = ctx.prepare_query_lambdas(qd, csr, asb.TauMode.Median)
    idx_d, sc_d, cnt_d = ctx.search_lambda_aware_batch(xd, lam_d, qd, lq_d, 10, 0.7, norms2=n2_d)
    idx_h, sc_h, cnt_h = ctx.search_lambda_aware_batch(x, lam_h, queries, lq_d.cpu().numpy(), 10, 0.7)
    # norms come from two kernels with different summation orders -> scores agree to an ulp or two
    assert np.array_equal(idx_d.cpu().numpy(), idx_h) and np.allclose(sc_d.cpu().numpy(), sc_h, rtol=0, atol=1e-14)


def test_cpp_host_mirror(asb):
    """include/arrowspace_b200.hpp (the compiled-language host mirror) end to end: build + search."""
    import subprocess
    exe = asb._build.build_cpp_example()
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "top: 3 (1.000000)" in out.stdout


# ===================================================== "next" rows: hybrid + range search
def test_hybrid_search_parity(ctx, asb, oracle, golden):
    """search_lambda_aware_hybrid (core.rs:802-928): union of {cos > 0.9999}, lambda top-k, semantic top-1."""
    db = golden["proteins"]
    csr = _graph(oracle, db[:20])
    lam = oracle.compute_taumode(db, csr, TAU_MEDIAN)
    rng = np.random.RandomState(4)
    near = db[:8] + 0.004 * rng.rand(8, 24)              # cos ~ 0.99995 > 0.9999, distinct from the originals
    dup = np.ascontiguousarray(np.vstack([db, near]))   # several cos > 0.9999 hits per query
    lam_d = np.concatenate([lam, lam[:8] + 0.3])
    aspace = asb.ArrowSpace(dup, asb.TauMode.Median, ctx)
    aspace.lambdas = lam_d
    for qi in (3, 10, 63):
        for k in (1, 3, 10, 40):
            for alpha in (0.9, 0.5):
                q = asb.ArrowItem.new(db[qi] * 1.02, float(lam[qi]))
                want = oracle.search_lambda_aware_hybrid(dup, lam_d, q.item, q.lambda_, k, alpha)
                got = aspace.search_lambda_aware_hybrid(q, k, alpha)
                assert [i for i, _ in got] == [i for i, _ in want], (qi, k, alpha)
                assert np.allclose([s for _, s in got], [s for _, s in want], rtol=0, atol=1e-12)
    assert aspace.search_lambda_aware_hybrid(asb.ArrowItem.new(db[0], 0.3), 0, 0.7) == []   # core.rs:810-812
    x = asb.synth.protein_like(5_000, 128, seed=42)
    lam5 = np.linspace(0.2, 0.6, 5_000)
    queries, _ = asb.synth.queries_from_items(x, 12, seed=43)
    idx, score, count = ctx.search_lambda_aware_hybrid_batch(x, lam5, queries, np.full(12, 0.4), 10, 0.7)
    for qn in range(12):
        want = oracle.search_lambda_aware_hybrid(x, lam5, queries[qn], 0.4, 10, 0.7)
        assert idx[qn, :count[qn]].tolist() == [i for i, _ in want]


def test_range_search_parity(ctx, asb, oracle, golden):
    """range_search (core.rs:944-976): signed lambda difference <= eps, index order, re-prepared zero lambda."""
    rng = np.random.RandomState(1)
    for n in (1, 31, 4096, 4097, 50_000):
        lam = rng.rand(n)
        for lq, eps in [(0.5, 0.1), (0.5, -0.2), (2.0, 0.5), (0.0, 10.0), (0.3, 0.0)]:
            widx, wdist = oracle.range_search(lam, lq, eps)
            gidx, gdist = ctx.range_search(lam, lq, eps)
            assert np.array_equal(gidx, widx) and np.array_equal(gdist, wdist), (n, lq, eps)
    db = golden["proteins"]
    csr = _graph(oracle, db[:20])
    aspace = asb.ArrowSpace(db, asb.TauMode.Median, ctx)
    gl = asb.GraphLaplacian(*csr, nnodes=64, graph_params=None)
    aspace.compute_taumode(gl)
    q = db[5] * 1.02
    lq = oracle.prepare_query_item(q, csr, TAU_MEDIAN)
    want = oracle.range_search(aspace.lambdas, lq, 0.01)
    got = aspace.range_search(asb.ArrowItem.new(q, 0.0), gl, 0.01)            # lambda 0 -> re-prepared (:953-957)
    assert [i for i, _ in got] == want[0].tolist()
    assert np.allclose([d for _, d in got], want[1], rtol=0, atol=1e-12)


# ============================================ SURVEY 8f rank 2 + 4: JL-projected build, every branch of the energy score
def _oracle_projected_build(oracle, x, maxk, radius, gp, proj):
    """The reference's with_dims_reduction flow (src/eigenmaps.rs:248-269, src/taumode.rs:233-245): centroids projected
    before the Laplacian (r x r graph); item lambdas read item[0 .. r) against it, tau and the denominator over all F."""
    cent, asg, sizes = oracle.cluster_incremental(x, maxk, radius)
    centp = oracle.project_matrix(cent, proj)
    csr = oracle.feature_laplacian(centp, **gp)
    r = proj.shape[1]
    lam = np.array([oracle.synthetic_lambda_prefix(row, csr, oracle.select_tau(row, TAU_MEDIAN), r) for row in x])
    return cent, asg, csr, lam


@pytest.mark.parametrize("spectral", [False, True])
def test_projected_build_and_energy_search(ctx, asb, oracle, spectral):
    n, f, maxk = 3_000, 96, 40
    x = asb.synth.protein_like(n, f, seed=42)
    radius = 1.5 * f * 0.0025 * 2
    gp = dict(eps=1.6, k=12, topk=4, p=2.0, sigma=0.5)           # projected centroids are signed: wide eps
    rng = np.random.default_rng(17)
    cent0, _, _ = oracle.cluster_incremental(x, maxk, radius)
    r = min(asb.host.compute_jl_dimension(cent0.shape[0], 0.5), f // 2)
    proj = rng.normal(size=(f, r))                                  # the host materialises the Gaussian matrix
    cent, asg, csr, lam = _oracle_projected_build(oracle, x, maxk, radius, gp, proj)
    b = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(gp["eps"], gp["k"], gp["topk"], gp["p"], gp["sigma"])
         .with_synthesis(asb.TauMode.Median).with_seed(42).with_inline_sampling(None)
         .with_dims_reduction(True, 0.5).with_projection(proj).with_cluster_params(maxk, radius).with_spectral(spectral))
    aspace, gl = b.build(x)
    assert aspace.reduced_dim == r and gl.indptr.shape[0] == r + 1
    assert np.array_equal(gl.init_data.view(np.uint64), cent.view(np.uint64))     # clustering is untouched by the projection
    _assert_csr_equal(gl.csr, csr)
    sig = oracle.spectral_signals(csr, **gp) if spectral else None
    if spectral:
        _assert_csr_equal(aspace.signals, sig)
        lam = np.array([oracle.synthetic_lambda_prefix(row, sig, oracle.select_tau(row, TAU_MEDIAN), r) for row in x])
    _assert_lambda_close(aspace.lambdas, lam)
    # prepare_query_item projects the query first (src/core.rs:540-545)
    queries, _ = asb.synth.queries_from_items(x, 9, seed=43)
    qp = oracle.project_matrix(queries, proj)
    lq_want = oracle.compute_taumode(qp, csr, TAU_MEDIAN)
    _assert_lambda_close(aspace.prepare_query_items_index(queries), lq_want)
    assert abs(aspace.prepare_query_item(queries[0], gl) - lq_want[0]) <= 1e-9 * abs(lq_want[0])
    # the lambda-aware search of a projected index panics in the reference (lengths differ, src/core.rs:157-161)
    with pytest.raises(asb.ArrowSpaceError) as ei:
        aspace.search_batch(queries, 5, 0.7)
    assert ei.value.status == 12
    # energy search: project_vec on both sides, projected_dirichlet through the signals when they exist
    idx, score, count = aspace.search_energy_batch(queries, 10, 1.0, 0.5)
    for qi in range(len(queries)):
        want = oracle.search_energy_ex(x, lam, queries[qi], float(lq_want[qi]), 10, 1.0, 0.5, projection=proj, signals=sig)
        assert [int(i) for i in idx[qi, :len(want)]] == [i for i, _ in want], qi
        assert np.allclose(score[qi, :len(want)], [s for _, s in want], rtol=0, atol=1e-12)
    one = aspace.search_energy(queries[2], gl, 7, 0.25, 2.0)
    want = oracle.search_energy_ex(x, lam, queries[2], float(lq_want[2]), 7, 0.25, 2.0, projection=proj, signals=sig)
    assert [i for i, _ in one] == [i for i, _ in want]


def test_energy_search_through_signals_without_projection(ctx, asb, oracle):
    """with_spectral(true) without a projection: projected_dirichlet runs through the F x F signals."""
    n, f, maxk = 4_000, 64, 40
    x = asb.synth.protein_like(n, f, seed=42)
    radius = 1.5 * f * 0.0025 * 2
    gp = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)
    aspace, gl = (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(0.5, 12, 4, 2.0, 0.25).with_seed(42)
                  .with_inline_sampling(None).with_cluster_params(maxk, radius).with_spectral(True).build(x))
    sig = aspace.signals
    queries, _ = asb.synth.queries_from_items(x, 6, seed=43)
    lq = oracle.compute_taumode(queries, gl.csr, TAU_MEDIAN)
    idx, score, count = aspace.search_energy_batch(queries, 12, 1.0, 0.5)
    for qi in range(len(queries)):
        want = oracle.search_energy_ex(x, aspace.lambdas, queries[qi], float(lq[qi]), 12, 1.0, 0.5, signals=sig)
        assert [int(i) for i in idx[qi, :len(want)]] == [i for i, _ in want], qi
        assert np.allclose(score[qi, :len(want)], [s for _, s in want], rtol=0, atol=1e-12)
