"""Pins the CPU oracle against every known-answer test / fixture the reference's own tests hold
for this path.  CPU only."""
import math

import numpy as np
import pytest

from oracle_binding import TAU_FIXED, TAU_MEAN, TAU_MEDIAN, TAU_PERCENTILE, OracleError

FLOOR = 1e-10
NAN, INF = float("nan"), float("inf")


# ---- select_tau table: src/tests/test_taumode.rs:14-159 --------------------------------------
def test_select_tau_fixed(oracle):
    e = [0.1, 0.5, 1.0]
    assert oracle.select_tau(e, TAU_FIXED, 0.3) == 0.3
    for bad in (-0.1, 0.0, NAN, INF):
        assert oracle.select_tau(e, TAU_FIXED, bad) == FLOOR


def test_select_tau_mean(oracle):
    assert abs(oracle.select_tau([1.0, 2.0, 3.0], TAU_MEAN) - 2.0) < 1e-12
    assert abs(oracle.select_tau([1.0, NAN, 3.0, INF, 2.0], TAU_MEAN) - 2.0) < 1e-12
    assert oracle.select_tau([NAN, INF, -INF], TAU_MEAN) == FLOOR
    assert oracle.select_tau([], TAU_MEAN) == FLOOR


def test_select_tau_median(oracle):
    assert oracle.select_tau([3.0, 1.0, 2.0], TAU_MEDIAN) == 2.0
    assert abs(oracle.select_tau([1.0, 2.0, 3.0, 4.0], TAU_MEDIAN) - 2.5) < 1e-12
    assert oracle.select_tau([5.0], TAU_MEDIAN) == 5.0
    assert oracle.select_tau([NAN, 1.0, 3.0, INF, 2.0], TAU_MEDIAN) == 2.0
    assert oracle.select_tau([NAN, INF], TAU_MEDIAN) == FLOOR
    assert oracle.select_tau([], TAU_MEDIAN) == FLOOR


def test_select_tau_percentile(oracle):
    e = [1.0, 2.0, 3.0, 4.0, 5.0]
    assert oracle.select_tau(e, TAU_PERCENTILE, 0.0) == 1.0
    assert oracle.select_tau(e, TAU_PERCENTILE, 1.0) == 5.0
    assert oracle.select_tau(e, TAU_PERCENTILE, 0.5) == 3.0
    assert oracle.select_tau(e, TAU_PERCENTILE, -0.1) == 1.0
    assert oracle.select_tau(e, TAU_PERCENTILE, 1.5) == 5.0
    assert oracle.select_tau([], TAU_PERCENTILE, 0.5) == FLOOR


def test_select_tau_floor(oracle):
    assert oracle.select_tau([FLOOR * 2.0], TAU_MEAN) == FLOOR * 2.0
    assert oracle.select_tau([FLOOR / 2.0], TAU_MEAN) == FLOOR
    assert oracle.select_tau([0.0], TAU_MEAN) == FLOOR


# ---- nearest_centroid: src/tests/test_clustering.rs:22-59 ------------------------------------
def test_nearest_centroid(oracle):
    i, d2 = oracle.nearest_centroid([1.1, 2.1], [[1.0, 2.0], [5.0, 6.0], [9.0, 10.0]])
    assert i == 0 and d2 < 0.03
    i, _ = oracle.nearest_centroid([4.9, 5.1], [[0.0, 0.0], [5.0, 5.0], [10.0, 10.0]])
    assert i == 1
    # first strict minimum wins (clustering.rs:922)
    i, _ = oracle.nearest_centroid([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    assert i == 0


# ---- synthetic lambda closed form + scale invariance: src/tests/test_taumode.rs:499-528 -------
def _dense_to_csr(m):
    m = np.asarray(m, dtype=np.float64)
    ip, ii, dd = [0], [], []
    for r in range(m.shape[0]):
        for c in range(m.shape[1]):
            if m[r, c] != 0.0:
                ii.append(c)
                dd.append(m[r, c])
        ip.append(len(ii))
    return np.array(ip, dtype=np.int64), np.array(ii, dtype=np.int64), np.array(dd, dtype=np.float64)


def test_synthetic_lambda_closed_form_and_scale_invariance(oracle):
    csr = _dense_to_csr([[1.0, 0.5], [0.5, 1.0]])
    l1 = oracle.synthetic_lambda([1.0, 2.0], csr, 0.5)
    l2 = oracle.synthetic_lambda([2.0, 4.0], csr, 0.5)
    # x=[1,2]: num 7, den 5, E=1.4, no negative off-diagonals -> G=0 -> 0.5*1.4/1.9
    assert abs(l1 - 0.5 * 1.4 / 1.9) < 1e-15
    assert abs(l1 - l2) <= 1e-10 * max(abs(l1), abs(l2))


def test_synthetic_lambda_with_dispersion(oracle):
    # path graph 0-1-2, unit weights: L = [[1,-1,0],[-1,2,-1],[0,-1,1]], x = [1,2,4]
    csr = _dense_to_csr([[1.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 1.0]])
    x = [1.0, 2.0, 4.0]
    tau = 0.25
    num, den = 1.0 + 4.0, 21.0           # x^T L x = (1-2)^2 + (2-4)^2
    edge = 2.0 * (1.0 + 4.0)             # both directions
    g = 2.0 * ((1.0 / edge) ** 2 + (4.0 / edge) ** 2)
    e = num / den
    want = tau * e / (e + tau) + (1 - tau) * g
    assert abs(oracle.synthetic_lambda(x, csr, tau) - want) < 1e-15


# ---- published search answer on the 64x24 table: paper.md:123-133 -----------------------------
def test_paper_top3_alpha1(oracle, golden):
    db = golden["proteins"]
    q = db[3] * 1.02
    res = oracle.search_lambda_aware(db, np.full(64, 0.25), q, 0.3, 3, 1.0)
    assert [i for i, _ in res] == [3, 6, 0]
    for (_, s), want in zip(res, (1.000000, 0.999573, 0.999325)):
        assert abs(s - want) < 5e-7


def test_search_contract(oracle, golden):
    db = golden["proteins"]
    lam = np.linspace(0.1, 0.9, 64)
    with pytest.raises(OracleError) as ei:  # core.rs:773-776
        oracle.search_lambda_aware(db, lam, db[0], 0.0, 3, 0.7)
    assert ei.value.status == 5
    res = oracle.search_lambda_aware(db, lam, db[5], 0.4, 100, 0.7)   # k >= N -> all N (core.rs:786)
    assert len(res) == 64
    scores = [s for _, s in res]
    assert all(scores[i] >= scores[i + 1] for i in range(63))
    # ties -> lower index (stable sort): duplicate rows with identical lambdas
    dup = np.vstack([db[:4], db[:4]])
    res = oracle.search_lambda_aware(dup, np.full(8, 0.5), db[2], 0.5, 8, 0.7)
    assert [i for i, _ in res][:2] == [2, 6]


# ---- Laplacian invariants: src/tests/test_laplacian.rs:51-152, test_graph_factory.rs:34-98 ---
def _lap_invariants(ip, ii, dd, f):
    assert len(ip) == f + 1 and ip[0] == 0 and ip[-1] == len(ii)
    dense = np.zeros((f, f))
    for r in range(f):
        cols = ii[ip[r]:ip[r + 1]]
        assert np.all(np.diff(cols) > 0), "columns sorted, unique"
        assert r in cols, "diagonal explicitly stored"
        dense[r, cols] = dd[ip[r]:ip[r + 1]]
    assert np.allclose(dense, dense.T, atol=0, rtol=0), "symmetric"
    assert np.all(np.abs(dense.sum(1)) < 1e-12), "row sums ~ 0"
    assert np.all(np.diag(dense) >= 0)
    off = dense - np.diag(np.diag(dense))
    assert np.all(off <= 0)
    return dense


def test_laplacian_invariants_and_shape(oracle, golden):
    cent = golden["proteins"][:20]                      # X=20 centroids x F=24 features
    ip, ii, dd = oracle.feature_laplacian(cent, eps=0.5, k=6, topk=3, p=2.0, sigma=0.25)
    dense = _lap_invariants(ip, ii, dd, 24)             # F x F (tests/test_builder.rs:272,339)
    assert len(ii) <= 24 * (1 + 2 * 4)
    assert (dense != 0).sum() > 24                      # non-degenerate on protein-like data


def test_laplacian_brute_force_adjacency(oracle, golden):
    """kNN semantics of tests/test_helpers.rs:104-201 (rectified cosine, (distance, index) order)."""
    cent = golden["proteins"][:16]
    feats = cent.T                                      # F x X
    f = feats.shape[0]
    topk, eps, sigma, p = 3, 0.5, 0.25, 2.0
    ip, ii, dd = oracle.feature_laplacian(cent, eps=eps, k=6, topk=topk, p=p, sigma=sigma, rectified=True)
    norms = np.sqrt((feats * feats).sum(1))
    adj = np.zeros((f, f), dtype=bool)
    for i in range(f):
        cand = []
        for j in range(f):
            if i == j:
                continue
            cs = float(feats[i] @ feats[j]) / (norms[i] * norms[j])
            d = 1.0 - max(cs, 0.0)
            cand.append((d, j))
        cand.sort()
        for d, j in cand[: topk + 1]:                   # the reference asks for topk+1 (laplacian.rs:211)
            if d <= eps:
                adj[i, j] = adj[j, i] = True
    got = np.zeros((f, f), dtype=bool)
    for r in range(f):
        for c in ii[ip[r]:ip[r + 1]]:
            if c != r:
                got[r, c] = True
    assert np.array_equal(adj, got)


def test_laplacian_errors(oracle, golden):
    with pytest.raises(OracleError) as ei:              # laplacian.rs:129-134
        oracle.feature_laplacian(golden["proteins"][:1], 0.5, 6, 3, 2.0, None)
    assert ei.value.status == 6
    with pytest.raises(OracleError) as ei:              # graph.rs:185-193
        oracle.feature_laplacian(golden["quora"], 1e-9, 6, 3, 2.0, None, sparsity_check=True)
    assert ei.value.status == 7


def test_empty_graph_gives_zero_lambda(oracle, golden):
    """Builder defaults (eps=1e-3) on generic data: empty graph -> lambda == 0 -> search panics."""
    cent = golden["quora"]
    ip, ii, dd = oracle.feature_laplacian(cent, 1e-3, 6, 4, 2.0, None)
    assert len(ii) == 384 and np.all(dd == 0.0)         # only the always-stored diagonal
    lam = oracle.compute_taumode(cent, (ip, ii, dd), TAU_MEDIAN)
    assert np.all(lam == 0.0)


# ---- clustering: B2 semantics -----------------------------------------------------------------
def test_cluster_incremental_small(oracle):
    rows = np.array([[0.0, 0.0], [0.1, 0.0], [5.0, 5.0], [5.1, 5.0], [0.05, 0.0], [20.0, 20.0], [2.6, 2.6]])
    cent, asg, sizes = oracle.cluster_incremental(rows, 2, 1.0)
    # row0 -> c0 ; row1 d2=.01<=.5 -> assign c0 (mean .05) ; row2 far -> c1 ; row3 -> c1 ;
    # row4 -> c0 ; row5: saturated, far -> dropped ; row6: d2 to c1 ~ 12 -> dropped
    assert asg.tolist() == [0, 0, 1, 1, 0, -1, -1]
    assert sizes.tolist() == [3, 2]
    assert abs(cent[0, 0] - 0.05) < 1e-15 and abs(cent[1, 0] - 5.05) < 1e-15
    # soft assignment: radius < d2 <= 1.5 radius counts but does not move the centroid
    rows = np.array([[0.0, 0.0], [1.1, 0.0]])
    cent, asg, sizes = oracle.cluster_incremental(rows, 1, 1.0)
    assert asg.tolist() == [0, 0] and sizes.tolist() == [2] and cent[0, 0] == 0.0


def test_twonn_and_bounds(oracle):
    rows = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 3.0], [5.0, 5.0]])
    d1, d2 = oracle.twonn_distances(rows, [0, 3])
    assert d1[0] == 1.0 and d2[0] == 3.0
    assert abs(d1[1] - math.sqrt(29.0)) < 1e-15
    # step1_bounds (clustering.rs:85-97) at the benchmark shapes with id_est = F
    assert oracle.step1_bounds(10_000, 128, 128)[1] == 100
    assert oracle.step1_bounds(100_000, 384, 384)[1] == 316
    assert oracle.step1_bounds(1_000_000, 384, 384)[1] == 384
    # intrinsic dim: ratios all 2 -> 1/ln 2 = 1.44 -> 1
    assert oracle.intrinsic_dim(100, 50, [1.0, 1.0], [2.0, 2.0]) == 1
    assert oracle.intrinsic_dim(5, 50, [1.0], [2.0]) == 2


def test_host_heuristics_match_oracle(oracle, asb):
    h = asb.heuristics
    for n, f, idv in [(10_000, 128, 128), (1_000_000, 384, 384), (500, 24, 3), (50, 8, 8)]:
        assert h.step1_bounds(n, f, idv) == oracle.step1_bounds(n, f, idv)
    rng = np.random.RandomState(0)
    d1 = rng.rand(50) + 0.5
    d2 = d1 * (1.0 + rng.rand(50))
    assert h.intrinsic_dim_from_distances(1000, 64, d1, d2) == oracle.intrinsic_dim(1000, 64, d1, d2)


# ---- "next" rows: hybrid / range search semantics (src/core.rs:802-976) -----------------------------
def test_hybrid_and_range_semantics(oracle, golden):
    db = golden["proteins"]
    lam = np.linspace(0.1, 0.9, 64)
    q = db[3] * 1.02
    # alpha = 1: the lambda top-k is the cosine top-k, so hybrid == plain search (paper answer again)
    res = oracle.search_lambda_aware_hybrid(db, lam, q, 0.3, 3, 1.0)
    assert [i for i, _ in res] == [3, 6, 0]
    # an item with cos > 0.9999 is kept with its COSINE as score even when its lambda is far away
    far = lam.copy()
    far[3] = 5.0
    res = oracle.search_lambda_aware_hybrid(db, far, q, 0.3, 2, 0.1)
    assert res[0][0] == 3 and abs(res[0][1] - 1.0) < 1e-12
    assert oracle.search_lambda_aware_hybrid(db, lam, q, 0.3, 0, 0.7) == []
    idx, dist = oracle.range_search(np.array([0.1, 0.5, 0.9]), 0.5, 0.1)
    assert idx.tolist() == [1, 2] and np.allclose(dist, [0.0, -0.4])        # signed difference, one-sided


def test_spectral_signals_is_the_laplacian_of_the_dense_laplacian(oracle, golden):
    """SURVEY 8f rank 3 (src/graph.rs:211-231): signals = build_laplacian_matrix(dense(L)^T, params).  The oracle
    entry point must equal the composition spelled out by hand, and keep the Laplacian invariants
    (src/tests/test_laplacian.rs:51-152: symmetric, diagonal stored, zero row sums)."""
    x = golden["proteins"]
    gp = dict(eps=0.5, k=6, topk=3, p=2.0, sigma=0.25)
    csr = oracle.feature_laplacian(x, **gp)
    f = len(csr[0]) - 1
    dense = np.zeros((f, f))
    for r in range(f):
        for e in range(csr[0][r], csr[0][r + 1]):
            dense[r, csr[1][e]] = csr[2][e]
    want = oracle.feature_laplacian(dense, **gp)
    got = oracle.spectral_signals(csr, **gp)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    sig = np.zeros((f, f))
    for r in range(f):
        cols = got[1][got[0][r]:got[0][r + 1]]
        assert r in cols                                   # diagonal always stored
        sig[r, cols] = got[2][got[0][r]:got[0][r + 1]]
    assert np.allclose(sig, sig.T, atol=1e-12)
    assert np.allclose(sig.sum(axis=1), 0.0, atol=1e-9)


def test_search_energy_semantics(oracle):
    """SURVEY 8f rank 4 (src/energymaps.rs:368-407, :838-895): energy = w_lambda |dlambda| + w_D min(d/(1+d), 1),
    results are (index, -energy) best first, ties to the lower index, truncated to k."""
    items = np.array([[0.0, 0.0], [3.0, 4.0], [0.0, 0.0], [1.0, 0.0]])
    lambdas = np.array([0.5, 0.5, 0.5, 0.1])
    q = np.array([0.0, 0.0])
    got = oracle.search_energy(items, lambdas, q, 0.5, 3, 1.0, 0.5)
    # item 0 and 2: energy 0 (tie -> 0 first); item 1: 0.5 * 5/6; item 3: 0.4 + 0.5 * 1/2
    assert [i for i, _ in got] == [0, 2, 1]
    assert got[0][1] == 0.0 and got[1][1] == 0.0
    assert np.isclose(got[2][1], -0.5 * 5.0 / 6.0, rtol=0, atol=1e-15)
    full = oracle.search_energy(items, lambdas, q, 0.5, 10, 1.0, 0.5)
    assert len(full) == 4 and np.isclose(full[3][1], -(0.4 + 0.25), atol=1e-15)
    assert oracle.search_energy(items, lambdas, q, 0.5, 0, 1.0, 0.5) == []


def test_jl_dimension_and_projection(oracle, asb):
    """SURVEY 8f rank 2 (src/reduction.rs:127-199; src/tests/test_reduction.rs: dimension formula, shape, scale)."""
    import math
    assert oracle.jl_dimension(1000, 0.1) == math.ceil(8 * math.log(1000) / 0.01)   # 5527
    assert oracle.jl_dimension(10, 0.9) == 32                                         # floor of 32
    assert asb.host.compute_jl_dimension(1000, 0.1) == oracle.jl_dimension(1000, 0.1)
    assert asb.host.compute_jl_dimension(100, 0.5) == oracle.jl_dimension(100, 0.5)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(7, 20))
    g = rng.normal(size=(20, 5))
    y = oracle.project_matrix(x, g)
    assert y.shape == (7, 5)
    assert np.allclose(y, x @ g / math.sqrt(5), rtol=1e-13, atol=1e-13)
    # the accumulation order is the reference's: features ascending, (x * g) * scale per term
    acc = 0.0
    for j in range(20):
        acc += x[2, j] * g[j, 3] * (1.0 / math.sqrt(5))
    assert y[2, 3] == acc
