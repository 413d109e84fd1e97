"""Workload for compute-sanitizer (memcheck / racecheck over K2, K4-K6, K8 and the prefilter).

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize.py [small|wide]
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py [small|wide]
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize.py [small|wide]

Two shapes: `small` (3000 x 64, K = 40) and `wide` (2500 x 384, K = 120: the pipelined clustering kernel's cluster of 16
CTAs, the tcgen05 tile with 12 K-chunks, the TMA-fed chain kernel).  Every clustering variant, the replay (with its
growth run and ring kernel), the Laplacian, taumode (symmetric + generic), both prefilter tiles (tcgen05 and mma.sync),
the exact search, Two-NN, hybrid / range / energy searches run once each; results are compared with each other (not
with the oracle: the sanitizer multiplies run time by 10-100x, the parity tests are the place for that).

Known and intended: racecheck only sees shared memory.  The one deliberate data race of the path -- the pipelined
clustering kernel reading FP32 centroid shadows that another CTA of the cluster is rewriting (torn reads, bounded by the
certified displacement: DESIGN.md K2) -- goes through distributed shared memory with st.async / mbarrier completion and
is reported by racecheck, if at all, as a hazard on `cshadow`; it is exempt by design.  Everything else must be clean."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import arrowspace_b200 as asb  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "small"
noseq = len(sys.argv) > 2 and sys.argv[2] == "noseq"
# second argument "noseq": skip the four hand-synchronised sequential clustering variants (racecheck needs minutes for
# them and reports their flag / mbarrier protocols as hazards, see profiles/r02_sanitizer.md) so that the tool reaches
# the kernels behind them; the reference walk then comes from the row-wise variant alone
n, f, maxk = (3000, 64, 40) if shape == "small" else (1500 if noseq else 2500, 384, 120)
ctx = asb.Context(0)
x = asb.synth.protein_like(n, f, seed=42)
radius = 1.5 * f * 0.0025 * 2
ref = None
for variant in ((2,) if noseq else (-2, -1, 0, 1, 2)):
    ctx.set_option("cluster_replay", 0)
    ctx.set_option("cluster_first_variant", variant)
    cent, asg, sizes = ctx.cluster_incremental(x, maxk, radius)
    used = ctx.kernel_ms("cluster_variant")
    if ref is None:
        ref = (cent.copy(), asg.copy())
    assert np.array_equal(cent.view(np.uint64), ref[0].view(np.uint64)) and np.array_equal(asg, ref[1]), variant
    print("cluster variant", variant, "->", used, "ok")
ctx.set_option("cluster_first_variant", 2 if noseq else -9)
ctx.set_option("cluster_replay", 1)
ctx.set_option("cluster_replay_prefix", 512)
ctx.set_option("cluster_replay_chunk", 512)
cent, asg, sizes = ctx.cluster_incremental(x, maxk, radius)
assert np.array_equal(cent.view(np.uint64), ref[0].view(np.uint64)) and np.array_equal(asg, ref[1])
print("replay ok: chunks", ctx.kernel_ms("cluster_replay_chunks"), "proven", ctx.kernel_ms("cluster_replay_chunks_ok"),
      "growth rows", ctx.kernel_ms("cluster_growth_rows"))
gp = asb.GraphParams(0.5, 12, 4, 2.0, 0.25)
csr = ctx.build_feature_laplacian(cent, gp)
csr_n = ctx.build_feature_laplacian(cent, asb.GraphParams(1.2, 12, 4, 2.0, 0.5, normalise=1))
print("laplacian ok: nnz", csr[0][-1], "normalised nnz", csr_n[0][-1])
lam, n2, st = ctx.compute_taumode(x, csr, asb.TauMode.Median, want_norms=True)
for opts in (dict(taumode_ipp=2), dict(taumode_regs=0)):
    for k_, v_ in opts.items():
        ctx.set_option(k_, v_)
    lam_v, _, _ = ctx.compute_taumode(x, csr, asb.TauMode.Median)
    assert np.allclose(lam, lam_v, rtol=1e-12)
ctx.set_option("taumode_ipp", 1)
ctx.set_option("taumode_regs", 1)
ctx.set_option("taumode_generic", 1)
lam_g, _, _ = ctx.compute_taumode(x, csr, asb.TauMode.Median)
ctx.set_option("taumode_generic", 0)
assert np.allclose(lam, lam_g, rtol=1e-12)
print("taumode ok")
q, _ = asb.synth.queries_from_items(x, 130, seed=43)
lq = ctx.prepare_query_lambdas(q, csr, asb.TauMode.Median)
res = {}
for name, opts in (("tcgen05", dict(search_prefilter=1, search_umma=1, search_umma_cluster=2)),
                   ("tcgen05_cl1_kc32", dict(search_prefilter=1, search_umma=1, search_umma_cluster=1, search_umma_kc=32)),
                   ("mma_sync", dict(search_prefilter=1, search_umma=0)), ("exact", dict(search_prefilter=0))):
    for k_, v_ in opts.items():
        ctx.set_option(k_, v_)
    idx, score, count = ctx.search_lambda_aware_batch(x, lam, q, lq, 10, 0.7, norms2=n2)
    res[name] = np.asarray(idx).copy()
    print("search", name, "ok (prefilter used:", ctx.kernel_ms("search_pf_used"), ")")
ctx.set_option("search_prefilter", 1)
ctx.set_option("search_umma", 1)
ctx.set_option("search_umma_kc", 16)
assert all(np.array_equal(v, res["exact"]) for v in res.values())
d1, d2 = ctx.twonn_distances(x, asb.heuristics.sample_indices(n, 200, 129))
ctx.set_option("twonn_prefilter", 1)
e1, e2 = ctx.twonn_distances(x, asb.heuristics.sample_indices(n, 200, 129))
ctx.set_option("twonn_prefilter", 0)
assert np.allclose(d1, e1, rtol=1e-9) and np.allclose(d2, e2, rtol=1e-9)
print("twonn ok")
ctx.search_lambda_aware_hybrid_batch(x, lam, q[:8], lq[:8], 5, 0.7, norms2=n2)
ctx.range_search(lam, float(lq[0]), 0.01)
ctx.search_energy_batch(x, lam, q[:8], lq[:8], 5, 1.0, 0.5)
print("extras ok")
print("sanitize workload finished:", shape)
