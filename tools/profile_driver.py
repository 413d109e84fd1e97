"""Runs ONE stage of the path at a moderate size so that `ncu` can capture its kernel quickly.
    python tools/profile_driver.py taumode|search|search_exact|cluster|twonn|laplacian [n] [f] [nq]
`search` runs the default path (certified TF32 prefilter + exact rescoring, search_pf_kernel), `search_exact` the
FP64 DMMA kernel (search_prefilter = 0)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import arrowspace_b200 as asb  # noqa: E402
import torch  # noqa: E402

stage = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
f = int(sys.argv[3]) if len(sys.argv) > 3 else 384
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
ctx = asb.Context(0)
for kv in filter(None, os.environ.get("ASB_OPTS", "").split(",")):      # e.g. ASB_OPTS=search_umma_kc=32,search_umma=0
    key, val = kv.split("=")
    ctx.set_option(key, float(val))
x = asb.synth.protein_like(n, f, seed=42)
xd = torch.from_numpy(x).cuda()
_, kmax = asb.heuristics.step1_bounds(1_000_000, f, f)
radius = asb.heuristics.pilot_radius(x, kmax, 128)
nc = min(n, 20_000)
cent, asg, sizes = ctx.cluster_incremental(xd[:nc], kmax, radius)
gp = asb.GraphParams(0.5, 12, 4, 2.0, 0.25)
csr = ctx.build_feature_laplacian(cent, gp)
print("clusters", cent.shape, "nnz", csr[0][-1], "cluster_ms", ctx.kernel_ms("cluster_kernel"))
if stage == "cluster":
    for _ in range(2):
        ctx.cluster_incremental(xd[:nc], kmax, radius)
    print("cluster_ms", ctx.kernel_ms("cluster_kernel"), "rows", nc)
elif stage == "laplacian":
    for _ in range(3):
        ctx.build_feature_laplacian(cent, gp)
elif stage == "taumode":
    for _ in range(3):
        lam, n2, st = ctx.compute_taumode(xd, csr, asb.TauMode.Median, want_norms=True)
    print("taumode_ms", ctx.kernel_ms("taumode_kernel"), "GB/s", n * (8 * f + 16) / ctx.kernel_ms("taumode_kernel") / 1e6)
elif stage in ("search", "search_exact"):
    lam, n2, st = ctx.compute_taumode(xd, csr, asb.TauMode.Median, want_norms=True)
    q = torch.from_numpy(asb.synth.rows_at(asb.synth.query_indices(n, nq, 43), f, 42) * 1.02).cuda()
    lq = ctx.prepare_query_lambdas(q, csr, asb.TauMode.Median)
    ctx.set_option("search_prefilter", 0 if stage == "search_exact" else 1)
    for _ in range(3):
        ctx.search_lambda_aware_batch(xd, lam, q, lq, 10, 0.7, norms2=n2)
    if ctx.kernel_ms("search_pf_used") == 1.0:
        ms = ctx.kernel_ms("search_pf_kernel")
        print("search_pf_ms", ms, "effective TFLOP/s", 2.0 * nq * n * f / ms / 1e9, "prep_ms", ctx.kernel_ms("search_pf_prep"),
              "finish_ms", ctx.kernel_ms("search_pf_finish"), "candidates/query", ctx.kernel_ms("search_pf_candidates") / nq,
              "rescored/query", ctx.kernel_ms("search_pf_rescored") / nq)
    else:
        ms = ctx.kernel_ms("search_kernel")
        print("search_ms", ms, "TFLOP/s", 2.0 * nq * n * f / ms / 1e9, "pf_flags", ctx.kernel_ms("search_pf_flags"))
elif stage == "twonn":
    si = torch.from_numpy(asb.heuristics.sample_indices(n, 500, 129)).cuda()
    for _ in range(3):
        ctx.twonn_distances(xd, si.cpu().numpy())
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.twonn_distances(xd, si.cpu().numpy())
    torch.cuda.synchronize()
    print("twonn_call_ms", (time.perf_counter() - t0) * 1e3, "twonn_kernel_ms", ctx.kernel_ms("twonn_kernel"), "l2_pf_kernel_ms",
          ctx.kernel_ms("l2_pf_kernel"), "pf_used", ctx.kernel_ms("twonn_pf_used"))
