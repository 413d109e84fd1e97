"""Why does a C4-shaped shard (1.25M x 768) search fall back to the exact kernel?  Device-drawn rows, prints the prefilter's
diagnostics.    python tools/c4_search_diag.py [n] [f] [nq]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import arrowspace_b200 as asb

n, f, nq = [int(v) for v in sys.argv[1:]] + [1_250_000, 768, 2048][len(sys.argv) - 1:]
ctx = asb.Context(0)
g = torch.Generator(device="cuda").manual_seed(42)
centres = torch.rand((64, f), dtype=torch.float64, device="cuda", generator=g)
lab = torch.randint(0, 64, (n,), device="cuda", generator=g)
xd = centres[lab]
xd += 0.05 * torch.randn((n, f), dtype=torch.float64, device="cuda", generator=g)
xd.clamp_(min=0.0)
_, kmax = asb.heuristics.step1_bounds(n, f, f)
radius = asb.heuristics.pilot_radius(xd[:50_000].cpu().numpy(), kmax, asb.heuristics.CLUSTERING_SEED)
cent, asg, sizes = ctx.cluster_incremental(xd[:100_000], kmax, radius)
csr = ctx.build_feature_laplacian(cent, asb.GraphParams(0.5, 12, 4, 2.0, 0.25))
lam, n2, st = ctx.compute_taumode(xd, csr, asb.TauMode.Median, want_norms=True)
q = (xd[torch.randint(0, n, (nq,), device="cuda", generator=g)] * 1.02).contiguous()
lq = ctx.prepare_query_lambdas(q, csr, asb.TauMode.Median)
for _ in range(2):
    ctx.search_lambda_aware_batch(xd, lam, q, lq, 10, 0.7, norms2=n2)
keys = ("search_pf_used", "search_pf_umma", "search_umma_bf16", "search_pf_flags", "search_pf_cap", "search_pf_slabs", "search_pf_band",
        "search_pf_candidates", "search_pf_rescored", "search_pf_kernel", "search_pf_finish", "search_kernel")
print({k: ctx.kernel_ms(k) for k in keys})
