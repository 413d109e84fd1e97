"""Clustering kernel time at the centroid counts of the multi-GPU runs (K = 449 / 634 / 896), 200k rows."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import arrowspace_b200 as asb, torch
ctx = asb.Context(0)
n, f = 200_000, 384
x = asb.synth.protein_like(n, f, seed=42)
xd = torch.from_numpy(x).cuda()
radius = asb.heuristics.pilot_radius(x, 384, 128)
for k in (384, 634, 700, 896, 1200):
    for _ in range(2):
        ctx.cluster_incremental(xd, k, radius)
    print(f"K={k}: variant {ctx.kernel_ms('cluster_variant'):.0f} ncta {ctx.kernel_ms('cluster_ncta'):.0f} ring {ctx.kernel_ms('cluster_ring_groups'):.0f} "
          f"{ctx.kernel_ms('cluster_kernel'):.1f} ms / {n} rows = {1e3 * ctx.kernel_ms('cluster_kernel') / n:.3f} us/row", flush=True)
