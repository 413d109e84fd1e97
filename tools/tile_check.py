"""Worst |tensor-core distance - FP64| / certified bound over a clustering walk (cluster_check_tile option)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import arrowspace_b200 as asb

ctx = asb.Context(0)
for n, f, maxk in ((20_000, 384, 384), (6_000, 132, 64), (3_000, 33, 20), (4_000, 768, 200)):
    x = asb.synth.protein_like(n, f, seed=5)
    ctx.set_option("cluster_check_tile", 1)
    ctx.cluster_incremental(x, maxk, 1.5 * f * 0.0025 * 2)
    print(f"n={n} f={f} K={maxk}: variant {ctx.kernel_ms('cluster_variant'):.0f}, worst error / bound = "
          f"{ctx.kernel_ms('cluster_phase47') * 1e-12:.4f}, exact rows {ctx.kernel_ms('cluster_exact_rows'):.0f}", flush=True)
    ctx.set_option("cluster_check_tile", 0)
    ctx.set_option("cluster_phase_times", 0)
