"""Small clustering walks for compute-sanitizer (all kernel variants)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import arrowspace_b200 as asb

ctx = asb.Context(0)
for n, f, maxk in ((1_500, 128, 48), (700, 36, 9)):
    x = asb.synth.protein_like(n, f, seed=7)
    for first in (-2, -1, 0, 2):
        ctx.set_option("cluster_first_variant", first)
        c, a, s = ctx.cluster_incremental(x, maxk, 1.5 * f * 0.0025 * 2)
        print(n, f, maxk, "variant", ctx.kernel_ms("cluster_variant"), "clusters", c.shape[0], flush=True)
