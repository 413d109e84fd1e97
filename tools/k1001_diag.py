import sys
sys.path.insert(0, '/root/repo')
import arrowspace_b200 as asb, numpy as np
ctx = asb.Context(0)
n, f, maxk = 1500, 768, 1001
x = asb.synth.protein_like(n, f, seed=77)
radius = 1.5 * f * 0.0025 * 2
for opt in (None, "cluster_rowwise"):
    if opt: ctx.set_option(opt, 1)
    try:
        c, a, s = ctx.cluster_incremental(x, maxk, radius)
        print(opt, "ok variant", ctx.kernel_ms("cluster_variant"), c.shape, flush=True)
    except Exception as e:
        print(opt, "FAILED", e, flush=True)
        break
