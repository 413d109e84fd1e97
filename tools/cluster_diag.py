import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import arrowspace_b200 as asb, torch, numpy as np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
f = 384
ctx = asb.Context(0)
import os
if os.environ.get('DIAG_TIMES', '1') == '1':
    ctx.set_option('cluster_phase_times', 1)
if os.environ.get('DIAG_NOF32', '0') == '1':
    ctx.set_option('cluster_no_f32', 1)
if os.environ.get('DIAG_TICK_TID'):
    ctx.set_option('cluster_tick_tid', int(os.environ['DIAG_TICK_TID']))
if os.environ.get('DIAG_NOPIPE', '0') == '1':
    ctx.set_option('cluster_no_pipeline', 1)
print('diag start', flush=True)
x = asb.synth.protein_like(n, f, seed=42)
xd = torch.from_numpy(x).cuda()
_, kmax = asb.heuristics.step1_bounds(1_000_000, f, f)
radius = asb.heuristics.pilot_radius(x, kmax, 128)
for upto in (n // 2, n):
    for it in range(2):
        ctx.cluster_incremental(xd[:upto], kmax, radius)
    print(f"variant={ctx.kernel_ms('cluster_variant'):.0f} rows={upto} ms={ctx.kernel_ms('cluster_kernel'):.2f} blocks={ctx.kernel_ms('cluster_blocks'):.0f} "
          f"rows/block={upto/max(ctx.kernel_ms('cluster_blocks'),1):.2f} exact={ctx.kernel_ms('cluster_exact_rows'):.0f} "
          f"us/row={1e3*ctx.kernel_ms('cluster_kernel')/upto:.3f} us/block={1e3*ctx.kernel_ms('cluster_kernel')/max(ctx.kernel_ms('cluster_blocks'),1):.2f}")
    ph = [ctx.kernel_ms(f"cluster_phase{k}") for k in range(7)]
    tot = sum(ph) or 1
    names = ["fetch+wait", "phase1", "phase2", "cluster.sync", "3a", "resolve", "apply"]
    if ctx.kernel_ms('cluster_variant') == -2:
        names = ["nonspec p1+p2", "w0 cluster wait", "w0 resolve", "w0 spec p1", "S1 barrier", "apply", "p2+arrive"]
    print("   cycles/block:", {n: round(v / max(ctx.kernel_ms('cluster_blocks'), 1)) for n, v in zip(names, ph)})
    if ctx.kernel_ms('cluster_variant') == -2:
        print("   last arriver at the block barrier (warp: count, mean cycles after warp 0):",
              {w: (int(ctx.kernel_ms(f"cluster_phase{8 + w}")) // 1000000, int(ctx.kernel_ms(f"cluster_phase{8 + w}")) % 1000000) for w in range(24) if ctx.kernel_ms(f"cluster_phase{8 + w}") > 0})
        nb_ = max(ctx.kernel_ms('cluster_blocks'), 1)
        print("   warp 0 detail cycles/block:", {k: round(ctx.kernel_ms(f"cluster_phase{i}") / nb_) for k, i in
              [("p2 argmin", 40), ("p2 fence", 41), ("p2 bulk issue", 42), ("res reduce16+bounds", 43), ("res fast cert", 44), ("res slow path", 45), ("res dec write", 46)]})
