import sys
sys.path.insert(0, '/root/repo')
import arrowspace_b200 as asb, torch, numpy as np
ctx = asb.Context(0, stream=torch.cuda.current_stream().cuda_stream or None)
n, f = 200_000, 384
x = asb.synth.protein_like(n, f, seed=42)
xd = torch.from_numpy(x).cuda()
_, kmax = asb.heuristics.step1_bounds(2_000_000, f, f)
radius = asb.heuristics.pilot_radius(x, kmax, 128)
cent = torch.zeros((kmax, f), dtype=torch.float64, device='cuda')
sizes = torch.zeros(kmax, dtype=torch.int64, device='cuda')
for it in range(2):
    cent.zero_(); sizes.zero_()
    k, asg = ctx.cluster_incremental_resume(xd, kmax, radius, cent, sizes, 0)
    print('resume: variant', ctx.kernel_ms('cluster_variant'), 'ms', ctx.kernel_ms('cluster_kernel'), 'k', k, flush=True)
c2, a2, s2 = ctx.cluster_incremental(xd, kmax, radius)
print('plain: variant', ctx.kernel_ms('cluster_variant'), 'ms', ctx.kernel_ms('cluster_kernel'))
