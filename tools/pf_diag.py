"""Exact search kernel vs the certified prefilter path (option search_prefilter) on device-resident data.

    python tools/pf_diag.py [n] [f] [nq] [k]      (defaults: 1_000_000 384 10_000 10)

Data are generated on the GPU (64 blobs + noise, the shape of synth.protein_like, from torch's generator: this is a
timing / agreement probe, not a parity test -- tests/test_search_prefilter.py is).  Prints per-kernel device times,
the candidate volume and whether both paths return the same ids.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

import arrowspace_b200 as asb


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    ctx = asb.Context(0)
    g = torch.Generator(device="cuda").manual_seed(42)
    centres = torch.rand((64, f), dtype=torch.float64, device="cuda", generator=g)
    lab = torch.randint(0, 64, (n,), device="cuda", generator=g)
    x = centres[lab]
    x += 0.05 * torch.randn((n, f), dtype=torch.float64, device="cuda", generator=g)
    x.clamp_(min=0.0)
    lam = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 0.2 + 0.1
    pick = torch.randint(0, n, (nq,), device="cuda", generator=g)
    q = (x[pick] * 1.02).contiguous()
    lq = (lam[pick] + 0.01).contiguous()
    n2 = (x * x).sum(1)
    out = {"n": n, "f": f, "nq": nq, "k": k}
    res = {}
    for name, opt in (("exact", 0), ("prefilter", 1)):
        ctx.set_option("search_prefilter", opt)
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx, score, count = ctx.search_lambda_aware_batch(x, lam, q, lq, k, 0.7, norms2=n2)
            e1.record()
            torch.cuda.synchronize()
        res[name] = (idx.clone(), score.clone())
        d = {"call_ms": e0.elapsed_time(e1)}
        keys = ["search_kernel"] if not opt else ["search_pf_prep", "search_pf_kernel", "search_pf_finish", "search_pf_used",
                                                  "search_pf_flags", "search_pf_cap", "search_pf_slabs", "search_pf_band",
                                                  "search_pf_candidates", "search_pf_rescored"]
        for kk in keys:
            d[kk] = ctx.kernel_ms(kk)
        out[name] = d
    same = bool((res["exact"][0] == res["prefilter"][0]).all())
    out["ids_equal"] = same
    out["ids_mismatch_rows"] = int(((res["exact"][0] != res["prefilter"][0]).any(1)).sum())
    out["max_score_diff"] = float((res["exact"][1] - res["prefilter"][1]).abs().max())
    flops = 2.0 * n * nq * f
    out["exact_tflops"] = flops / (out["exact"]["search_kernel"] * 1e-3) / 1e12 if out["exact"]["search_kernel"] else None
    if out["prefilter"]["search_pf_kernel"]:
        out["prefilter_effective_tflops"] = flops / (out["prefilter"]["search_pf_kernel"] * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
