"""Sequential clustering kernel vs certified parallel replay (option cluster_replay) on device-resident rows.

    python tools/replay_diag.py [n] [f] [prefix] [chunk] [tf32]  (defaults: 1_000_000 384 2048 1024 0)

Rows have the bench data's shape (drawn on the GPU), max_clusters / radius come from the bench's own rule.  Prints the wall
time of both paths (CUDA events around the C-ABI call), how many chunks were proven, and whether centroids,
assignments and sizes are identical.  A timing / agreement probe; tests/test_cluster_replay.py is the parity test."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import arrowspace_b200 as asb


def main():
    a = [int(v) for v in sys.argv[1:]]
    n, f, prefix, chunk, tf32 = a + [1_000_000, 384, 2_048, 1_024, 0][len(a):]
    ctx = asb.Context(0)
    # the bench data's shape (64 non-negative blobs + 0.05 noise) drawn on the GPU: synth.protein_like takes ~26 s of
    # host time at 1M x 384, which a gpurun call pays for in box minutes
    g = torch.Generator(device="cuda").manual_seed(42)
    centres = torch.rand((64, f), dtype=torch.float64, device="cuda", generator=g)
    lab = torch.randint(0, 64, (n,), device="cuda", generator=g)
    xd = centres[lab]
    xd += 0.05 * torch.randn((n, f), dtype=torch.float64, device="cuda", generator=g)
    xd.clamp_(min=0.0)
    _, kmax = asb.heuristics.step1_bounds(n, f, f)
    radius = asb.heuristics.pilot_radius(xd[: min(n, 50_000)].cpu().numpy(), kmax, asb.heuristics.CLUSTERING_SEED)
    out = {"n": n, "f": f, "max_clusters": int(kmax), "radius": radius, "prefix": prefix, "chunk": chunk, "tf32": tf32}
    res = {}
    for name, opt in (("sequential", 0), ("replay", 1)):
        ctx.set_option("cluster_replay", opt)
        ctx.set_option("cluster_replay_tf32", tf32 if opt else 0)   # nearest / runner-up pass on the certified TF32 ranking
        ctx.set_option("cluster_replay_prefix", prefix)
        ctx.set_option("cluster_replay_chunk", chunk)
        if opt:
            for kv in filter(None, os.environ.get("ASB_OPTS", "").split(",")):   # e.g. ASB_OPTS=cluster_first_variant=0
                key, val = kv.split("=")
                ctx.set_option(key, float(val))
        for _ in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cent, asg, sizes = ctx.cluster_incremental(xd, kmax, radius)
            e1.record()
            torch.cuda.synchronize()
        res[name] = (np.asarray(cent), asg.cpu().numpy() if hasattr(asg, "cpu") else np.asarray(asg), np.asarray(sizes))
        out[name] = {"call_ms": e0.elapsed_time(e1), "clusters": int(len(res[name][0])),
                     "chunks": ctx.kernel_ms("cluster_replay_chunks"), "chunks_ok": ctx.kernel_ms("cluster_replay_chunks_ok"),
                     "rows_replayed": ctx.kernel_ms("cluster_replay_rows"),
                     "sequential_ms": ctx.kernel_ms("cluster_replay_seq_ms") if opt else ctx.kernel_ms("cluster_kernel"),
                     "top2_ms": ctx.kernel_ms("cluster_replay_top2_ms"), "chain_ms": ctx.kernel_ms("cluster_replay_chain_ms"),
                     "near_retries": ctx.kernel_ms("cluster_replay_near_retries"), "growth_rows": ctx.kernel_ms("cluster_growth_rows"),
                     "chain_rows": {k: ctx.kernel_ms("cluster_chain_%s" % k) for k in ("rows_grouped", "rows_by_row", "exact_steps", "checkpoints")},
                     "probe": {k: ctx.kernel_ms("cluster_probe_%s" % k) for k in ("rows", "span_us", "longest_rows", "longest_us", "last_block_rows")},
                     "wall_ms": {k: round(ctx.kernel_ms("cluster_wall_%s_ms" % k), 3) for k in ("growth", "prefix", "prepare", "run", "fallback")}}
    s, r = res["sequential"], res["replay"]
    out["centroids_bit_identical"] = bool(s[0].shape == r[0].shape and np.array_equal(
        np.ascontiguousarray(s[0]).view(np.uint64), np.ascontiguousarray(r[0]).view(np.uint64)))
    out["assignments_equal"] = bool(np.array_equal(s[1], r[1]))
    out["sizes_equal"] = bool(np.array_equal(s[2], r[2]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
