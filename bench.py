#!/usr/bin/env python
"""bench.py -- lambda-tau build items/s + lambda-aware search QPS@k=10 on B200 (BASELINE.json).

A "step" is one pass of the hot path over one batch of synthetic input:
    Two-NN scan (K1) -> incremental clustering (K2) -> feature Laplacian (K3+K4) -> taumode (K5+K6)
    = ArrowSpaceBuilder::build, then prepare_query (K7) + search_lambda_aware k=10 (K8) for Q queries.
Workload at N=1: BASELINE configs[2] "1M x 384 lambda-tau build + 10k-query lambda-aware search k=10
on 1 B200" (the configuration the metric is quoted on; 3.07 GB, fits one GPU).  At N>1 every rank
holds its own 1M x 384 shard of one virtual dataset (weak scaling, row sharding).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

`value` = items/s of the whole build with inputs resident in HBM; `e2e` = the same through the
public host-buffer API (pinned host rows -> H2D inside the timed region, results read back).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GRAPH = dict(eps=0.5, k=12, topk=4, p=2.0, sigma=0.25)   # with_lambda_graph(0.5, 12, 4, 2.0, Some(0.25))
ALPHA, TOPK = 0.7, 10
DATA_SEED, QUERY_SEED = 42, 43


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--items", dest="n", type=int, default=1_000_000, help="items per GPU")
    ap.add_argument("--features", dest="f", type=int, default=384)
    ap.add_argument("--queries", dest="nq", type=int, default=10_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cluster-replay", action="store_true",
                    help="walk every row on the sequential clustering kernel (option cluster_replay = 0; the default is "
                         "the certified parallel replay, same bits)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --items is the GLOBAL row count, split evenly over the GPUs (BASELINE configs[3]: "
                         "--items 10000000 --features 768 --strong --device-data)")
    ap.add_argument("--device-data", action="store_true",
                    help="draw the synthetic rows on the GPU (same blob model, torch generator) instead of on the host: for "
                         "configurations whose host generation would take minutes; implies --no-e2e --no-cpu-baseline")
    ap.add_argument("--no-unfriendly", action="store_true",
                    help="skip the second, unfriendly clustering workload (512 blobs on K = 128, reported beside the main line)")
    ap.add_argument("--ref-full", action="store_true",
                    help="with --impl reference: run the WHOLE workload once on the host cores instead of a bounded sample")
    ap.add_argument("--exact-search", action="store_true",
                    help="search with the exact FP64 DMMA kernel only (option search_prefilter = 0)")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bf16_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["bf16_tflops"]), ("dense bf16 measured on this pool (MEASURED_PEAKS.json: %.0f TFLOP/s burst, %.0f sustained)"
                                         % (d["bf16_tflops"], d.get("bf16_tflops_sustained", float("nan"))))
    return 2250.0, "fallback: nominal dense bf16 (B200_PROFILING.md)"


def umma_tf32_peak():
    p = ROOT / "profiles" / "r02_umma_tf32_peak.json"
    if p.exists():
        try:
            d = json.loads(p.read_text().splitlines()[0])
            return float(d["tflops"]), "tcgen05 TF32 micro-benchmark on this pool (tools/umma_peak.cu, profiles/r02_umma_tf32_peak.json)"
        except Exception:
            pass
    return 1116.3, "tcgen05 TF32 micro-benchmark on this pool (tools/umma_peak.cu): 1116 TFLOP/s"


def ncu_traffic(kernel_substr, n, f, nq):
    """DRAM bytes per launch of a kernel from the committed ncu launch list of THIS workload (profiles/r02_traffic.json,
    written by tools/ncu_launches.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum ... bench.py`); None
    when the capture is of another configuration."""
    p = ROOT / "profiles" / "r02_traffic.json"
    if not p.exists():
        return None
    try:
        d = json.loads(p.read_text())
        cfg = d.get("_config", {})
        if (cfg.get("n"), cfg.get("f"), cfg.get("nq")) != (n, f, nq):
            return None
        for name, v in d.items():
            if name != "_config" and kernel_substr in name:
                return float(v["dram_bytes_per_launch"])
    except Exception:
        return None
    return None


def fp64_peak():
    p = ROOT / "profiles" / "r01_fp64_peak.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["dmma_tflops"]), float(d["dfma_tflops"])
    return 34.1, 36.7


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/asb_clocks_{os.getpid()}.csv"

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def cluster_inputs(n_global: int, f: int, sample_rows: np.ndarray):
    """(max_clusters, radius): max_clusters = step1_bounds k_max with id_est = F;
    radius = the seeded pilot rule (heuristics.pilot_radius), both printed with the result."""
    import arrowspace_b200 as asb
    _, k_max = asb.heuristics.step1_bounds(n_global, f, f)
    radius = asb.heuristics.pilot_radius(sample_rows, k_max, asb.heuristics.CLUSTERING_SEED)
    return int(k_max), float(radius)


# ----------------------------------------------------------------------------- reference arm
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_times(o, x, queries, maxk, radius, n_build, n_search, cores):
    """One pass of the oracle port over a sample: per-stage wall times (s).  Every stage of the reference path is
    linear in the row count except the feature Laplacian (fixed: F nodes), so each stage is timed on its own and the
    caller scales stage by stage instead of scaling the sum."""
    from oracle_binding import TAU_MEDIAN
    t = {}
    t0 = time.perf_counter()
    o.twonn_distances(x[:n_build], asb_heur().sample_indices(n_build, 500, 129))
    t["twonn"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cent, asg, sizes = o.cluster_incremental(x[:n_build], maxk, radius)      # sequential in the deterministic mode
    t["cluster"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    csr = o.feature_laplacian(cent, **GRAPH)
    t["laplacian"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    lam_b = o.compute_taumode(x[:n_build], csr, TAU_MEDIAN)
    t["taumode"] = time.perf_counter() - t0
    lam = lam_b if n_search == n_build else o.compute_taumode(x[:n_search], csr, TAU_MEDIAN)
    lq = o.compute_taumode(queries, csr, TAU_MEDIAN)
    t0 = time.perf_counter()
    o.search_lambda_aware_batch(x[:n_search], lam, queries, lq, TOPK, ALPHA)
    t["search"] = time.perf_counter() - t0
    t["n_clusters"] = int(len(cent))
    return t


def asb_heur():
    import arrowspace_b200 as asb
    return asb.heuristics


def cpu_port_model(t, n_build, n_search, nq_sample, n_full):
    """Full-size figures from the sample's stage times: row-linear stages x (n_full / sample rows), Laplacian as is."""
    sb = n_full / n_build
    build_s = (t["twonn"] + t["cluster"] + t["taumode"]) * sb + t["laplacian"]
    qps = (nq_sample / t["search"]) * (n_search / n_full)
    return build_s, n_full / build_s, qps


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: the Rust crate cannot be compiled here --
    DESIGN.md) on the box's host cores, each step a bounded sample of the workload; `--ref-full` runs the whole
    workload once instead (minutes; the record that checks the sample's extrapolation is kept under profiles/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import arrowspace_b200 as asb
    from oracle_binding import Oracle
    o = Oracle()
    # a launcher may have exported OMP_NUM_THREADS=1 (torch.distributed.run does): the baseline uses every host core
    cores = o.set_num_threads(host_threads())
    n, f = args.n, args.f
    full = bool(args.ref_full)
    n_s = n if full else min(n, 20_000)            # rows of the build sample
    n_items_search = n if full else min(n, 200_000)
    q_s = max(cores, 16) * (4 if full else 1)
    x = asb.synth.protein_like(n_items_search, f, seed=DATA_SEED)
    maxk, radius = cluster_inputs(n * args.gpus, f, x[: min(len(x), 50_000)])
    queries = asb.synth.rows_at(asb.synth.query_indices(n_items_search, q_s, QUERY_SEED), f, DATA_SEED) * 1.02
    steps, warm = (1, 0) if full else (args.steps, args.warmup)
    acc = []
    for step in range(warm + steps):
        t = cpu_port_times(o, x, queries, maxk, radius, n_s, n_items_search, cores)
        if step >= warm:
            acc.append(t)
    t = {k: float(np.mean([a[k] for a in acc])) for k in acc[0]}
    build_s, items_s, qps = cpu_port_model(t, n_s, n_items_search, q_s, n)
    sample = (f"build: {'all' if full else 'first'} {n_s} rows of the {n}x{f} workload, stage by stage (Two-NN 500 samples, "
              f"sequential clustering, taumode scaled by {n}/{n_s}; feature Laplacian unscaled); search: {q_s} queries x "
              f"{n_items_search} items, QPS scaled by {n_items_search}/{n}")
    line = {
        "impl": "reference", "metric": "lambda_tau_build_items_per_s", "value": items_s, "unit": "items/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": (build_s + args.nq / qps) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}x{f} lambda-tau build + {args.nq}-query lambda-aware search k={TOPK}"
                               + ("" if full else " (bounded sample, stage-wise linear model)"),
                   "max_clusters": maxk, "radius": radius, "graph": GRAPH, "taumode": "Median", "alpha": ALPHA},
        "search_qps": qps, "same_config": full, "extrapolated": not full,
        "cpu_baseline": {"value": items_s, "unit": "items/s", "cores": cores, "kind": "port", "sample": sample,
                         "search_qps": qps, "stage_s_on_sample": t, "omp_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "full_size_check": full_size_record()},
        "e2e": {"value": items_s, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def full_size_record():
    """The one full-size run of the oracle port kept under profiles/ (`bench.py --impl reference --ref-full`): what the
    bounded sample's stage-wise model is checked against."""
    p = ROOT / "profiles" / "r02_reference_full_1m.json"
    if not p.exists():
        return None
    try:
        d = json.loads(p.read_text())
        return {"items_per_s": d["value"], "search_qps": d["search_qps"], "cores": d["cpu_baseline"]["cores"],
                "file": "profiles/r02_reference_full_1m.json"}
    except Exception:
        return None


# --------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import arrowspace_b200 as asb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at VERSION level; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    asb._build.build_cuda()
    stream = torch.cuda.current_stream().cuda_stream
    ctx = asb.Context(local_rank, stream=stream if stream else None)
    comm = None
    if world > 1:
        # the data path's collectives run inside the C ABI on its own NCCL communicator (asb_comm_*): torch.distributed
        # only ships the 128-byte ncclUniqueId and the timing reductions
        box = [asb.host.Comm.make_unique_id(ctx) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = asb.host.Comm(ctx, world, rank, box[0])

    f, nq = args.f, args.nq
    if args.strong:      # --items is the global row count; every rank takes an equal slice (the last one the remainder)
        n_global = args.n
        per = (n_global + world - 1) // world
        lo = min(n_global, rank * per)
        n = min(n_global, lo + per) - lo
    else:                # weak scaling: --items rows per GPU
        n = args.n
        n_global = n * world
        lo = rank * n
    if args.device_data:
        args.no_e2e = True
        args.no_cpu_baseline = True
    t_gen = time.perf_counter()
    if args.device_data:
        # the same model (64 non-negative blobs + 0.05 noise), drawn on the GPU: centres from one seed on every rank, rows
        # from a per-rank stream -- the global dataset is the concatenation of the shards
        g0 = torch.Generator(device=dev).manual_seed(DATA_SEED)
        centres = torch.rand((64, f), dtype=torch.float64, device=dev, generator=g0)
        g1 = torch.Generator(device=dev).manual_seed(DATA_SEED * 1000 + rank)
        rows_d = torch.empty((n, f), dtype=torch.float64, device=dev)
        for r0 in range(0, n, 262144):
            r1 = min(n, r0 + 262144)
            lab = torch.randint(0, 64, (r1 - r0,), device=dev, generator=g1)
            blk = centres[lab]
            blk += 0.05 * torch.randn((r1 - r0, f), dtype=torch.float64, device=dev, generator=g1)
            rows_d[r0:r1] = blk.clamp_(min=0.0)
        # queries = the first rows of rank 0's shard x 1.02, broadcast
        queries_d = (rows_d[:nq] * 1.02).contiguous() if rank == 0 else torch.empty((nq, f), dtype=torch.float64, device=dev)
        if world > 1:
            dist.broadcast(queries_d, src=0)
        rows_h = queries_h = None
        head = rows_d[: min(n, 50_000)].cpu().numpy()
    else:
        rows_h = torch.empty((n, f), dtype=torch.float64).pin_memory()
        asb.synth.protein_like(n, f, seed=DATA_SEED, out=rows_h.numpy(), row0=lo)
        q_idx = asb.synth.query_indices(n_global, nq, QUERY_SEED)
        queries_h = torch.from_numpy(asb.synth.rows_at(q_idx, f, DATA_SEED) * 1.02).pin_memory()
        rows_d = rows_h.to(dev, non_blocking=True)
        queries_d = queries_h.to(dev, non_blocking=True)
        head = rows_h.numpy()[: min(n, 50_000)]
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    maxk, radius = cluster_inputs(n_global, f, head)
    if world > 1:
        t = torch.tensor([radius], dtype=torch.float64, device=dev)
        dist.broadcast(t, src=0)
        radius = float(t.item())
    gp = asb.GraphParams(GRAPH["eps"], GRAPH["k"], GRAPH["topk"], GRAPH["p"], GRAPH["sigma"])
    tm = asb.TauMode.Median
    # Two-NN samples: 500 rows of the GLOBAL dataset (at N > 1 the per-shard two nearest are merged across ranks)
    sample_idx_h = asb.heuristics.sample_indices(n_global, 500, 129)
    sample_idx = torch.from_numpy(sample_idx_h).to(dev)
    d1 = torch.empty(500, dtype=torch.float64, device=dev)
    d2 = torch.empty(500, dtype=torch.float64, device=dev)
    lib = ctx.lib
    import ctypes as C

    def builder():
        return (asb.ArrowSpaceBuilder.new(ctx).with_lambda_graph(GRAPH["eps"], GRAPH["k"], GRAPH["topk"], GRAPH["p"],
                                                                 GRAPH["sigma"])
                .with_synthesis(tm).with_seed(DATA_SEED).with_inline_sampling(None).with_dims_reduction(False, None)
                .with_cluster_params(maxk, radius))

    compute = asb.parallel.GpuCompute(ctx)
    if args.exact_search:
        ctx.set_option("search_prefilter", 0)
    if args.no_cluster_replay:
        ctx.set_option("cluster_replay", 0)
    kernel_acc = {"twonn_kernel": [], "cluster_kernel": [], "taumode_kernel": [], "search_kernel": [],
                  "search_pf_kernel": [], "search_pf_prep": [], "search_pf_finish": []}
    pf_diag = {}
    shard_diag = {}
    stage_acc = {"twonn": [], "cluster": [], "laplacian": [], "taumode": [], "search": []}

    def twonn():
        if world == 1:
            ctx.check(lib.asb_twonn_distances(ctx.handle, rows_d.data_ptr(), n, f, sample_idx.data_ptr(), 500,
                                              d1.data_ptr(), d2.data_ptr()))
        else:
            ctx.check(lib.asb_twonn_distances_sharded(ctx.handle, comm.handle, rows_d.data_ptr(), n, f, lo,
                                                      sample_idx.data_ptr(), 500, d1.data_ptr(), d2.data_ptr()))

    def step_resident(record: bool):
        """One step with every input already resident in HBM.  Returns (build_ms, search_ms)."""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        twonn()
        if world == 1:
            bp = asb.host.BuildParamsC(gp.to_c(), tm.mode, tm.value, maxk, radius, 0)
            h = C.c_void_p()
            ctx.check(lib.asb_index_build(ctx.handle, rows_d.data_ptr(), n, f, C.byref(bp), C.byref(h)))
            e[1].record()
            idx = torch.empty((nq, TOPK), dtype=torch.int64, device=dev)
            score = torch.empty((nq, TOPK), dtype=torch.float64, device=dev)
            count = torch.empty(nq, dtype=torch.int64, device=dev)
            ctx.check(lib.asb_index_search(ctx.handle, h, queries_d.data_ptr(), nq, TOPK, ALPHA, idx.data_ptr(),
                                           score.data_ptr(), count.data_ptr(), None))
            e[2].record()
            torch.cuda.synchronize()
            info = asb.host.IndexInfoC()
            lib.asb_index_info_get(h, C.byref(info))
            if record:
                stage_acc["cluster"].append(info.ms_cluster)
                stage_acc["laplacian"].append(info.ms_laplacian)
                stage_acc["taumode"].append(info.ms_taumode)
            result = (idx, score, info)
            lib.asb_index_destroy(h)
        else:
            bp = asb.host.BuildParamsC(gp.to_c(), tm.mode, tm.value, maxk, radius, 0)
            index = asb.host.ShardedIndex(ctx, comm, rows_d, lo, n_global, bp)
            e[1].record()
            idx, score, count, _ = index.search(queries_d, TOPK, ALPHA)
            e[2].record()
            torch.cuda.synchronize()
            info = index.info()
            if record:
                shard_diag["search_local_ms"] = ctx.kernel_ms("sharded_search_local_ms")
                shard_diag["search_exchange_ms"] = ctx.kernel_ms("sharded_search_exchange_ms")
            if record:
                stage_acc["cluster"].append(info.ms_cluster)
                stage_acc["laplacian"].append(info.ms_laplacian)
                stage_acc["taumode"].append(info.ms_taumode)
                for key in ("cluster_shard_speculative", "cluster_shard_fallback", "shard_t_snapshot_ms", "shard_t_ranked_ms",
                            "shard_t_state_in_ms", "shard_t_walked_ms", "shard_t_done_ms"):
                    shard_diag[key] = ctx.kernel_ms(key)
            result = (idx, score, info)
            index.close()
        if record:
            pf_used = ctx.kernel_ms("search_pf_used") == 1.0
            for kname in kernel_acc:
                if kname.startswith("search_pf") and not pf_used:
                    continue      # the exact kernel answered (option off, or the prefilter fell back)
                if kname == "search_kernel" and pf_used:
                    continue
                kernel_acc[kname].append(ctx.kernel_ms(kname))
            if pf_used:
                for kname in ("search_pf_candidates", "search_pf_rescored", "search_pf_cap", "search_pf_slabs",
                              "search_pf_band"):
                    pf_diag[kname] = ctx.kernel_ms(kname)
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident(False)
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    build_ms, search_ms = [], []
    t_start.record()
    last = None
    for _ in range(args.steps):
        b, s, last = step_resident(True)
        build_ms.append(b)
        search_ms.append(s)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    launches = ctx.kernel_launches - launches0

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = max_over_ranks(total_ms)
    build_ms_mean = max_over_ranks(float(np.mean(build_ms)))
    search_ms_mean = max_over_ranks(float(np.mean(search_ms)))
    ms_per_step = total_ms / args.steps
    items_per_s = n_global / (build_ms_mean * 1e-3)
    qps = nq / (search_ms_mean * 1e-3)

    # ---- end to end through the host-buffer API (pinned host -> H2D inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_build, e2e_search = [], []
        h2d = n * f * 8 + nq * f * 8
        d2h = 0
        for it in range(1 + max(1, min(args.steps, 2))):
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                aspace, gl = builder().build(rows_h.numpy())          # H2D of the rows + D2H of every output
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                idx, score, count, lq = aspace.search_batch(queries_h.numpy(), TOPK, ALPHA)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                d2h = (aspace.lambdas.nbytes + gl.init_data.nbytes + aspace.cluster_assignments.nbytes +
                       aspace.cluster_sizes.nbytes + gl.indptr.nbytes + gl.indices.nbytes + gl.data.nbytes +
                       idx.nbytes + score.nbytes + count.nbytes + lq.nbytes)
                aspace._release()
            else:
                bp = asb.host.BuildParamsC(gp.to_c(), tm.mode, tm.value, maxk, radius, 0)
                index = asb.host.ShardedIndex(ctx, comm, rows_h.numpy(), lo, n_global, bp)   # host rows: H2D inside
                lam_h = index.lambdas()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                idx_h, score_h, count_h, _ = index.search(queries_h.numpy(), TOPK, ALPHA)     # host in, host out
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                d2h = lam_h.size * 8 + idx_h.size * 8 + score_h.size * 8 + count_h.size * 8
                index.close()
            if it > 0:
                e2e_build.append(max_over_ranks((t1 - t0) * 1e3))
                e2e_search.append(max_over_ranks((t2 - t1) * 1e3))
        e2e = {"value": n_global / (float(np.mean(e2e_build)) * 1e-3), "unit": "items/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "search_qps": nq / (float(np.mean(e2e_search)) * 1e-3),
               "build_ms": float(np.mean(e2e_build)), "search_ms": float(np.mean(e2e_search))}

    if world > 1:
        all_diag = [None] * world
        dist.all_gather_object(all_diag, shard_diag)
        shard_diag = {f"rank{i}": d for i, d in enumerate(all_diag)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (algorithmic work per launch / CUDA-event duration of that kernel)
    hbm_peak, peak_src = measured_peaks()
    dmma_peak, dfma_peak = fp64_peak()
    kms = {k: float(np.mean(v)) if v else 0.0 for k, v in kernel_acc.items()}
    info_nnz = int(last[2].nnz)
    kernels = {}
    if kms["taumode_kernel"] > 0:
        by = n * (8 * f + 16)        # read the item once, write lambda + norm  (8F + 8 per item)
        ach = by / (kms["taumode_kernel"] * 1e-3) / 1e9
        kernels["taumode_kernel"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": ach / hbm_peak, "ms": kms["taumode_kernel"], "traffic": None}
    if kms["search_kernel"] > 0:
        fl = 2.0 * nq * n * f
        ach = fl / (kms["search_kernel"] * 1e-3) / 1e12
        kernels["search_kernel"] = {"bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                                    "frac": ach / dmma_peak, "ms": kms["search_kernel"], "traffic": None,
                                    "peak_source": "FP64 DMMA m8n8k4 micro-benchmark on this pool "
                                                   "(profiles/r01_fp64_peak.json); DFMA peak %.1f" % dfma_peak}
    if kms["search_pf_kernel"] > 0:
        # certified prefilter (csrc/search_pf.cuh): the same 2 Q N F algorithmic flops, executed as 3 TF32 MMAs per
        # product (3xTF32) on the mma.sync tensor path; peak = the mma.sync TF32 rate measured on this pool by
        # tools/mma_peak.cu (0.5 m16n8k8 MMA / clk / SM at the boost clock the clocks line reports)
        fl = 2.0 * nq * n * f
        ach = fl / (kms["search_pf_kernel"] * 1e-3) / 1e12
        umma = ctx.kernel_ms("search_pf_umma") == 1.0
        bf16 = umma and ctx.kernel_ms("search_umma_bf16") == 1.0
        if bf16:
            # BF16x3 planes on tcgen05.mma kind::f16 (csrc/search_umma.cuh): peak = the dense bf16 throughput the driver
            # measured on this pool (MEASURED_PEAKS.json, cuBLAS 8192^3: the burst figure -- the kernel is timed alone)
            tf32_peak, peak_src_pf = bf16_peak()
        elif umma:
            # the tile runs on tcgen05.mma kind::tf32 (csrc/search_umma.cuh); peak = the tcgen05 TF32 issue rate measured
            # on this pool by tools/umma_peak.cu (M128 N256 K8 at 128 cycles per instruction on every SM)
            tf32_peak, peak_src_pf = umma_tf32_peak()
        else:
            tf32_peak = 148 * 0.5 * 2048 * 1.965e9 / 1e12
            peak_src_pf = ("mma.sync TF32 m16n8k8 micro-benchmark on this pool (tools/mma_peak.cu): 0.5 MMA/clk/SM = 298 "
                           "TFLOP/s (search_umma = 0: the mma.sync tile)")
        kernels["search_pf_kernel"] = {"bound": "tensor", "achieved": 3.0 * ach, "peak": tf32_peak, "unit": "TFLOP/s",
                                       "frac": 3.0 * ach / tf32_peak, "ms": kms["search_pf_kernel"], "traffic": None,
                                       "algorithmic_tflops": ach, "algorithmic_frac": ach / tf32_peak,
                                       "peak_source": peak_src_pf + "; `achieved` counts the 3 MMAs the kernel executes per "
                                                      "algorithmic product (split operands: hi*lo + lo*hi + hi*hi), "
                                                      "algorithmic_tflops = 2 Q N F / time",
                                       "tile": ("tcgen05.mma kind::f16 (BF16x3) + TMA + TMEM" if bf16 else
                                                "tcgen05.mma kind::tf32 (3xTF32) + TMA + TMEM" if umma else "mma.sync + cp.async"),
                                       "vs_fp64_dmma_peak": ach / dmma_peak,
                                       "prep_ms": kms["search_pf_prep"], "finish_ms": kms["search_pf_finish"],
                                       "candidates_per_query": pf_diag.get("search_pf_candidates", 0.0) / max(nq, 1),
                                       "rescored_per_query": pf_diag.get("search_pf_rescored", 0.0) / max(nq, 1),
                                       "band": pf_diag.get("search_pf_band"), "slabs": pf_diag.get("search_pf_slabs")}
    if kms["twonn_kernel"] > 0:
        fl = 2.0 * 500 * n * f
        ach = fl / (kms["twonn_kernel"] * 1e-3) / 1e12
        kernels["twonn_kernel"] = {"bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                                   "frac": ach / dmma_peak, "ms": kms["twonn_kernel"], "traffic": None}
    if kms["cluster_kernel"] > 0:
        by = n * 8 * f
        ach = by / (kms["cluster_kernel"] * 1e-3) / 1e9
        kernels["cluster_kernel"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": ach / hbm_peak, "ms": kms["cluster_kernel"], "traffic": None,
                                     "note": "order-dependent walk: dependency bound by design (32-row blocks, the "
                                             "resolve of block b overlaps the tensor-core distance tile of block b+1), "
                                             "rows/s = %.0f" % (n / (kms["cluster_kernel"] * 1e-3)),
                                     "variant": ctx.kernel_ms("cluster_variant"),
                                     "exact_rows": ctx.kernel_ms("cluster_exact_rows")}
    # DRAM traffic per launch: from the committed ncu launch list of this very command (profiles/r02_traffic.json)
    if world == 1:
        for kname, sub in (("taumode_kernel", "taumode_reg_kernel"), ("search_pf_kernel", "search_umma_kernel"),
                           ("search_kernel", "search_kernel"), ("twonn_kernel", "search_kernel"),
                           ("cluster_kernel", "cluster_f32p_kernel")):
            if kname in kernels:
                kernels[kname]["traffic"] = ncu_traffic(sub, n, f, nq)
    dominant = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    roofline = dict(kernels[dominant], kernel=dominant) if dominant else None
    if roofline is not None:
        # HBM peaks come from MEASURED_PEAKS.json (or the profiling guide's fallback); that file has no FP64 entry,
        # so FP64 tensor-bound kernels are held against the DMMA rate measured on this pool by tools/fp64_peak.cu
        roofline["peak_kind"] = peak_src if roofline["bound"] == "hbm" else (
            kernels[dominant].get("peak_source", "") if dominant == "search_pf_kernel" else
            "measured on this pool by tools/fp64_peak.cu (profiles/r01_fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure")

    line = {
        "metric": "lambda_tau_build_items_per_s", "value": items_per_s, "unit": "items/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (drawn on the GPU)" if args.device_data else "synthetic",
        "config": {"workload": f"{n_global}x{f} lambda-tau build ({n} rows/GPU) + {nq}-query lambda-aware search k={TOPK}",
                   "build": "Two-NN scan + incremental clustering + feature Laplacian + taumode (ArrowSpaceBuilder::build)",
                   "max_clusters": maxk, "radius": radius, "graph": GRAPH, "taumode": "Median", "alpha": ALPHA,
                   "l2": "inputs (3.07 GB items per GPU) larger than the 126 MB L2; no flush needed",
                   "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                   "nnz_laplacian": info_nnz, "data_gen_s": t_gen},
        "search_qps": qps, "build_ms": build_ms_mean, "search_ms": search_ms_mean,
        "stages_ms": {k: float(np.mean(v)) for k, v in stage_acc.items() if v},
        "cluster_replay": ({"chunks": ctx.kernel_ms("cluster_replay_chunks"), "proven": ctx.kernel_ms("cluster_replay_chunks_ok"),
                            "rows": ctx.kernel_ms("cluster_replay_rows"), "sequential_ms": ctx.kernel_ms("cluster_replay_seq_ms"),
                            "top2_ms": ctx.kernel_ms("cluster_replay_top2_ms"), "chain_ms": ctx.kernel_ms("cluster_replay_chain_ms"),
                            "note": "kernels.cluster_kernel.ms is the LAST sequential launch only when the replay is on"}
                           if not args.no_cluster_replay else None),
        "sharded": ({"collectives": "inside the C ABI (asb_comm_*: dlopen'd NCCL): Two-NN sample all-reduce + 2-min all-gather, "
                                    "centroid-state broadcast / send / recv, CSR broadcast, lambda-stat all-reduce, top-k "
                                    "all-gather + merge", "per_rank": shard_diag} if world > 1 else None),
        "kernels": kernels, "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
    }

    # ---- a second, UNFRIENDLY clustering workload (outside the timed region): far more natural clusters than centroids
    # (512 blobs on K = 128), so every centroid averages several blobs, keeps drifting and rows sit between centroids --
    # the certified replay proves little here and the sequential kernel carries the walk
    if world == 1 and not args.no_unfriendly:
        nu, fu, ku = 200_000, f, 128
        gu = torch.Generator(device=dev).manual_seed(7)
        cu = torch.rand((512, fu), dtype=torch.float64, device=dev, generator=gu)
        xu = cu[torch.randint(0, 512, (nu,), device=dev, generator=gu)]
        xu += 0.05 * torch.randn((nu, fu), dtype=torch.float64, device=dev, generator=gu)
        xu.clamp_(min=0.0)
        ru = asb.heuristics.pilot_radius(xu[:50_000].cpu().numpy(), ku, asb.heuristics.CLUSTERING_SEED)
        res_u = {}
        for name, opt in (("replay", 1), ("sequential", 0)):
            ctx.set_option("cluster_replay", opt)
            for _ in range(2):
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                cent_u, asg_u, sizes_u = ctx.cluster_incremental(xu, ku, ru)
                ev1.record()
                torch.cuda.synchronize()
            res_u[name] = {"ms": ev0.elapsed_time(ev1), "us_per_row": ev0.elapsed_time(ev1) * 1e3 / nu,
                           "chunks": ctx.kernel_ms("cluster_replay_chunks"), "proven": ctx.kernel_ms("cluster_replay_chunks_ok"),
                           "rows_replayed": ctx.kernel_ms("cluster_replay_rows"), "exact_rows": ctx.kernel_ms("cluster_exact_rows"),
                           "clusters": int(len(cent_u))}
        ctx.set_option("cluster_replay", 0 if args.no_cluster_replay else 1)
        line["cluster_unfriendly"] = dict(res_u, workload=f"{nu}x{fu}, 512 blobs, max_clusters {ku}, radius {ru:.4f} (device-drawn)")
        del xu

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (rank 0, N=1 only)
    if world == 1 and not args.no_cpu_baseline:
        from oracle_binding import Oracle
        o = Oracle()
        cores = o.set_num_threads(host_threads())
        n_s = min(n, 20_000)
        n_items_search = min(n, 200_000)
        q_s = max(cores, 16)
        t = cpu_port_times(o, rows_h.numpy(), queries_h.numpy()[:q_s], maxk, radius, n_s, n_items_search, cores)
        build_s, items_s, cqps = cpu_port_model(t, n_s, n_items_search, q_s, n)
        line["cpu_baseline"] = {
            "value": items_s, "unit": "items/s", "cores": cores, "kind": "port",
            "sample": f"build on the first {n_s} rows, stage by stage (row-linear stages scaled by {n}/{n_s}, Laplacian "
                      f"unscaled; clustering is sequential in the reference's deterministic mode); search {q_s} queries x "
                      f"{n_items_search} items scaled to N={n}",
            "search_qps": cqps, "stage_s_on_sample": t, "build_s_model": build_s, "full_size_check": full_size_record()}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def _emit(line: dict) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout; everything else that writes to fd 1 while the
    bench runs (NCCL's version banner, library chatter) has been pointed at stderr."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
