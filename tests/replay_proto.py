"""CPU feasibility study for a PARALLEL, still exact, replacement of the order-dependent clustering walk
(src/clustering.rs:547-928) once the centroid count has saturated -- "optimistic replay with certification".

The walk is sequential because row r sees the centroids every earlier row left behind.  But:
  * with K = max_clusters no centroid is created any more; a row only (a) picks its nearest centroid b, (b) compares
    d^2 with radius / 1.5 radius, (c) if d^2 <= radius moves c_b by (x - c_b) / k  (clustering.rs:711-781);
  * GIVEN the assignment vector A of a chunk of rows, the K centroid trajectories are independent sequential chains
    (each centroid only sees its own rows, in row order) -- K-way parallel, element-wise IEEE arithmetic, hence
    bit-identical to the walk whenever A is the walk's assignment;
  * GIVEN trajectories, every row's decision can be re-derived in parallel: distances to the chunk-start snapshot S0
    (one dense contraction) bound the distance to any centroid at the row's time by the path length that centroid has
    travelled inside the chunk so far; only rows whose bounds do not separate need exact distances to the few
    competitors.
  * Fixed point: start from A = decisions against S0, recompute trajectories, re-derive decisions, repeat until nothing
    changes.  If the decisions derived from the trajectories of A equal A, then by induction over r (row r depends only
    on rows < r) A IS the sequential walk's assignment -- exactness needs no bound on the number of sweeps.

This script replays the saturated part of a walk that way (numpy; updates follow the guess A, decisions are re-derived
row by row against the guessed trajectories -- exactly what a parallel implementation would see) and reports sweeps per
chunk, the share of rows that needed exact competitor distances, and whether assignments / centroids / counts match
the oracle's sequential walk bit for bit.   python tests/replay_proto.py [n] [f] [chunk]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import arrowspace_b200 as asb  # noqa: E402  (synthetic data + the host heuristics only; no GPU code is touched)
from oracle_binding import Oracle  # noqa: E402  (test infrastructure: lives under tests/ because it runs the oracle)


def replay_chunk(X, C0, cnt0, radius, stats):
    """One chunk of rows against the exact chunk-start state (C0, cnt0).  Returns (assignments, C, cnt)."""
    m, K = len(X), len(C0)
    x2 = np.einsum("ij,ij->i", X, X)
    c2 = np.einsum("ij,ij->i", C0, C0)
    D0 = np.maximum(x2[:, None] + c2[None, :] - 2.0 * (X @ C0.T), 0.0)
    d0 = np.sqrt(D0)                                     # distance to the snapshot, error ~1e-7 relative at worst
    slack = 1e-6 * (np.sqrt(x2).max() + np.sqrt(c2).max())   # covers the GEMM-form cancellation error

    def classify(d2):
        return 0 if d2 <= radius else (1 if d2 <= 1.5 * radius else 2)   # update / count only / drop

    A = D0.argmin(1)
    cls = np.array([classify(D0[r, A[r]]) for r in range(m)], dtype=np.int8)
    sweeps = 0
    while True:
        sweeps += 1
        C, cnt = C0.copy(), cnt0.astype(np.float64).copy()
        disp = np.zeros(K)                               # path length of every centroid inside the chunk so far
        A2, cls2 = A.copy(), cls.copy()
        exact_rows = 0
        for r in range(m):
            x = X[r]
            b = A[r]
            diff = x - C[b]
            db2 = float(diff @ diff)                     # distance to b's state at row r (reference: sequential sum)
            db = np.sqrt(db2)
            lower = d0[r] - disp - slack                 # every other centroid is at least this far at row r
            lower[b] = np.inf
            # certified: b strictly nearest (ties -> lower index needs strictness against lower ids only; be strict)
            ok = db < lower.min()
            thr_close = min(abs(db2 - radius), abs(db2 - 1.5 * radius)) < 1e-9 * max(radius, 1.0)
            if ok and not thr_close:
                nb, ncls = b, classify(db2)
            else:                                        # exact distances to the competitors' states at row r
                exact_rows += 1
                cand = np.nonzero(d0[r] - disp - slack <= db)[0]
                cand = np.union1d(cand, [b])
                dd = C[cand] - x
                dc2 = np.einsum("ij,ij->i", dd, dd)
                j = int(dc2.argmin())                    # first minimum = lowest id among exact ties
                nb, ncls = int(cand[j]), classify(float(dc2[j]))
            A2[r], cls2[r] = nb, ncls
            # the trajectories follow the GUESS (what a parallel implementation computed before verifying)
            if cls[r] == 0:
                k_new = cnt[b] + 1.0
                step = (x - C[b]) / k_new                # clustering.rs:747-751, element-wise IEEE
                C[b] += step
                disp[b] += np.sqrt(float(step @ step)) * (1.0 + 1e-12)
                cnt[b] = k_new
            elif cls[r] == 1:
                cnt[b] += 1.0
        changed = int(((A2 != A) | (cls2 != cls)).sum())
        stats["exact_rows"].append(exact_rows)
        stats["changed"].append(changed)
        if changed == 0:
            break
        A, cls = A2, cls2
    stats["sweeps"].append(sweeps)
    asg = np.where(cls == 2, -1, A)
    return asg, C, cnt


def single_sweep_chunk(X, C0, cnt0, radius, kmax):
    """The GPU-shaped variant: ONE sweep, all or nothing.  (1) top-2 snapshot distances per row (a dense contraction
    with a fused 2-min, the existing Two-NN kernel shape); (2) K independent chains apply the rows of each centroid in
    order, classify each row from the exact current distance and track the centroid's NET displacement from the
    snapshot; (3) a row is certified when its current distance to b is below (second-best snapshot distance) - (largest
    net displacement any centroid reached in the chunk).  All rows certified -> the guess was the walk's assignment
    (induction over rows) and the chain results are the walk's bits; otherwise the caller runs the sequential kernel
    on this chunk from the same start state (asb_cluster_incremental_resume).  Works unsaturated too: a row that
    would open a new centroid (d^2 > radius / 2 while K < max) fails the chunk."""
    m, K = len(X), len(C0)
    x2 = np.einsum("ij,ij->i", X, X)
    c2 = np.einsum("ij,ij->i", C0, C0)
    D0 = np.maximum(x2[:, None] + c2[None, :] - 2.0 * (X @ C0.T), 0.0)
    A = D0.argmin(1)
    part = np.partition(D0, 1, axis=1)[:, :2]
    second = np.sqrt(part[:, 1])                          # distance to the runner-up at the snapshot
    slack = 1e-6 * (np.sqrt(x2).max() + np.sqrt(c2).max())
    C, cnt = C0.copy(), cnt0.copy()
    db = np.empty(m)
    cls = np.empty(m, dtype=np.int8)
    max_disp = 0.0
    borderline = 0
    for j in np.unique(A):                                # the chains are independent: any order, here one by one
        rows = np.nonzero(A == j)[0]
        c, k = C[j].copy(), cnt[j]
        for r in rows:
            diff = X[r] - c
            d2 = float(diff @ diff)
            db[r] = np.sqrt(d2)
            if K < kmax and d2 > 0.5 * radius:
                cls[r] = 3                                # would create a centroid: structural change -> fail
            elif d2 <= radius:
                cls[r] = 0
                k += 1.0
                c += diff / k
                dn = c - C0[j]
                max_disp = max(max_disp, float(np.sqrt(dn @ dn)))
            elif d2 <= 1.5 * radius:
                cls[r] = 1
                k += 1.0
            else:
                cls[r] = 2
            for thr in (0.5 * radius, radius, 1.5 * radius):
                if abs(d2 - thr) < 1e-9 * radius:
                    borderline += 1
        C[j], cnt[j] = c, k
    margin = second - max_disp - slack - db               # > 0: certified
    # a row farther than sqrt(1.5 radius) from EVERY centroid at its own time is dropped whichever centroid is the
    # nearest (no update, no count, assignment None: clustering.rs:759-815) -- e.g. a blob no centroid was opened for
    best = np.sqrt(part[:, 0])
    far = np.maximum(best - max_disp - slack, 0.0) ** 2 > 1.5 * radius * (1.0 + 1e-9)
    margin = np.where(far & (cls == 2), np.inf, margin)
    ok = bool((margin > 0).all()) and not (cls == 3).any() and borderline == 0
    info = {"ok": ok, "uncertified": int((margin <= 0).sum()), "creates": int((cls == 3).sum()),
            "dropped_whoever_is_nearest": int((far & (cls == 2)).sum()),
            "max_net_displacement": max_disp, "min_margin": float(margin.min()),
            "median_margin": float(np.median(margin)), "active_centroids": int(len(np.unique(A[cls == 0]))),
            "longest_chain": int(np.bincount(A, minlength=K).max())}
    return ok, np.where(cls == 2, -1, A), C, cnt, info


def grouped_sweep_chunk(X, C0, cnt0, radius, kmax, delta_p, T=4):
    """Single sweep with CANDIDATE SETS: robust to centroids that contest the same rows.  delta_p is an a-priori bound
    on every centroid's net displacement inside the chunk (checked afterwards).  A row's nearest centroid at its own
    time is one of the snapshot's top-T whose snapshot distance is within 2 delta_p of the best (everything farther
    cannot overtake).  Centroids that share a row's candidate set are merged into a group (union-find); each group is
    ONE sequential chain over its rows, evaluating the few candidates exactly; groups are independent.  Fails when a
    candidate set is not closed inside the top-T, when a centroid moves farther than delta_p, on a near-tie between
    candidates, on a creation or on a threshold within the guard band."""
    m, K = len(X), len(C0)
    T = min(T, K)
    x2 = np.einsum("ij,ij->i", X, X)
    c2 = np.einsum("ij,ij->i", C0, C0)
    D0 = np.maximum(x2[:, None] + c2[None, :] - 2.0 * (X @ C0.T), 0.0)
    top = np.argsort(D0, axis=1, kind="stable")[:, :T]
    d0 = np.sqrt(np.take_along_axis(D0, top, axis=1))
    slack = 1e-6 * (np.sqrt(x2).max() + np.sqrt(c2).max())
    within = d0 <= d0[:, :1] + 2.0 * delta_p + slack            # candidate mask (column 0 always true)
    far = np.maximum(d0[:, 0] - delta_p - slack, 0.0) ** 2 > 1.5 * radius * (1.0 + 1e-9)   # dropped whoever is nearest
    within[far, 1:] = False
    info = {"delta_p": delta_p, "dropped_whoever_is_nearest": int(far.sum())}
    if T < K and within[:, T - 1].any():
        info.update(ok=False, reason="candidate set not closed inside the top-T", rows=int(within[:, T - 1].sum()))
        return False, None, None, None, info
    parent = list(range(K))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    multi = np.nonzero(within.sum(1) > 1)[0]
    for r in multi:
        a = find(int(top[r, 0]))
        for i in range(1, T):
            if within[r, i]:
                b = find(int(top[r, i]))
                if a != b:
                    parent[b] = a
    label = np.array([find(j) for j in range(K)])
    row_group = label[top[:, 0]]
    C, cnt = C0.copy(), cnt0.copy()
    asg = np.empty(m, dtype=np.int64)
    max_disp, bad = 0.0, 0
    for g in np.unique(row_group):                               # groups are independent chains
        for r in np.nonzero(row_group == g)[0]:
            cands = top[r][within[r]]
            dd = C[cands] - X[r]
            dc2 = np.einsum("ij,ij->i", dd, dd)
            order = np.argsort(dc2, kind="stable")
            if len(cands) > 1 and dc2[order[1]] - dc2[order[0]] <= 1e-9 * radius:
                bad += 1                                         # near tie: the reference's summation order decides
            j, d2 = int(cands[order[0]]), float(dc2[order[0]])
            if any(abs(d2 - t) < 1e-9 * radius for t in (0.5 * radius, radius, 1.5 * radius)):
                bad += 1
            if K < kmax and d2 > 0.5 * radius:
                bad += 1
                asg[r] = -2
            elif d2 <= radius:
                cnt[j] += 1.0
                C[j] += (X[r] - C[j]) / cnt[j]
                dn = C[j] - C0[j]
                max_disp = max(max_disp, float(np.sqrt(dn @ dn)))
                asg[r] = j
            elif d2 <= 1.5 * radius:
                cnt[j] += 1.0
                asg[r] = j
            else:
                asg[r] = -1
    ok = bool(bad == 0 and max_disp <= delta_p)
    sizes = np.bincount(label, minlength=K)
    info.update(ok=ok, bad=bad, max_net_displacement=max_disp, multi_candidate_rows=int(len(multi)),
                groups=int(len(np.unique(label))), largest_group=int(sizes.max()),
                longest_group_chain=int(np.bincount(row_group, minlength=K).max()))
    return ok, asg, C, cnt, info


def main_grouped(n, f, kmax_arg, chunk, first):
    o = Oracle()
    x = asb.synth.protein_like(n, f, seed=42)
    kmax = kmax_arg if kmax_arg > 0 else asb.heuristics.step1_bounds(n, f, f)[1]
    radius = asb.heuristics.pilot_radius(x[: min(n, 50_000)], kmax, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = o.cluster_incremental(x, kmax, radius)
    C, a0, cnt = o.cluster_incremental(x[:first], kmax, radius)
    cnt = cnt.astype(np.float64)
    got, infos, lo = [a0], [], first
    delta_p = 0.05 * np.sqrt(radius)
    while lo < n:
        hi = min(lo + chunk, n)
        for attempt in range(2):
            ok, a, C2, cnt2, info = grouped_sweep_chunk(x[lo:hi], C, cnt, radius, kmax, delta_p)
            info["rows"] = [lo, hi]
            infos.append(info)
            if ok or "max_net_displacement" not in info or info.get("bad", 0):
                break
            delta_p = 2.0 * info["max_net_displacement"]          # moved farther than predicted: widen once
        if ok:
            C, cnt = C2, cnt2
            got.append(a)
            delta_p = max(1.5 * info["max_net_displacement"], 1e-6)
        else:
            got.append(asg[lo:hi])
            C, _, cntn = o.cluster_incremental(x[:hi], kmax, radius)
            cnt = cntn.astype(np.float64)
        lo = hi
    got = np.concatenate(got)
    out = {"mode": "grouped", "n": n, "f": f, "max_clusters": int(kmax), "clusters": int(len(cent)), "radius": radius,
           "sequential_prefix_rows": first, "chunk": chunk, "attempts": len(infos),
           "chunks_ok": int(sum(i["ok"] for i in infos)),
           "assignments_equal": bool(np.array_equal(got, asg)),
           "centroids_bit_identical": bool(C.shape == cent.shape and np.array_equal(C.view(np.uint64), cent.view(np.uint64))),
           "counts_equal": bool(np.array_equal(cnt.astype(np.uint64), sizes)), "per_attempt": infos}
    print(json.dumps(out))


def main_single(n, f, chunk, first):
    """Sequential walk (oracle) for the first `first` rows and for every chunk that fails; single-sweep replay else."""
    o = Oracle()
    x = asb.synth.protein_like(n, f, seed=42)
    _, kmax = asb.heuristics.step1_bounds(n, f, f)
    radius = asb.heuristics.pilot_radius(x[: min(n, 50_000)], kmax, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = o.cluster_incremental(x, kmax, radius)
    C, a0, cnt = o.cluster_incremental(x[:first], kmax, radius)
    cnt = cnt.astype(np.float64)
    got = [a0]
    infos = []
    lo = first
    while lo < n:
        hi = min(lo + chunk, n)
        ok, a, C2, cnt2, info = single_sweep_chunk(x[lo:hi], C, cnt, radius, kmax)
        info["rows"] = [lo, hi]
        infos.append(info)
        if ok:
            C, cnt = C2, cnt2
            got.append(a)
        else:   # sequential fallback from the chunk-start state == the oracle's walk restricted to these rows
            got.append(asg[lo:hi])
            Cn, _, cntn = o.cluster_incremental(x[:hi], kmax, radius)
            if len(Cn) != len(C):
                C = Cn
            else:
                C = Cn
            cnt = cntn.astype(np.float64)
        lo = hi
    got = np.concatenate(got)
    out = {"mode": "single_sweep", "n": n, "f": f, "max_clusters": int(kmax), "clusters": int(len(cent)),
           "radius": radius, "sequential_prefix_rows": first, "chunk": chunk,
           "chunks_ok": int(sum(i["ok"] for i in infos)), "chunks": len(infos),
           "rows_replayed_share": float(sum(i["rows"][1] - i["rows"][0] for i in infos if i["ok"]) / n),
           "assignments_equal": bool(np.array_equal(got, asg)),
           "centroids_bit_identical": bool(C.shape == cent.shape and np.array_equal(C.view(np.uint64), cent.view(np.uint64))),
           "counts_equal": bool(np.array_equal(cnt.astype(np.uint64), sizes)),
           "per_chunk": infos}
    print(json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "grouped":
        a = [int(v) for v in sys.argv[2:]]
        return main_grouped(*(a + [60_000, 384, 200, 8_192, 12_288][len(a):]))
    if len(sys.argv) > 1 and sys.argv[1] == "single":
        a = [int(v) for v in sys.argv[2:]]
        return main_single(*(a + [200_000, 384, 32_768, 16_384][len(a):]))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 8_192
    o = Oracle()
    x = asb.synth.protein_like(n, f, seed=42)
    _, kmax = asb.heuristics.step1_bounds(n, f, f)
    radius = asb.heuristics.pilot_radius(x[: min(n, 50_000)], kmax, asb.heuristics.CLUSTERING_SEED)
    cent, asg, sizes = o.cluster_incremental(x, kmax, radius)
    # rows that created a centroid: the assignment is the next unused id
    seen, last_new = 0, -1
    for r in range(n):
        if asg[r] == seen:
            seen += 1
            last_new = r
    r0 = last_new + 1
    out = {"n": n, "f": f, "max_clusters": int(kmax), "clusters": int(len(cent)), "radius": radius,
           "saturated_at_row": r0, "chunk": chunk}
    if len(cent) < kmax:
        out["note"] = "centroid count never saturates on this input: the replay scheme does not apply as is"
        print(json.dumps(out))
        return
    C, _, cnt = o.cluster_incremental(x[:r0], kmax, radius)        # exact state where the replay starts
    cnt = cnt.astype(np.float64)
    stats = {"sweeps": [], "exact_rows": [], "changed": []}
    got = []
    t0 = time.perf_counter()
    for lo in range(r0, n, chunk):
        a, C, cnt = replay_chunk(x[lo:lo + chunk], C, cnt, radius, stats)
        got.append(a)
    out["proto_seconds"] = time.perf_counter() - t0
    got = np.concatenate(got) if got else np.zeros(0, dtype=np.int64)
    out["assignments_equal"] = bool(np.array_equal(got, asg[r0:]))
    out["assignment_mismatches"] = int((got != asg[r0:]).sum())
    out["centroids_bit_identical"] = bool(np.array_equal(C.view(np.uint64), cent.view(np.uint64)))
    out["counts_equal"] = bool(np.array_equal(cnt.astype(np.uint64), sizes))
    out["chunks"] = len(stats["sweeps"])
    out["sweeps_per_chunk"] = {"mean": float(np.mean(stats["sweeps"])), "max": int(np.max(stats["sweeps"]))}
    rows_swept = sum(min(chunk, n - lo) * s for lo, s in zip(range(r0, n, chunk), stats["sweeps"]))
    out["rows_needing_exact_competitors_share"] = float(sum(stats["exact_rows"]) / max(rows_swept, 1))
    out["decisions_changed_after_first_sweep_share"] = float(
        sum(c for c in stats["changed"]) / max(n - r0, 1))
    out["sweeps_by_chunk"] = stats["sweeps"]
    first = np.cumsum([0] + stats["sweeps"][:-1])
    out["first_sweep_exact_share_by_chunk"] = [round(stats["exact_rows"][i] / min(chunk, n - lo), 4)
                                               for i, lo in zip(first, range(r0, n, chunk))]
    out["first_sweep_changed_by_chunk"] = [stats["changed"][i] for i in first]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
