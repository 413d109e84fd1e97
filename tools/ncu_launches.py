"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv --log-file X` launch
list by kernel: launches, total / mean device time, share of the listed time and (when present) DRAM bytes per launch.
    python tools/ncu_launches.py gpurun_out/launches.csv [out.json] > profiles/rNN_launches.md
The per-launch times are cold-cache and serialised (B200_PROFILING.md): compare SHARES with the bench's stage times, not
absolutes.  With out.json the per-kernel DRAM traffic is also written as JSON (bench.py reads profiles/r02_traffic.json)."""
import csv
import json
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r and "Metric Name" in r:
                hdr = r
            continue
        rows.append(r)
    if hdr is None:
        print("no launch table found")
        return
    iid, ik, im, iu, iv = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = OrderedDict()
    for r in rows:
        key = r[iid]
        d = launches.setdefault(key, {"kernel": r[ik]})
        val = float(r[iv].replace(",", "")) if r[iv] not in ("", "n/a") else 0.0
        unit = r[iu]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[im]] = val * scale
    agg = OrderedDict()
    for d in launches.values():
        name = d["kernel"].split("(")[0].replace("void ", "").strip()
        name = name.replace("<unnamed>::", "")
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values()) or 1.0
    print(f"| kernel | launches | total ms | mean us | share | DRAM read+write per launch |")
    print("|---|---:|---:|---:|---:|---:|")
    out = {}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        per = (a["rd"] + a["wr"]) / a["n"]
        print(f"| `{name[:70]}` | {a['n']} | {a['us'] / 1e3:.3f} | {a['us'] / a['n']:.1f} | {100 * a['us'] / tot:.1f} % | "
              f"{per / 1e6:.1f} MB |")
        out[name] = {"launches": a["n"], "total_ms": a["us"] / 1e3, "share": a["us"] / tot,
                     "dram_bytes_per_launch": per, "dram_bytes_total": a["rd"] + a["wr"]}
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
