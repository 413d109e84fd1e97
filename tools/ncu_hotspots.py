"""Summarise `ncu --page source --csv` output: stall samples / executed instructions per SASS region.
    ncu -i X.ncu-rep --page source --csv | python tools/ncu_hotspots.py [bin]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[iS]) for r in data) or 1
tot_e = sum(int(r[iE]) for r in data) or 1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 150
print(f"total samples {tot_s}  total warp-instructions {tot_e}  sass lines {len(data)}")
for b in range(0, len(data), B):
    seg = data[b:b + B]
    s = sum(int(r[iS]) for r in seg)
    e = sum(int(r[iE]) for r in seg)
    if s / tot_s > 0.01 or e / tot_e > 0.01:
        ops = {}
        for r in seg:
            t = r[1].strip().split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
            ops[op] = ops.get(op, 0) + int(r[iE])
        st = {}
        for c in stall_cols:
            v = sum(int(r[c] or 0) for r in seg)
            if v:
                st[hdr[c]] = v
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        tst = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"[{b:5d}-{b + B:5d}] samples {100 * s / tot_s:5.1f}%  inst {100 * e / tot_e:5.1f}%  "
              f"{[(k, round(v / 1e6, 1)) for k, v in top]}  {[(k, round(100 * v / tot_s, 1)) for k, v in tst]}")
print("== top single instructions by samples")
for i in sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:14]:
    r = data[i]
    print(f"  #{i:5d} {100 * int(r[iS]) / tot_s:5.1f}%  exec {r[iE]:>10}  {r[1].strip()[:80]}")
