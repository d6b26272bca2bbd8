"""per-kernel times of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import csv, sys
f = sys.argv[1]; nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
out = []
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    out.append((r[ki].split("(")[0][:70], v))
per = len(out) // nsteps
last = out[-per:]
agg = {}
for k, v in last:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in last)
print(f"# last of {nsteps} steps: {per} launches, {tot:.3f} ms (ncu: cold cache, serialised)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v:8.3f} ms {100*v/tot:5.1f}%  {c:3d} x {k}")
