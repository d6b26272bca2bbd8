"""per-kernel times per stream (= rank of an in-process group) of the LAST step in an ncu launch list"""
import csv, sys
f = sys.argv[1]; nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi, ui, si = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Stream")
by = {}
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    by.setdefault(r[si], []).append((r[ki].split("(")[0][:60], v))
agg = {}
tot = {}
for s, out in by.items():
    per = len(out) // nsteps
    last = out[-per:]
    tot[s] = (per, sum(v for _, v in last))
    for k, v in last:
        a = agg.setdefault(k, {}); a[s] = a.get(s, 0.0) + v
print("# last step per stream: " + ", ".join(f"{s}: {n} launches {t:.3f} ms" for s, (n, t) in tot.items()))
print(f"# {'kernel':60s} {'max':>8s} {'mean':>8s} {'min':>8s}  over {len(by)} streams (ms)")
for k, a in sorted(agg.items(), key=lambda kv: -max(kv[1].values())):
    vals = [a.get(s, 0.0) for s in by]
    print(f"  {k:60s} {max(vals):8.3f} {sum(vals)/len(vals):8.3f} {min(vals):8.3f}")
