mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/perf_probe.py 2>&1 | grep -v rep0 | tail -9
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python tools/prof_one.py 1000000 2 > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r1d.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
data=rows[1:]; half=data[len(data)//2:]
agg=collections.OrderedDict()
for r in half:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
for k,v in agg.items():
    if v[1]>3e4: print(f"{v[1]/1e6:8.3f} ms {v[0]:3d} {k}")
PY
