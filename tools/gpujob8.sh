mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches2.csv python tools/prof_one.py 1000000 2 > gpurun_out/ncu_launch2.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches2.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1+len(rows)//2:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:70]
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print('second rep total ms', tot/1e6)
for k,v in agg.items(): print(f"{v[1]/1e6:8.3f} ms {v[0]:3d} {k}")
PY
