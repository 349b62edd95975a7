mkdir -p gpurun_out
CMD="python bench.py --workload batch --rows 10000000 --steps 2 --batch-mode 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_r02d.csv $CMD > gpurun_out/ncu_batch_launches.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_batch_r02d.csv")))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); n=d["Kernel Name"][:60]+" "+d["Grid Size"]; v=float(d["Metric Value"].replace(",",""))
        agg.setdefault(n,[]).append(v)
for n,v in agg.items(): print(f"{n:90s} x{len(v):3d}  min {min(v)/1e6:9.4f} ms  max {max(v)/1e6:9.4f} ms")
PY
