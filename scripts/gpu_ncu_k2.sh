# K2 evidence: launch list + one full capture of the headline command (each after a plain run exited 0),
# then the plain benches the docs quote (default, 1M, 10M x 768 top-100) and the latency sweep
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_single_$R.csv $CMD > gpurun_out/ncu_l1.log 2>&1
$CMD > gpurun_out/plain_k2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/k2_scan_$R $CMD > gpurun_out/ncu_f1.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1
timeout 900 python bench.py --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
timeout 900 python bench.py --rows 10000000 --dim 768 --k 100 --no-cpu-baseline > gpurun_out/bench_768.log 2>&1
timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_sweep.log 2>&1
tail -c 400 gpurun_out/plain_k2.log
for f in bench_default bench_1m bench_768; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"]*1e3,2),"us", round(d["value"],1),"qps  e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4), d["verified"], d["gpu_launches"])
PY
done
cat gpurun_out/k2_sweep.log
