# last evidence of the round: config-5 query shape (10M x 768, k = 100) as a stream, persistent against chained, and the
# ncu launch list of the maintenance workload (after the plain command exited 0)
mkdir -p gpurun_out
B="python bench.py --rows 10000000 --dim 768 --k 100 --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-verify"
timeout 300 $B > gpurun_out/k768_p.log 2>&1; timeout 300 $B --launch-per-query > gpurun_out/k768_q.log 2>&1
timeout 300 $B > gpurun_out/k768_p2.log 2>&1; timeout 300 $B --launch-per-query > gpurun_out/k768_q2.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k768_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], round(d["ms_per_step"], 4), "ms", round(d["value"], 1), "qps e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"], d["repeats"]["device_ms_per_step"]["all"])
PY
CMD="python bench.py --workload maintenance --maint-rows 2000000 --maint-save-rows 200000"
timeout 300 $CMD > gpurun_out/plain_maint.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_maint.csv $CMD > gpurun_out/ncu_l_maint.log 2>&1
grep -c compact gpurun_out/launches_maint.csv
