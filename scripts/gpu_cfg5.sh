mkdir -p gpurun_out
timeout 900 python bench.py --workload ingest --rows 10000000 --dim 768 --k 100 > gpurun_out/bench_ingest768.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ingest768.log
tail -c 2200 gpurun_out/bench_ingest768.log
for v in 0 1; do
timeout 900 python bench.py --rows 10000000 --dim 768 --k 100 --steps 100 --warmup 10 --variant $v --no-cpu-baseline > gpurun_out/bench_768_v$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_768_v$v.log
python - <<PY
import json
for l in open("gpurun_out/bench_768_v$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("768 k=100 variant $v", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "GB/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), d["verified"], d["clocks"])
PY
done
