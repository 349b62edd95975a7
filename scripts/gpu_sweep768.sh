mkdir -p gpurun_out
for v in 0 2; do
timeout 900 python bench.py --rows 10000000 --dim 768 --k 100 --steps 100 --warmup 10 --variant $v --no-cpu-baseline > gpurun_out/bench_768_v$v.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bench_768_v$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("768 k=100 variant $v", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps", "GB/s", round(d["roofline"]["achieved"],1), d["verified"], d["clocks"])
PY
done
