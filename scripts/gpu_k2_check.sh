mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not k3 and not 10m" > gpurun_out/pytest_k2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k2.log; tail -3 gpurun_out/pytest_k2.log
for rows in 1000000 10000000; do
timeout 600 python bench.py --steps 300 --warmup 20 --rows $rows --no-cpu-baseline > gpurun_out/bench_r$rows.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bench_r$rows.log"):
    if l.startswith("{"):
        d=json.loads(l); print("rows=$rows", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "GB/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), d["verified"], d["clocks"])
PY
done
