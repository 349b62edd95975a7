set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_store_gpu.py -x -q -m gpu -k "not k3 and not k0" > gpurun_out/pytest_k2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k2.log
tail -5 gpurun_out/pytest_k2.log
timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_sweep.log 2>&1; cat gpurun_out/k2_sweep.log
timeout 600 python bench.py --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bench_1m.log"):
    if l.startswith("{"):
        d=json.loads(l); print("1M", round(d["ms_per_step"]*1e3,2),"us", round(d["value"],1),"qps  e2e", round(d["e2e"]["value"],1), d["e2e"]["latency"], d["verified"])
PY
