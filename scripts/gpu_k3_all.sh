mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -4 gpurun_out/pytest_k3.log
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 > gpurun_out/bench_batch.log 2>&1; echo "rc=$?" >> gpurun_out/bench_batch.log
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --rows 2000000 > gpurun_out/bench_batch2m.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --nq 128 > gpurun_out/bench_batch128.log 2>&1
for f in bench_batch bench_batch2m bench_batch128; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"],2),"ms", round(d["value"]),"qps", round(d["roofline"]["achieved"],1),"TF", d["roofline"]["issued_frac"], d["config"]["k3_fallback_queries"], d["verified_against_k2"], d["clocks"])
PY
done
