mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -4 gpurun_out/pytest_k3.log
for c in 1 2 4; do
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --k3-cluster $c > gpurun_out/bench_batch_c$c.log 2>&1; echo "rc=$?" >> gpurun_out/bench_batch_c$c.log
done
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --nq 128 > gpurun_out/bench_batch128.log 2>&1
for f in bench_batch_c1 bench_batch_c2 bench_batch_c4 bench_batch128; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"],2),"ms", round(d["value"]),"qps", round(d["roofline"]["achieved"],1),"TF", round(d["roofline"]["issued_frac"],3), d["config"]["k3_fallback_queries"], d["verified_against_k2"], d["clocks"])
PY
tail -2 gpurun_out/$f.log | cut -c1-300 | grep -v "^{"
done
