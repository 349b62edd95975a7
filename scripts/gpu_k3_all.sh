mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -4 gpurun_out/pytest_k3.log
for m in 2 3; do
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode $m > gpurun_out/bench_batch_m$m.log 2>&1; echo "rc=$?" >> gpurun_out/bench_batch_m$m.log
done
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 --k3-cluster 2 > gpurun_out/bench_batch_m3c2.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 --nq 128 > gpurun_out/bench_batch_m3q128.log 2>&1
for f in bench_batch_m2 bench_batch_m3 bench_batch_m3c2 bench_batch_m3q128; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"],2),"ms", round(d["value"]),"qps", round(d["roofline"]["achieved"],1),"TF", d["config"]["k3_fallback_queries"], d["verified_against_k2"], d["clocks"])
PY
tail -2 gpurun_out/$f.log | cut -c1-300 | grep -v "^{"
done
