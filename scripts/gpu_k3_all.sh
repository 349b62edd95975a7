mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -4 gpurun_out/pytest_k3.log
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 2 > gpurun_out/bench_batch_m2.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 > gpurun_out/bench_batch_m3.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 --k3-cluster 4 > gpurun_out/bench_batch_m3c4.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 --k3-cluster 1 > gpurun_out/bench_batch_m3c1.log 2>&1
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --batch-mode 3 --dim 768 --rows 5000000 > gpurun_out/bench_batch_m3d768.log 2>&1
for f in bench_batch_m2 bench_batch_m3 bench_batch_m3c4 bench_batch_m3c1 bench_batch_m3d768; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"],2),"ms", round(d["value"]),"qps", round(d["roofline"]["achieved"],1),"TF", d["config"]["k3_fallback_queries"], d["verified_against_k2"], d["clocks"])
PY
tail -2 gpurun_out/$f.log | cut -c1-300 | grep -v "^{"
done
