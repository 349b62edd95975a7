# Final evidence run for the round: full GPU suite, smoke, K2 + K0 ncu evidence (each after the plain
# command exited 0), and every plain bench line the docs quote.
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_single_$R.csv $CMD > gpurun_out/ncu_l1.log 2>&1
$CMD > gpurun_out/plain_k2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/k2_scan_$R $CMD > gpurun_out/ncu_f1.log 2>&1
CMDP="python bench.py --workload pool --steps 4 --warmup 3 --pool-texts 16384 --pool-full-mask"
$CMDP > gpurun_out/plain_k0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pool_kernel -s 3 -c 1 -o gpurun_out/k0_pool_$R $CMDP > gpurun_out/ncu_f0.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log
timeout 900 python bench.py --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
timeout 900 python bench.py --rows 10000000 --dim 768 --k 100 --no-cpu-baseline > gpurun_out/bench_768.log 2>&1
timeout 600 python bench.py --workload pool --steps 20 --warmup 3 --pool-texts 16384 > gpurun_out/bench_pool_16384.log 2>&1
timeout 600 python bench.py --workload pool --steps 20 --warmup 3 --pool-texts 16384 --pool-full-mask > gpurun_out/bench_pool_full_16384.log 2>&1
timeout 900 python bench.py --workload batch --steps 5 --warmup 3 > gpurun_out/bench_batch.log 2>&1; echo "rc=$?" >> gpurun_out/bench_batch.log
timeout 900 python bench.py --workload ingest --dim 768 --k 100 > gpurun_out/bench_ingest768.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ingest768.log
timeout 600 python bench.py --workload config1 > gpurun_out/bench_config1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_config1.log
for f in bench_default bench_1m bench_768; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$f", round(d["ms_per_step"]*1e3,2),"us", round(d["value"],1),"qps  e2e", round(d["e2e"]["value"],1), d["e2e"].get("latency"), "frac", round(d["roofline"]["frac"],4), d["verified"], d["gpu_launches"], d["clocks"])
PY
done
for f in bench_reference bench_batch bench_ingest768 bench_config1 bench_pool_16384 bench_pool_full_16384; do echo "== $f"; tail -c 1800 gpurun_out/$f.log; done
