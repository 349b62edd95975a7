# full GPU suite + default bench + reference arm + ncu refresh of K3 (after the plain run exits 0)
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log
timeout 900 python bench.py --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
CMDB="python bench.py --workload batch --steps 2 --warmup 3"
$CMDB > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_$R.csv $CMDB > gpurun_out/ncu_l2.log 2>&1
$CMDB > gpurun_out/plain_k3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_scan -s 2 -c 1 -o gpurun_out/k3_batch_$R $CMDB > gpurun_out/ncu_f2.log 2>&1
tail -c 1500 gpurun_out/bench_default.log; tail -c 600 gpurun_out/bench_reference.log; tail -c 700 gpurun_out/bench_1m.log
