set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_single_$R.csv $CMD > gpurun_out/ncu_l1.log 2>&1
$CMD > gpurun_out/plain_k2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/k2_scan_$R $CMD > gpurun_out/ncu_f1.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 900 python bench.py --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
tail -c 400 gpurun_out/plain_k2.log
