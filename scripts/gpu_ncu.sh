set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 -o gpurun_out/k2_scan_r01 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
