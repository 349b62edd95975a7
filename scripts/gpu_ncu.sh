# ncu evidence for profiles/ (each ncu run only after the same command exited 0 without ncu)
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_single_$R.csv $CMD > gpurun_out/ncu_l1.log 2>&1
$CMD > gpurun_out/plain_k2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/k2_scan_$R $CMD > gpurun_out/ncu_f1.log 2>&1
CMDB="python bench.py --workload batch --steps 2 --warmup 3"
$CMDB > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_$R.csv $CMDB > gpurun_out/ncu_l2.log 2>&1
$CMDB > gpurun_out/plain_k3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_scan -s 2 -c 1 -o gpurun_out/k3_batch_$R $CMDB > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out | tail -12
