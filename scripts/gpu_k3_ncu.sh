mkdir -p gpurun_out
CMD="python bench.py --workload batch --rows 2000000 --steps 1 --warmup 3"
$CMD > gpurun_out/plain_k3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_scan -s 2 -c 1 -o gpurun_out/k3_batch_r01 $CMD > gpurun_out/ncu_k3.log 2>&1
tail -c 600 gpurun_out/plain_k3.log; tail -5 gpurun_out/ncu_k3.log
