# ncu --set full of the K3 kernels (single-pass stage of the cascade, three-pass split; plus the CTA-pair variant) on a
# 2M-row corpus; each capture only after the same command exited 0 without ncu
mkdir -p gpurun_out
TAG=${TAG:-r02c}
for V in "3 0 batch_scan single" "2 0 batch_scan x3"; do
set -- $V
CMD="python bench.py --workload batch --rows 2000000 --steps 1 --batch-mode $1 --k3-pair $2"
$CMD > gpurun_out/plain_k3_$4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 4 -c 2 -o gpurun_out/k3_$4_$TAG -f $CMD > gpurun_out/ncu_k3_$4.log 2>&1
tail -c 200 gpurun_out/plain_k3_$4.log; tail -2 gpurun_out/ncu_k3_$4.log
done
# launch list of the batch workload (kernel shares of a step)
CMD="python bench.py --workload batch --rows 10000000 --steps 2 --batch-mode 0"
$CMD > gpurun_out/plain_batch_launches.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_$TAG.csv $CMD > gpurun_out/ncu_batch_launches.log 2>&1
tail -3 gpurun_out/launches_batch_$TAG.csv | cut -c1-200
