# ncu --set full of the two K3 kernels (single-pass stage of the cascade, bf16x3) on a 2M-row corpus; each capture only
# after the same command exited 0 without ncu
mkdir -p gpurun_out
TAG=${TAG:-r02a}
for M in 3 2; do
CMD="python bench.py --workload batch --rows 2000000 --steps 1 --batch-mode $M"
$CMD > gpurun_out/plain_k3_m$M.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:batch_scan -s 2 -c 1 -o gpurun_out/k3_m${M}_$TAG $CMD > gpurun_out/ncu_k3_m$M.log 2>&1
tail -c 300 gpurun_out/plain_k3_m$M.log; tail -3 gpurun_out/ncu_k3_m$M.log
done
