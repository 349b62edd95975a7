# ncu --set full of the K3 kernels (pair single-pass, single-CTA single-pass, three-pass) on a 2M-row corpus; each capture only
# after the same command exited 0 without ncu
mkdir -p gpurun_out
TAG=${TAG:-r02b}
for V in "3 1 pair_scan" "3 0 batch_scan" "2 1 batch_scan"; do
set -- $V
CMD="python bench.py --workload batch --rows 2000000 --steps 1 --batch-mode $1 --k3-pair $2"
$CMD > gpurun_out/plain_k3_m$1_p$2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o gpurun_out/k3_m$1_p$2_$TAG -f $CMD > gpurun_out/ncu_k3_m$1_p$2.log 2>&1
tail -c 300 gpurun_out/plain_k3_m$1_p$2.log; tail -3 gpurun_out/ncu_k3_m$1_p$2.log
done
