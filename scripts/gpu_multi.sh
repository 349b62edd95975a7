# usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n$N.log
tail -c 3000 gpurun_out/bench_n$N.log
