# usage: bash scripts/gpu_batch_multi.sh N  (under gpurun --gpus N): sharded batched search (config 3 over N GPUs)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched_virtual or sharded_searcher or k4" > gpurun_out/pytest_bshard.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bshard.log; tail -4 gpurun_out/pytest_bshard.log
fi
for mode in 2 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --workload batch --batch-mode $mode --steps 10 --warmup 3 > gpurun_out/bench_batch_n${N}_m$mode.log 2>&1
echo "rc=$?" >> gpurun_out/bench_batch_n${N}_m$mode.log
python - <<PY
import json
for l in open("gpurun_out/bench_batch_n${N}_m$mode.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N mode=$mode", round(d["ms_per_step"],3),"ms", round(d["value"]),"qps e2e", round(d["e2e"]["value"]), "TF/GPU", round(d["roofline"]["achieved"],1), d["verified"], d["gpu_launches"], d["clocks"])
PY
tail -3 gpurun_out/bench_batch_n${N}_m$mode.log | grep -v "^{" | cut -c1-300
done
