# usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shard_group or k4" > gpurun_out/pytest_shard.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_shard.log; tail -3 gpurun_out/pytest_shard.log
fi
for ex in ${EXCH:-fused nccl}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 300 --warmup 20 --exchange $ex > gpurun_out/bench_n${N}_$ex.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n${N}_$ex.log
python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_$ex.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N $ex", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "pipelined", (d["e2e"].get("pipelined") or {}).get("value"), "frac", round(d["roofline"]["frac"],3), d["config"]["exchange"], d["verified"], d.get("stream_self_consistent"), d["gpu_launches"])
PY
tail -3 gpurun_out/bench_n${N}_$ex.log | grep -v "^{" | cut -c1-400
done
