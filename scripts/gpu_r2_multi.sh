# usage: bash scripts/gpu_r2_multi.sh N   (under gpurun --gpus N): the driver's command at N GPUs (reference arm first), the
# shard-group tests that need more than one device, and the driver-style line's multi-GPU extras
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shard_group or k4 or local_group" > gpurun_out/pytest_shard_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_shard_n$N.log; tail -3 gpurun_out/pytest_shard_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/r2_bench_n$N.log
python - <<PY
import json
for l in open("gpurun_out/r2_bench_n$N.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N=$N", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "pipelined", (d["e2e"].get("pipelined") or {}).get("value"), "frac", round(d["roofline"]["frac"],3), d["verified"], d["verification"], d["gpu_launches"], d["clocks"])
        print("per-rank local ms", d["setup"].get("per_rank_local_scan_ms"), "exchange cost", d["setup"].get("exchange_cost_ms_per_step"))
        print("single process", json.dumps(d["e2e"].get("single_process"))[:800])
        for k,v in d.get("configs",{}).items(): print("==",k, json.dumps(v)[:900])
PY
tail -3 gpurun_out/r2_bench_n$N.log | grep -v "^{" | cut -c1-400
# longer stream: the steady-state number (300 steps)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 300 --warmup 20 --no-extra > gpurun_out/r2_bench_n${N}_300.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/r2_bench_n${N}_300.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N=$N steps=300", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["verified"])
        print("per-rank local ms", d["setup"].get("per_rank_local_scan_ms"), "exchange cost", d["setup"].get("exchange_cost_ms_per_step"))
PY
