# usage: bash scripts/gpu_r2_stream_multi.sh N (under gpurun --gpus N): the persistent stream kernel with the REAL peer
# exchange in its finisher warp — the driver's command, then two soaks on small shards (every result of the stream is
# compared with the pool query's precomputed result), each against the launch-per-query form
N=${1:-2}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/sm_bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sm_bench_n$N.log
$T --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --launch-per-query --no-cpu-baseline > gpurun_out/sm_bench_n${N}_q.log 2>&1; echo "rc=$?" >> gpurun_out/sm_bench_n${N}_q.log
$T --master-port 29513 bench.py --gpus $N --rows $((100000 * N)) --steps 5000 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/sm_soak_small_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sm_soak_small_n$N.log
$T --master-port 29514 bench.py --gpus $N --rows $((1250000 * N)) --steps 1000 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/sm_soak_shard_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sm_soak_shard_n$N.log
$T --master-port 29515 bench.py --gpus $N --rows $((1250000 * N)) --steps 1000 --warmup 20 --no-extra --no-cpu-baseline --launch-per-query > gpurun_out/sm_soak_shard_n${N}_q.log 2>&1; echo "rc=$?" >> gpurun_out/sm_soak_shard_n${N}_q.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/sm_*_n$N*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], round(d["ms_per_step"] * 1e3, 2), "us", round(d["value"], 1), "qps e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4),
                  "launches", d["gpu_launches"], "verified", d["verified"], d["verification"], d["setup"].get("per_rank_local_scan_ms"), d["setup"].get("exchange_cost_ms_per_step"))
            for k, v in d.get("configs", {}).items():
                if "100M" in k or "shard" in k: print("  ==", k, json.dumps(v)[:700])
    print(f.split("/")[-1], open(f).read().strip().split("\n")[-1][:200])
PY
