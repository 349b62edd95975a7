# usage: bash scripts/gpu_r2_stream_n8.sh N (under gpurun --gpus N): the driver's command with the persistent stream
# kernel, then the same region as one launch per query (no extras) for the A/B
N=${1:-8}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s8_bench_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/s8_bench_n$N.log
$T --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --launch-per-query --no-cpu-baseline > gpurun_out/s8_bench_n${N}_q.log 2>&1; echo "rc=$?" >> gpurun_out/s8_bench_n${N}_q.log
$T --master-port 29513 bench.py --gpus $N --steps 1000 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/s8_soak_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/s8_soak_n$N.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/s8_*_n$N*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], round(d["ms_per_step"] * 1e3, 2), "us", round(d["value"], 1), "qps e2e", round(d["e2e"]["value"], 1), "pipelined", (d["e2e"].get("pipelined") or {}).get("value"),
                  "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"], "verified", d["verified"], d["verification"], d["repeats"]["device_ms_per_step"]["all"])
            print("  per-rank local ms", d["setup"].get("per_rank_local_scan_ms"), "exchange cost", d["setup"].get("exchange_cost_ms_per_step"))
            print("  single process", json.dumps(d["e2e"].get("single_process"))[:600])
            for k, v in d.get("configs", {}).items():
                if "100M" in k or "shard" in k: print("  ==", k, json.dumps(v)[:500])
    print(f.split("/")[-1], open(f).read().strip().split("\n")[-1][:200])
PY
