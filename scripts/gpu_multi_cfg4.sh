# usage: bash scripts/gpu_multi_cfg4.sh N   (run under gpurun --gpus N): BASELINE configs[3], 100M x 384 sharded
N=${1:-8}
mkdir -p gpurun_out
for rows in 100000000 10000000; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 200 --warmup 20 --rows $rows --exchange fused > gpurun_out/bench_n${N}_rows$rows.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n${N}_rows$rows.log
python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_rows$rows.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N rows=$rows", round(d["ms_per_step"],4),"ms", round(d["value"],1),"qps e2e", round(d["e2e"]["value"],1), "GB/s/GPU", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), d["config"]["exchange"], d["verified"], d["clocks"])
PY
tail -3 gpurun_out/bench_n${N}_rows$rows.log | grep -v "^{" | cut -c1-300
done
