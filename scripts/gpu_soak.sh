# stress the chained query streams (and, with N > 1, the fused exchange): long streams, small shards
N=${1:-1}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  if [ "$N" = "1" ]; then timeout 600 python bench.py --no-cpu-baseline --no-extra --repeats 2 "$@" > gpurun_out/soak_n${N}_$tag.log 2>&1
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --no-extra --repeats 2 "$@" > gpurun_out/soak_n${N}_$tag.log 2>&1; fi
  echo "rc=$?" >> gpurun_out/soak_n${N}_$tag.log
  python - <<PY
import json
for l in open("gpurun_out/soak_n${N}_$tag.log"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N $tag", round(d["ms_per_step"]*1e3,2),"us", round(d["value"],1),"qps  e2e", round(d["e2e"]["value"],1), "verified", d["verified"], "consistent", d["verification"]["stream_equals_precomputed_pool_results"], d["gpu_launches"])
PY
  tail -2 gpurun_out/soak_n${N}_$tag.log | grep -v "^{" | cut -c1-200
}
if [ "${SOAK_SET:-full}" = "short" ]; then      # two runs: tiny shards (exchange-bound) and the bench's own shard size
run small --rows 400000 --steps 50000 --warmup 20
run mid --rows 10000000 --steps 5000 --warmup 20
else
run tiny --rows 5000 --steps 100000 --warmup 20
run small --rows 100000 --steps 50000 --warmup 20
run k50 --rows 300000 --k 50 --steps 20000 --warmup 20
run mid --rows 1000000 --steps 20000 --warmup 20
run d768 --rows 500000 --dim 768 --k 100 --steps 10000 --warmup 20
fi
