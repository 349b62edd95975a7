# query streams (chained K2 launches) + host-query path: tests, then benches with comparisons
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or host_query or k2_ or shard_group or golden or empty or tombstone" > gpurun_out/pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stream.log
tail -15 gpurun_out/pytest_stream.log
for rows in 10000000 1000000; do
for fl in "" "--no-chain" "--no-stream" "--staged-host-path"; do
  tag=$(echo "$fl" | tr -d ' -')
  timeout 600 python bench.py --rows $rows --no-cpu-baseline $fl > gpurun_out/bs_${rows}_${tag:-default}.log 2>&1; echo "rc=$?" >> gpurun_out/bs_${rows}_${tag:-default}.log
  python - <<PY
import json
for l in open("gpurun_out/bs_${rows}_${tag:-default}.log"):
    if l.startswith("{"):
        d=json.loads(l); print("rows=$rows [$fl]", round(d["ms_per_step"]*1e3,2),"us", round(d["value"],1),"qps  e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],4), d["verified"], d["gpu_launches"])
PY
  tail -2 gpurun_out/bs_${rows}_${tag:-default}.log | grep -v "^{" | cut -c1-300
done
done
