# round 2: the driver's two commands (reference arm first), then a few standalone workloads
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_ref_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ref_n1.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench_n1.log
tail -c 600 gpurun_out/r2_ref_n1.log
python - <<'PY'
import json
for l in open("gpurun_out/r2_bench_n1.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"], 1),
              "verified", d["verified"], d["verification"], "launches", d["gpu_launches"], d["clocks"])
        print("repeats", d["repeats"])
        print("cpu", d.get("cpu_baseline"))
        for k, v in d.get("configs", {}).items():
            print("==", k, json.dumps(v)[:1500])
PY
tail -5 gpurun_out/r2_bench_n1.log | cut -c1-400
