# persistent stream kernel: driver-shaped bench lines (persistent vs one launch per query, alternating), then the ncu
# launch list and one full capture of scan_stream_kernel — each ncu run only after the plain command exited 0
set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
for i in 1 2; do
  timeout 300 $B > gpurun_out/stream_bench_p$i.log 2>&1; echo "rc=$?" >> gpurun_out/stream_bench_p$i.log
  timeout 300 $B --launch-per-query > gpurun_out/stream_bench_q$i.log 2>&1; echo "rc=$?" >> gpurun_out/stream_bench_q$i.log
done
for r in 1000000 1250000; do for i in 1 2; do
  timeout 300 $B --rows $r --no-verify --steps 200 > gpurun_out/stream_bench_${r}_p$i.log 2>&1
  timeout 300 $B --rows $r --no-verify --steps 200 --launch-per-query > gpurun_out/stream_bench_${r}_q$i.log 2>&1
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/stream_bench_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], round(d["ms_per_step"] * 1e3, 2), "us", round(d["value"], 1), "qps frac", round(d["roofline"]["frac"], 4),
                  "launches", d["gpu_launches"], "verified", d["verified"], d["repeats"]["device_ms_per_step"]["all"], d["clocks"]["sm_mhz"])
PY
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-verify"
$CMD > gpurun_out/plain_stream.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_stream.csv $CMD > gpurun_out/ncu_l_stream.log 2>&1
$CMD > gpurun_out/plain_stream_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 2 -c 1 -o gpurun_out/k2_stream_r02 $CMD > gpurun_out/ncu_f_stream.log 2>&1
tail -3 gpurun_out/ncu_f_stream.log
grep -c scan_stream gpurun_out/launches_stream.csv
