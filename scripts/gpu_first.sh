set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in 0 1 2; do
  timeout 600 python bench.py --steps 200 --warmup 20 --variant $v --no-cpu-baseline > gpurun_out/bench_v$v.log 2>&1
done
timeout 600 python bench.py --steps 200 --warmup 20 --rows 1000000 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
tail -3 gpurun_out/*.log
