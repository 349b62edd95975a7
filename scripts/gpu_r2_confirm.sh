# round 2 re-entry: confirm HEAD on a fresh box — full GPU suite, smoke, the driver's two bench arms (timed)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
S=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$? secs=$(( $(date +%s) - S ))" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
S=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1; echo "rc=$? secs=$(( $(date +%s) - S ))" >> gpurun_out/bench_reference.log
S=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; echo "rc=$? secs=$(( $(date +%s) - S ))" >> gpurun_out/bench_default.log
tail -c 600 gpurun_out/bench_reference.log
tail -c 6000 gpurun_out/bench_default.log
