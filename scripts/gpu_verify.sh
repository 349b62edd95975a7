# quick end-of-session verification: full GPU suite, smoke, default bench + reference arm
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log
tail -c 3000 gpurun_out/bench_default.log; tail -c 400 gpurun_out/bench_reference.log
