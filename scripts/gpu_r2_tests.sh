# round 2: full GPU suite + smoke (the new sequential-order K1 is exercised by every test that appends)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
