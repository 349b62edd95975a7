set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_full_size_gpu.py -x -q -m gpu --durations=8 > gpurun_out/pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full.log
tail -25 gpurun_out/pytest_full.log
