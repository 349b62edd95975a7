mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_store_gpu.py -x -q -m gpu > gpurun_out/pytest_store.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_store.log; tail -15 gpurun_out/pytest_store.log
