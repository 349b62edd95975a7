mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "very_large_batches" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log; tail -15 gpurun_out/pytest_new.log
