mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -40 gpurun_out/pytest_k3.log
