set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_store_gpu.py -x -q -m gpu -k "not k3 and not k0" > gpurun_out/pytest_k2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k2.log
tail -5 gpurun_out/pytest_k2.log
for K in 10 50 100; do K=$K timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_sweep_k$K.log 2>&1; echo "k=$K"; cat gpurun_out/k2_sweep_k$K.log; done
