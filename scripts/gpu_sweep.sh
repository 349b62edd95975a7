mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or k2_search" > gpurun_out/pytest_var.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_var.log; tail -3 gpurun_out/pytest_var.log
VARIANTS=0,2,3 timeout 900 python scripts/k2_sweep.py > gpurun_out/k2_sweep.log 2>&1; echo "rc=$?" >> gpurun_out/k2_sweep.log
grep -E "rows= *(1000|100000|1000000|5000000|10000000) " gpurun_out/k2_sweep.log; tail -2 gpurun_out/k2_sweep.log
