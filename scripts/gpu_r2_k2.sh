set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_k2_variants_agree or stream" > gpurun_out/pytest_k2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k2.log; tail -3 gpurun_out/pytest_k2.log
VARIANTS=0,5,0,5 ROWS=10000,100000,1000000,1250000,10000000 timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_sweep_r2.log 2>&1; cat gpurun_out/k2_sweep_r2.log
