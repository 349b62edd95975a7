set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_query_stream" > gpurun_out/pytest_k2a.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_k2a.log; tail -5 gpurun_out/pytest_k2a.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|rror" gpurun_out/pytest_k2a.log | head; nvidia-smi | head -15; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or shard" > gpurun_out/pytest_k2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k2.log; tail -3 gpurun_out/pytest_k2.log
VARIANTS=1400,1401,1400,1401 ROWS=10000,100000,1000000,1250000,10000000 timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_sweep_r2.log 2>&1; cat gpurun_out/k2_sweep_r2.log
