# persistent stream kernel (k2_stream.cuh): its tests first (short timeout: a hang must not hold the box), then every
# stream / shard test, then the A/B sweep: 902 = one chained launch per query, 901 = one persistent launch
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent" > gpurun_out/pytest_stream_a.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_stream_a.log; tail -5 gpurun_out/pytest_stream_a.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|rror" gpurun_out/pytest_stream_a.log | head -30; nvidia-smi | head -15; exit 0; fi
timeout 600 python -m pytest tests -x -q -m gpu -k "stream or shard or full_size" > gpurun_out/pytest_stream_b.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stream_b.log; tail -3 gpurun_out/pytest_stream_b.log
VARIANTS=902,901,902,901 ROWS=${ROWS:-100000,300000,1000000,1250000,10000000} timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_stream_sweep.log 2>&1; cat gpurun_out/k2_stream_sweep.log
K=100 VARIANTS=902,901,902,901 ROWS=1000000,1250000 timeout 600 python scripts/k2_sweep.py > gpurun_out/k2_stream_sweep_k100.log 2>&1; cat gpurun_out/k2_stream_sweep_k100.log
