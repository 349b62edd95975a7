set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_k3_batch_matches" > gpurun_out/pytest_pair.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_pair.log
tail -3 gpurun_out/pytest_pair.log
timeout 600 python scripts/k3_probe.py > gpurun_out/k3_probe_r2.log 2>&1; echo "rc=$?" >> gpurun_out/k3_probe_r2.log
cat gpurun_out/k3_probe_r2.log
