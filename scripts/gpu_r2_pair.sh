# round 2: first runs of the CTA-pair K3 kernel, smallest case first, each step under its own timeout
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_k3_pair_kernel and pairs and 65-384" > gpurun_out/pytest_pair0.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_pair0.log
tail -15 gpurun_out/pytest_pair0.log
if [ $rc -ne 0 ]; then nvidia-smi > gpurun_out/nvsmi_after.log 2>&1; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "k3 or batch" > gpurun_out/pytest_k3.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_k3.log
tail -30 gpurun_out/pytest_k3.log
