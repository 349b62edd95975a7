set -x
mkdir -p gpurun_out
timeout 600 python scripts/k3_probe.py > gpurun_out/k3_probe_r2.log 2>&1; echo "rc=$?" >> gpurun_out/k3_probe_r2.log
cat gpurun_out/k3_probe_r2.log
