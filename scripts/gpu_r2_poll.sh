mkdir -p gpurun_out; : > gpurun_out/e2e_poll_ab.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.limit,pcie.link.gen.current,pcie.link.width.current --format=csv >> gpurun_out/e2e_poll_ab.log
timeout 200 python scripts/e2e_poll_ab.py >> gpurun_out/e2e_poll_ab.log 2>&1
cat gpurun_out/e2e_poll_ab.log
