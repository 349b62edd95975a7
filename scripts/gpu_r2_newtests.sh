mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "format_follows or repeated_batches or single_pass_kernels_follow or prefetch_knob" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log; tail -12 gpurun_out/pytest_new.log
