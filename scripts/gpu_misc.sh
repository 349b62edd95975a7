mkdir -p gpurun_out
timeout 300 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -5 gpurun_out/sanitize_plain.log; tail -12 gpurun_out/sanitize_memcheck.log
timeout 600 python bench.py --workload config1 > gpurun_out/bench_config1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_config1.log
tail -c 1800 gpurun_out/bench_config1.log
