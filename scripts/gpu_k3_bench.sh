mkdir -p gpurun_out
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 > gpurun_out/bench_batch.log 2>&1; echo "rc=$?" >> gpurun_out/bench_batch.log
timeout 600 python bench.py --workload batch --steps 5 --warmup 3 --nq 128 > gpurun_out/bench_batch128.log 2>&1
tail -c 2500 gpurun_out/bench_batch.log; tail -c 1200 gpurun_out/bench_batch128.log
