set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k0 or k1_" > gpurun_out/pytest_k0.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k0.log
tail -15 gpurun_out/pytest_k0.log
timeout 300 python bench.py --workload pool --steps 20 --warmup 3 > gpurun_out/bench_pool.log 2>&1; echo "rc=$?" >> gpurun_out/bench_pool.log
timeout 300 python bench.py --workload pool --steps 20 --warmup 3 --pool-full-mask > gpurun_out/bench_pool_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_pool_full.log
tail -c 900 gpurun_out/bench_pool.log; tail -c 900 gpurun_out/bench_pool_full.log
