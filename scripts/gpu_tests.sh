# full GPU suite + smoke + the K0 bench at two batch sizes
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
for n in 4096 16384; do
timeout 300 python bench.py --workload pool --steps 20 --warmup 3 --pool-texts $n > gpurun_out/bench_pool_$n.log 2>&1; echo "rc=$?" >> gpurun_out/bench_pool_$n.log
timeout 300 python bench.py --workload pool --steps 20 --warmup 3 --pool-texts $n --pool-full-mask > gpurun_out/bench_pool_full_$n.log 2>&1; echo "rc=$?" >> gpurun_out/bench_pool_full_$n.log
done
grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*\|"verified": [a-z]*\|"value": [0-9.]*' gpurun_out/bench_pool_*.log
