# compaction rewrite (direct gather once a chunk's sources lie beyond its destination, one synchronise) + the
# maintenance workload of bench.py (SURVEY 8(f) rows 1-2 measured)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "compact or tombstone or save or store" --durations=5 > gpurun_out/pytest_maint.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_maint.log; tail -12 gpurun_out/pytest_maint.log
timeout 600 python bench.py --workload maintenance > gpurun_out/bench_maint.log 2>&1; echo "rc=$?" >> gpurun_out/bench_maint.log; tail -c 3500 gpurun_out/bench_maint.log
