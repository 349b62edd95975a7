mkdir -p gpurun_out; : > gpurun_out/k3_lib_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k3 or batch" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log; tail -3 gpurun_out/pytest_k3.log
for rnd in 1 2; do for L in ${LIBS:-libsema_b200.so}; do
SEMA_B200_LIB=$PWD/sema_b200/$L timeout 200 python scripts/k3_time.py >> gpurun_out/k3_lib_ab.log 2>&1
done; done
cat gpurun_out/k3_lib_ab.log
timeout 300 python bench.py --workload batch --batch-mode 0 --steps 8 > gpurun_out/bb_r2_cascade.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bb_r2_cascade.log"):
    if l.startswith("{"):
        d=json.loads(l); print("cascade", round(d["ms_per_step"],3), "ms", round(d["roofline"]["frac"],3), d["clocks"], d["verified_against_k2"], d["batch"]["k3_fallback_queries"], "e2e", round(d["e2e"]["ms_per_step"],3))
PY
