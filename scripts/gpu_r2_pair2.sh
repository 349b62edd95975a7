# round 2: K3 tests (all), pair-kernel probes, batch bench A/B of the single-pass variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "k3 or batch" > gpurun_out/pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k3.log
tail -8 gpurun_out/pytest_k3.log
timeout 600 python scripts/k3_probe.py > gpurun_out/k3_probe_r2.log 2>&1; echo "rc=$?" >> gpurun_out/k3_probe_r2.log
cat gpurun_out/k3_probe_r2.log
for P in 1 2 0; do
timeout 300 python bench.py --workload batch --batch-mode 3 --k3-pair $P --steps 10 > gpurun_out/bb_r2_m3_p$P.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bb_r2_m3_p$P.log"):
    if l.startswith("{"):
        d=json.loads(l); print("pair=$P", d["batch"]["mode"], round(d["ms_per_step"],3), "ms", round(d["roofline"]["frac"],3), d["clocks"], d["verified_against_k2"], d["batch"]["k3_fallback_queries"])
PY
done
for M in 2 0; do for PR in 0 1; do
timeout 300 python bench.py --workload batch --batch-mode $M --precision $PR --steps 8 > gpurun_out/bb_r2_m${M}_prec$PR.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/bb_r2_m${M}_prec$PR.log"):
    if l.startswith("{"):
        d=json.loads(l); print("mode=$M prec=$PR", d["batch"]["mode"], round(d["ms_per_step"],3), "ms", round(d["roofline"]["frac"],3), d["clocks"], d["verified_against_k2"], d["batch"]["k3_fallback_queries"])
PY
done; done
