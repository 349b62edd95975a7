# where the persistent stream kernel starts to pay, by k (901 = persistent, 902 = one chained launch per query)
mkdir -p gpurun_out
K=10 VARIANTS=902,901,902,901 ROWS=10000,25000,50000 timeout 300 python scripts/k2_sweep.py > gpurun_out/thr_k10.log 2>&1; cat gpurun_out/thr_k10.log
K=32 VARIANTS=902,901,902,901 ROWS=100000,300000,600000 timeout 300 python scripts/k2_sweep.py > gpurun_out/thr_k32.log 2>&1; cat gpurun_out/thr_k32.log
K=50 VARIANTS=902,901,902,901 ROWS=100000,300000,600000 timeout 300 python scripts/k2_sweep.py > gpurun_out/thr_k50.log 2>&1; cat gpurun_out/thr_k50.log
K=100 VARIANTS=902,901,902,901 ROWS=100000,300000,600000 timeout 300 python scripts/k2_sweep.py > gpurun_out/thr_k100.log 2>&1; cat gpurun_out/thr_k100.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
