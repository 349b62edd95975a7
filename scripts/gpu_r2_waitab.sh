mkdir -p gpurun_out; : > gpurun_out/k3_lib_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_k3_batch_matches or test_k3_pair_kernel or tombstones" > gpurun_out/pytest_pair.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair.log; tail -3 gpurun_out/pytest_pair.log
for rnd in 1 2 3; do for L in ${LIBS:-libsema_b200.so libsema_b200_noreg.so}; do
SEMA_B200_LIB=$PWD/sema_b200/$L timeout 200 python scripts/k3_time.py >> gpurun_out/k3_lib_ab.log 2>&1
done; done
cat gpurun_out/k3_lib_ab.log
