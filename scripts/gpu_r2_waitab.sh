mkdir -p gpurun_out; : > gpurun_out/k3_lib_ab.log
for rnd in 1 2 3; do for L in ${LIBS:-libsema_b200.so libsema_b200_heap16.so}; do
SEMA_B200_LIB=$PWD/sema_b200/$L timeout 200 python scripts/k3_time.py >> gpurun_out/k3_lib_ab.log 2>&1
done; done
cat gpurun_out/k3_lib_ab.log
