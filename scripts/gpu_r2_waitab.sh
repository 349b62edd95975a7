mkdir -p gpurun_out; : > gpurun_out/k3_wait_ab.log
for rnd in 1 2; do for L in libsema_b200.so libsema_b200_waitSPIN.so libsema_b200_waitNOHINT.so; do
SEMA_B200_LIB=$PWD/sema_b200/$L timeout 200 python scripts/k3_time.py >> gpurun_out/k3_wait_ab.log 2>&1
done; done
cat gpurun_out/k3_wait_ab.log
