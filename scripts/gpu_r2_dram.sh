# DRAM traffic of the K3 single-pass kernels on the full 10M x 384 corpus (do the query-group clusters still share L2?)
mkdir -p gpurun_out
for P in 1 0; do
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:"pair_scan|batch_scan" -s 2 -c 2 --csv --log-file gpurun_out/dram_p$P.csv python bench.py --workload batch --rows 10000000 --steps 1 --batch-mode 3 --k3-pair $P > gpurun_out/dram_p$P.log 2>&1
grep -E "pair_scan|batch_scan" gpurun_out/dram_p$P.csv | cut -d, -f5,13- | head -12
done
