#!/bin/bash
# End-of-round evidence on one B200 (run under gpurun): latency table, ncu launch list of one forward of the bench
# configuration (+ L2 / DRAM traffic), full ncu captures of the top kernels, the two HBM-bound kernels, and the
# families' deviation on un-damped weights. Outputs under gpurun_out/r02_final_*.
OUT=gpurun_out
mkdir -p $OUT
python scripts/latency.py > $OUT/r02_final_latency.txt 2>&1
python scripts/one_step.py f16 > $OUT/r02_final_one_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --csv --log-file $OUT/r02_final_launches.csv python scripts/one_step.py f16 > $OUT/r02_final_ncu_launches.log 2>&1
python scripts/hbm_kernels.py softargmax > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:softargmax_partial --launch-skip 2 -c 1 -f -o $OUT/r02_final_softargmax python scripts/hbm_kernels.py softargmax > $OUT/r02_final_ncu_sa.log 2>&1
python scripts/hbm_kernels.py fk > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fk_gen --launch-skip 2 -c 1 -f -o $OUT/r02_final_fk python scripts/hbm_kernels.py fk > $OUT/r02_final_ncu_fk.log 2>&1
python bench.py > $OUT/r02_final_bench.json 2> $OUT/r02_final_bench.err
python bench.py --backbone hrnet32 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r02_final_bench_hrnet32.json 2> $OUT/r02_final_bench_hrnet32.err
