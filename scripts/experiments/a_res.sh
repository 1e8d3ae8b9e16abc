#!/bin/bash
# A/B of the activation-resident walk of the fused heatmap head (HRP_TC_A_RES=0|1)
mkdir -p gpurun_out; rm -f gpurun_out/r2_ar_*
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_heatmap or fullnet_against_reference_golden or fullnet_tensor_core_families or config_batch or batch64" > gpurun_out/r2_ar_tests.log 2>&1; tail -2 gpurun_out/r2_ar_tests.log
for c in 0 1; do
  HRP_TC_A_RES=$c timeout 120 python scripts/dump_ops.py gpurun_out/r2_ar_ops$c.csv f16 > /dev/null 2>&1
  echo "A_RES=$c: $(grep ',64,64,256,64,64,448,' gpurun_out/r2_ar_ops$c.csv)"
done
for c in 0 1 0 1; do
  echo "== HRP_TC_A_RES=$c" >> gpurun_out/r2_ar_bench.txt
  HRP_TC_A_RES=$c timeout 200 python bench.py --steps 20 --warmup 5 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" >> gpurun_out/r2_ar_bench.txt 2>&1
done
cat gpurun_out/r2_ar_bench.txt
