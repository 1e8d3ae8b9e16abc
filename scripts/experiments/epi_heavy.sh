#!/bin/bash
# A/B of the epilogue-heavy rule (two CTAs x 128-wide tiles x two epilogue groups) and the second group in the fused heatmap head
mkdir -p gpurun_out; rm -f gpurun_out/r2_eh_*.txt
SH="64,32,128,1,1,1 64,32,128,1,1,0 64,256,128,1,1,0 16,256,1024,1,1,1 8,512,2048,1,1,1 8,256,1024,1,1,1 16,128,512,1,1,1 32,64,256,1,1,1 64,64,256,1,1,1 32,128,512,1,1,1 32,256,128,1,1,0 16,256,128,1,1,0"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_layer_tensor_core_families or fullnet_against_reference_golden or fullnet_tensor_core_families or fused_basic_block or branch_chain or fused_heatmap or config_batch" > gpurun_out/r2_eh_tests.log 2>&1; tail -2 gpurun_out/r2_eh_tests.log
for c in 0 1; do
  echo "== HRP_TC_EPI_HEAVY=$c (quarter GPU)" >> gpurun_out/r2_eh_layers.txt
  HRP_TC_DEBUG=1 HRP_BENCH_PCT=25 HRP_TC_EPI_HEAVY=$c timeout 300 python scripts/conv_bench.py f16 64 $SH 2>&1 | awk '!seen[$0]++' >> gpurun_out/r2_eh_layers.txt
done
for c in 0 1 0 1 0 1; do
  echo "== all three switches = $c" >> gpurun_out/r2_eh_bench.txt
  HRP_TC_EPI_DUAL=$c HRP_TC_EPI_HEAVY=$c HRP_TC_SA_DUAL=$c timeout 300 python bench.py --steps 60 --warmup 6 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" >> gpurun_out/r2_eh_bench.txt 2>&1
done
for c in "HRP_TC_SA_DUAL=0" "HRP_TC_EPI_HEAVY=0"; do
  echo "== $c (others on)" >> gpurun_out/r2_eh_bench.txt
  env $c timeout 300 python bench.py --steps 60 --warmup 6 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" >> gpurun_out/r2_eh_bench.txt 2>&1
done
grep -v "^conv_tc" gpurun_out/r2_eh_layers.txt; cat gpurun_out/r2_eh_bench.txt
