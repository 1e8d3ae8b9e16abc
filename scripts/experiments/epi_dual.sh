#!/bin/bash
# A/B of the second epilogue group in conv_tc (HRP_TC_EPI_DUAL=0|1): layer times at the quarter-GPU cap, tests, whole network
mkdir -p gpurun_out; rm -f gpurun_out/r2_ed_*.txt
SH="64,64,256,1,1,1 64,64,256,1,1,0 32,128,512,1,1,1 16,256,1024,1,1,1 64,256,64,1,1,0 64,256,128,1,1,0 16,1024,256,1,1,0 8,512,2048,1,1,1 64,32,128,1,1,1 16,256,256,3,1,0 32,512,128,1,1,0 64,128,256,3,2,1 32,64,32,1,1,0 64,32,32,1,1,0 64,256,32,3,1,0 64,32,64,3,2,1 64,32,32,3,2,0 128,64,64,3,2,0 32,128,128,3,1,0"
timeout 120 python scripts/conv_bench.py f16 64 64,64,256,1,1,1 64,256,64,1,1,0 > gpurun_out/r2_ed_first.txt 2>&1 || { echo "first run failed"; tail -5 gpurun_out/r2_ed_first.txt; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_layer_tensor_core_families or fullnet_against_reference_golden or fullnet_tensor_core_families or fused_basic_block or branch_chain" > gpurun_out/r2_ed_tests.log 2>&1; tail -2 gpurun_out/r2_ed_tests.log
for c in 0 1; do
  echo "== HRP_TC_EPI_DUAL=$c (quarter GPU)" >> gpurun_out/r2_ed_layers.txt
  HRP_BENCH_PCT=25 HRP_TC_EPI_DUAL=$c timeout 300 python scripts/conv_bench.py f16 64 $SH 2>&1 >> gpurun_out/r2_ed_layers.txt
done
for c in 0 1 0 1; do
  echo "== HRP_TC_EPI_DUAL=$c" >> gpurun_out/r2_ed_bench.txt
  HRP_TC_EPI_DUAL=$c timeout 300 python bench.py --steps 12 --warmup 4 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_ed_bench.txt 2>&1
done
cat gpurun_out/r2_ed_layers.txt gpurun_out/r2_ed_bench.txt
