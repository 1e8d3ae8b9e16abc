#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_t64_*.txt
SH="64,32,32,3,2,0 64,32,64,3,2,1 32,32,128,3,2,1 32,32,32,3,2,0 16,32,256,3,2,1 64,32,32,1,1,0 64,32,128,1,1,1 64,32,32,3,1,0"
for c in 0 1 2; do
  echo "== HRP_TC_NO_TMA64=$c (quarter GPU)" >> gpurun_out/r2_t64_layers.txt
  HRP_TC_DEBUG=1 HRP_BENCH_PCT=25 HRP_TC_NO_TMA64=$c timeout 300 python scripts/conv_bench.py f16 64 $SH 2>&1 | awk '!seen[$0]++' >> gpurun_out/r2_t64_layers.txt
done
timeout 600 env HRP_TC_NO_TMA64=1 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_layer or fullnet_against_reference_golden" > gpurun_out/r2_t64_tests.log 2>&1; tail -2 gpurun_out/r2_t64_tests.log
for c in 0 1 2 0 1 2; do
  echo "== HRP_TC_NO_TMA64=$c" >> gpurun_out/r2_t64_bench.txt
  HRP_TC_NO_TMA64=$c timeout 300 python bench.py --steps 12 --warmup 4 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_t64_bench.txt 2>&1
done
grep -v "^conv_tc" gpurun_out/r2_t64_layers.txt; cat gpurun_out/r2_t64_bench.txt
