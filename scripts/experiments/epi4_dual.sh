#!/bin/bash
# layers the rule gives eight epilogue warps and the SM to themselves: two CTAs x (4 + 4 dual) epilogue warps instead?
mkdir -p gpurun_out; rm -f gpurun_out/r2_e4_*.txt
SH="64,32,128,1,1,1 64,32,128,1,1,0 64,256,128,1,1,0 16,256,1024,1,1,1 8,512,2048,1,1,1 16,128,512,1,1,1 16,128,512,1,1,0 8,256,1024,1,1,1 8,256,1024,1,1,0 32,64,256,1,1,1 32,64,256,1,1,0 32,512,256,1,1,0 16,1024,512,1,1,0 8,2048,512,1,1,0 32,256,512,3,2,1"
run() { echo "== $*" >> gpurun_out/r2_e4_layers.txt; env HRP_TC_DEBUG=1 HRP_BENCH_PCT=25 "$@" timeout 300 python scripts/conv_bench.py f16 64 $SH 2>&1 | awk '!seen[$0]++' >> gpurun_out/r2_e4_layers.txt; }
run HRP_X=0
run HRP_TC_EPI=4
run HRP_TC_EPI=4 HRP_TC_BN=128
cat gpurun_out/r2_e4_layers.txt
for c in 0 1 0 1 0 1; do
  echo "== HRP_TC_EPI_DUAL=$c" >> gpurun_out/r2_e4_bench.txt
  HRP_TC_EPI_DUAL=$c timeout 300 python bench.py --steps 60 --warmup 6 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" >> gpurun_out/r2_e4_bench.txt 2>&1
done
cat gpurun_out/r2_e4_bench.txt
