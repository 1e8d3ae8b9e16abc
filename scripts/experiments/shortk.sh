#!/bin/bash
# short-K / epilogue-bound layers at the quarter-GPU cap: two CTAs x four epilogue warps (rule) against one CTA x eight (deeper operand ring)
mkdir -p gpurun_out; rm -f gpurun_out/r2_sk_*.txt
SH="64,64,256,1,1,1 64,64,256,1,1,0 32,128,512,1,1,1 16,256,1024,1,1,1 64,256,64,1,1,0 64,256,128,1,1,0 8,512,2048,1,1,1 64,32,128,1,1,1 32,64,256,1,1,1 16,128,512,1,1,1"
run() { echo "== $*" >> gpurun_out/r2_sk_layers.txt; env HRP_TC_DEBUG=1 HRP_BENCH_PCT=25 "$@" timeout 300 python scripts/conv_bench.py f16 64 $SH 2>&1 | awk '!seen[$0]++' >> gpurun_out/r2_sk_layers.txt; }
run HRP_X=0
run HRP_TC_EPI=8
run HRP_TC_EPI=8 HRP_TC_BN=256
run HRP_TC_CTAS=1
run HRP_TC_NO_SHORT_STG2=1
cat gpurun_out/r2_sk_layers.txt
