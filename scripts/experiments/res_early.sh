#!/bin/bash
# residual prefetch at the start of the tile (HRP_TC_RES_EARLY=1) against at the end (0)
mkdir -p gpurun_out; rm -f gpurun_out/r2_re_*.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_layer or fullnet_against_reference_golden or fullnet_tensor_core" > gpurun_out/r2_re_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_re_tests.log
tail -3 gpurun_out/r2_re_tests.log
SH="64,64,256,1,1,1 32,128,512,1,1,1 16,256,1024,1,1,1 8,512,2048,1,1,1 64,32,64,3,2,1 32,64,128,3,2,1"
for c in 0 1; do
  echo "== HRP_TC_RES_EARLY=$c" >> gpurun_out/r2_re_layers.txt
  HRP_TC_RES_EARLY=$c timeout 300 python scripts/conv_bench.py f16 64 $SH >> gpurun_out/r2_re_layers.txt 2>&1
done
for c in 0 1 0 1; do
  echo "== HRP_TC_RES_EARLY=$c" >> gpurun_out/r2_re_bench.txt
  HRP_TC_RES_EARLY=$c timeout 400 python bench.py --steps 12 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_re_bench.txt 2>&1
done
cat gpurun_out/r2_re_layers.txt gpurun_out/r2_re_bench.txt
