#!/bin/bash
# weight-tile multicast across thread-block clusters in conv_tc: parity, single layers, whole network
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_layer or fused_heatmap or fullnet_against_reference_golden or fullnet_tensor_core" > gpurun_out/r2_cl_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_cl_tests.log
tail -3 gpurun_out/r2_cl_tests.log
SH="64,64,256,1,1,1 64,256,64,1,1,0 32,128,512,1,1,1 16,256,1024,1,1,1 32,128,128,3,1,0 16,256,256,3,1,0 8,512,512,3,1,0 16,1024,256,1,1,0 8,2048,512,1,1,0 8,512,2048,1,1,1 32,256,512,3,2,0"
for c in 1 2 4; do
  echo "== HRP_TC_CLUSTER=$c" >> gpurun_out/r2_cl_layers.txt
  HRP_TC_CLUSTER=$c timeout 300 python scripts/conv_bench.py f16 64 $SH >> gpurun_out/r2_cl_layers.txt 2>&1
done
for c in 1 2 4 2 1; do
  echo "== HRP_TC_CLUSTER=$c" >> gpurun_out/r2_cl_bench.txt
  HRP_TC_CLUSTER=$c timeout 400 python bench.py --steps 12 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_cl_bench.txt 2>&1
done
cat gpurun_out/r2_cl_layers.txt gpurun_out/r2_cl_bench.txt
