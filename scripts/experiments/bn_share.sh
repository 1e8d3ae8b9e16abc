#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_bs_*.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fullnet_against_reference_golden or fullnet_tensor_core or config_batch" > gpurun_out/r2_bs_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_bs_tests.log
tail -3 gpurun_out/r2_bs_tests.log
for c in "HRP_TC_BN_FULL_GPU=1" "HRP_TC_BN_FULL_GPU=0" "HRP_TC_CTAS_SHARE=1" "HRP_TC_BN_FULL_GPU=1" "HRP_TC_BN_FULL_GPU=0" "HRP_TC_CTAS_SHARE=1"; do
  echo "== $c" >> gpurun_out/r2_bs_bench.txt
  env $c timeout 400 python bench.py --steps 12 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_bs_bench.txt 2>&1
done
cat gpurun_out/r2_bs_bench.txt
