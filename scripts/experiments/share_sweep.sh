#!/bin/bash
# lane share per launch / plans in flight after the tile-width rule (f16, whole network)
mkdir -p gpurun_out; rm -f gpurun_out/r2_ss_bench.txt
for c in "HRP_PCT_HI=25 HRP_PCT_LO=25" "HRP_PCT_HI=20 HRP_PCT_LO=20" "HRP_PCT_HI=33 HRP_PCT_LO=33" "HRP_PCT_HI=33 HRP_PCT_LO=20" "HRP_PCT_HI=25 HRP_PCT_LO=25 HRP_SLOTS=5" "HRP_PCT_HI=20 HRP_PCT_LO=20 HRP_SLOTS=5" "HRP_PCT_HI=25 HRP_PCT_LO=25" "HRP_PCT_HI=33 HRP_PCT_LO=33 HRP_SLOTS=3"; do
  echo "== $c" >> gpurun_out/r2_ss_bench.txt
  env $c timeout 300 python bench.py --steps 12 --warmup 4 --no-families --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/r2_ss_bench.txt 2>&1
done
cat gpurun_out/r2_ss_bench.txt
