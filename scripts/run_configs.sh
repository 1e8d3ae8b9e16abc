#!/bin/bash
# BASELINE.json configs 3-5 at N GPUs of one box (run under `gpurun --gpus N`): Kuka batch 256 per GPU with the NCCL output
# gather, Baxter batch 128 per GPU (N = 8: 1024 frames in total), and the FK + projection sweep with and without the
# output all-gather. One JSON (line) file per run under gpurun_out/.   usage: scripts/run_configs.sh N [tag]
N=${1:-1}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
run() {   # name, then bench.py arguments
  local name=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  else
    timeout 600 python bench.py --gpus 1 "$@" > $OUT/${TAG}_${name}_n$N.json 2> $OUT/${TAG}_${name}_n$N.err
  fi
}
run kuka_b256 --robot kuka --batch 256 --steps 10 --warmup 3 --no-cpu-baseline
if [ "$N" = "8" ]; then run baxter_b128 --robot baxter --batch 128 --steps 10 --warmup 3 --no-cpu-baseline; fi
if [ "$N" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/fk_sweep.py --gather --iters 10 > $OUT/${TAG}_fk_sweep_n$N.jsonl 2> $OUT/${TAG}_fk_sweep_n$N.err
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_smi_n$N.txt 2>&1
