#!/bin/bash
# A/B of execution-strategy switches at several batch sizes: scripts/ab_env_sizes.sh TAG "sizes" "VAR=val ..." ["VAR=val ..." ...]
TAG=$1; SIZES=$2; shift; shift
mkdir -p gpurun_out
for n in $SIZES; do
  for v in "" "$@"; do
    env $v python bench.py --envs $n --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --rollout-k 0 > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
    echo -n "n=$n [${v:-default}] " | tee -a gpurun_out/${TAG}.txt
    python scripts/print_bench.py gpurun_out/${TAG}_tmp.json | tee -a gpurun_out/${TAG}.txt
  done
done
