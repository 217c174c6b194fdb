#!/bin/bash
# A/B of library builds (build_variants/lib_*.so) on the GPU at 65,536 envs (+ optional sizes): scripts/ab_libs.sh TAG [envs...]
TAG=$1; shift
mkdir -p gpurun_out
for lib in build_variants/lib_*.so; do
  for n in 65536 "$@"; do
    HK_LIB_PATH=$PWD/$lib python bench.py --envs $n --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --rollout-k 0 > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
    python - "$lib" "$n" gpurun_out/${TAG}_tmp.json <<'PY' | tee -a gpurun_out/${TAG}.txt
import json,sys
try:
    d=json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print('%-28s n=%-8s value=%.4g ms=%.4f' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']), {k:round(v,4) for k,v in d['kernel_ms_per_tick'].items() if v > 0.01})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
  done
done
