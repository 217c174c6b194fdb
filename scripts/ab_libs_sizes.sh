for lib in build_variants/lib_*.so; do
  for n in 131072 262144 524288; do
    HK_LIB_PATH=$PWD/$lib python bench.py --envs $n --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --rollout-k 0 > gpurun_out/r2j_tmp.json 2> gpurun_out/r2j_tmp.err
    python - "$lib" "$n" gpurun_out/r2j_tmp.json <<'PY' | tee -a gpurun_out/r2j_optlevel.txt
import json,sys
try:
    d=json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print('%-28s n=%-8s value=%.4g ms=%.4f' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']), {k:round(v,4) for k,v in d['kernel_ms_per_tick'].items() if v > 0.01})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
  done
done
