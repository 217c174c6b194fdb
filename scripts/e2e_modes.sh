timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "step_host" 2>&1 | tail -3
for m in overlap copy zero_copy; do
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --rollout-k 0 --e2e-mode $m > gpurun_out/r2h_e2e_$m.json 2> gpurun_out/r2h_err.txt
python - $m <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2h_e2e_{m}.json").read().strip().splitlines()[-1])
    print(m, "value=%.4g ms=%.4f e2e=%.4g e2e_ms=%.4f ratio=%.3f match=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"]/d["value"], d["e2e"]["replay_matches_recording"]))
except Exception as e:
    print(m, "ERR", e, open("gpurun_out/r2h_err.txt").read()[-600:])
PY
done
