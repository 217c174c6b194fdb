#!/bin/bash
# Multi-GPU evidence (gpurun --gpus N): usage scripts/evidence_multi.sh <tag> <N>
T=${1:-r2m}; N=${2:-8}; O=gpurun_out; mkdir -p $O
run() {  # config, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    --steps 200 --warmup 20 --config $1 $2 > $O/${T}_bench_${N}gpu_$1.json 2> $O/${T}_${N}gpu_$1.err
  echo "$1 rc=$?"
}
nvidia-smi topo -m > $O/${T}_${N}gpu_topo.txt 2>&1
run normal65k
run normal1M
run actor262k
run defense65k
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --rollout-k 0 2>&1 | grep -E "NCCL INFO (Channel|Connected|comm|ncclComm|NVLS)|\"metric\"" | head -12 > $O/${T}_${N}gpu_nccl.txt
python - $O $T $N <<'PY'
import json, sys, glob
o, t, n = sys.argv[1:4]
for f in sorted(glob.glob(f"{o}/{t}_bench_{n}gpu_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.4g" % d["value"], "ms=%.4f" % d["ms_per_step"], "e2e=%.4g" % ((d.get("e2e") or {}).get("value", 0)),
              "rollout=%.4g" % ((d.get("rollout") or {}).get("value", 0)), "episodes", d["episode_stats"]["episodes"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace("_bench_", "_").replace(".json", ".err")).read()[-500:])
PY
