O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 300 --no-cpu-baseline > $O/bench_r1d_8gpu.json 2> $O/bench_r1d_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --envs 131072 --steps 200 --warmup 300 --no-cpu-baseline > $O/bench_r1d_8gpu_1M_total.json 2> $O/bench_r1d_8gpu_1M_total.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --envs 32768 --steps 200 --warmup 300 --no-cpu-baseline > $O/bench_r1d_8gpu_262k_total.json 2> $O/bench_r1d_8gpu_262k_total.err
for f in 8gpu 8gpu_1M_total 8gpu_262k_total; do tail -1 $O/bench_r1c_$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
tail -3 $O/bench_r1d_8gpu.err
