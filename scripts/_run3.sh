python scripts/parity_sweep.py 8192 300 > gpurun_out/parity_sweep_r1e.txt 2>&1; tail -4 gpurun_out/parity_sweep_r1e.txt
HK_ENV_WARPS=3 python scripts/parity_sweep.py 4096 200 > gpurun_out/parity_sweep_r1f.txt 2>&1; tail -2 gpurun_out/parity_sweep_r1f.txt
{
scripts/ab_sweep.sh 65536 "HK_ENV_WARPS=5 HK_SLOW_BLOCK=160" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=160 HK_PHASE_SYNC=7" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=256" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=384" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=384 HK_PHASE_SYNC=7" "HK_ENV_WARPS=6 HK_SLOW_BLOCK=384" "HK_ENV_WARPS=4 HK_SLOW_BLOCK=384" "HK_ENV_WARPS=3 HK_SLOW_BLOCK=384" "HK_ENV_WARPS=8 HK_SLOW_BLOCK=384"
scripts/ab_sweep.sh 131072 "HK_ENV_WARPS=12" "HK_ENV_WARPS=12 HK_PHASE_SYNC=7" "HK_ENV_WARPS=10" "HK_ENV_WARPS=8" "HK_ENV_WARPS=6"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_ENV_WARPS=12" "HK_ENV_WARPS=12 HK_PHASE_SYNC=7" "HK_ENV_WARPS=10" "HK_ENV_WARPS=8"
} > gpurun_out/ab_r1e.txt 2>&1
cat gpurun_out/ab_r1e.txt
