python scripts/parity_sweep.py 8192 300 > gpurun_out/parity_sweep_r1g.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1g.txt; grep -i "mismatch\|bad\|error" gpurun_out/parity_sweep_r1g.txt | head; tail -1 gpurun_out/parity_sweep_r1g.txt
{
scripts/ab_sweep.sh 65536 "HK_ENV_WARPS=5 HK_SLOW_BLOCK=160 HK_PHASE_SYNC=15" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=160" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=256" "HK_ENV_WARPS=5 HK_SLOW_BLOCK=384"  "HK_ENV_WARPS=5 HK_SLOW_BLOCK=384 HK_PHASE_SYNC=15" "HK_ENV_WARPS=6 HK_SLOW_BLOCK=384" "HK_ENV_WARPS=6 HK_SLOW_BLOCK=192" "HK_ENV_WARPS=8 HK_SLOW_BLOCK=384"
scripts/ab_sweep.sh 131072 "HK_ENV_WARPS=12 HK_PHASE_SYNC=15" "HK_ENV_WARPS=12" "HK_ENV_WARPS=10" "HK_ENV_WARPS=8"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_ENV_WARPS=12 HK_PHASE_SYNC=15" "HK_ENV_WARPS=12"
} > gpurun_out/ab_r1f.txt 2>&1
cat gpurun_out/ab_r1f.txt
python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1k.txt 2>&1
grep -A9 "blocks with work" gpurun_out/lane_trace_r1k.txt | head -22
