python scripts/parity_sweep.py 4096 200 > gpurun_out/parity_sweep_r1i.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1i.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1i.txt | head -3
{
scripts/ab_sweep.sh 65536 "HK_CLASS_WARPS=0" "HK_X=auto" "HK_TARGET_BLOCKS=148" "HK_TARGET_BLOCKS=132" "HK_TARGET_BLOCKS=120" "HK_CLASS_WARPS=8444" "HK_CLASS_WARPS=a444" "HK_CLASS_WARPS=c444" "HK_CLASS_WARPS=8555" "HK_CLASS_WARPS=a555" "HK_CLASS_WARPS=8444 HK_SLOW_BLOCK=384" "HK_CLASS_WARPS=8444 HK_CLASS_LANES=5544"
scripts/ab_sweep.sh 32768 "HK_CLASS_WARPS=0" "HK_X=auto" "HK_CLASS_WARPS=4222" "HK_CLASS_WARPS=4444" "HK_CLASS_WARPS=5333" "HK_CLASS_WARPS=0 HK_ENV_WARPS=3"
scripts/ab_sweep.sh 131072 "HK_CLASS_WARPS=0" "HK_X=auto" "HK_TARGET_BLOCKS=148"
scripts/ab_sweep.sh 4096 "HK_CLASS_WARPS=0" "HK_X=auto"
scripts/ab_sweep.sh 262144 "HK_CLASS_WARPS=0" "HK_X=auto"
} > gpurun_out/ab_r1j.txt 2>&1
cat gpurun_out/ab_r1j.txt
