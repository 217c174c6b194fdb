HK_CLASS_WARPS=8333 python scripts/parity_sweep.py 4096 200 > gpurun_out/parity_sweep_r1h.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1h.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1h.txt | head -3
{
scripts/ab_sweep.sh 65536 "HK_CLASS_WARPS=0" "HK_CLASS_WARPS=5555" "HK_CLASS_WARPS=6444" "HK_CLASS_WARPS=7444" "HK_CLASS_WARPS=8444" "HK_CLASS_WARPS=8333" "HK_CLASS_WARPS=a333" "HK_CLASS_WARPS=6333" "HK_CLASS_WARPS=5444" "HK_CLASS_WARPS=5333" "HK_CLASS_WARPS=6433" "HK_CLASS_WARPS=6343"
scripts/ab_sweep.sh 131072 "HK_CLASS_WARPS=0" "HK_CLASS_WARPS=cccc" "HK_CLASS_WARPS=c888" "HK_CLASS_WARPS=c666" "HK_CLASS_WARPS=caaa" "HK_CLASS_WARPS=c888 HK_CLASS_LANES=5555"
} > gpurun_out/ab_r1i.txt 2>&1
cat gpurun_out/ab_r1i.txt
