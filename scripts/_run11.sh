{
scripts/ab_sweep.sh 65536 "HK_X=auto" "HK_CLASS_WARPS=8443" "HK_CLASS_WARPS=8442" "HK_CLASS_WARPS=8434" "HK_CLASS_WARPS=8343" "HK_CLASS_WARPS=9443" "HK_CLASS_WARPS=8543" "HK_CLASS_WARPS=7443"
} > gpurun_out/ab_r1k.txt 2>&1
cat gpurun_out/ab_r1k.txt
python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1n.txt 2>&1
grep -A9 "  class 0" gpurun_out/lane_trace_r1n.txt | head -44
