python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
scripts/ab_sweep.sh 65536 "HK_X=auto" "HK_CLASS_WARPS=8555" "HK_CLASS_WARPS=7444" "HK_CLASS_WARPS=9555" "HK_CLASS_WARPS=a555" "HK_CLASS_WARPS=7555" "HK_CLASS_WARPS=8454" "HK_CLASS_WARPS=8544" "HK_X=auto"
} > gpurun_out/ab_r1s.txt 2>&1; cat gpurun_out/ab_r1s.txt
python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1r.txt 2>&1
grep -A3 "  class 0" gpurun_out/lane_trace_r1r.txt | head -8; grep -A8 "blocks with work" gpurun_out/lane_trace_r1r.txt | head -20
