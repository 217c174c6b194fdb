python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1l.txt 2>&1
grep -B1 -A5 "  class 0" gpurun_out/lane_trace_r1l.txt | head -30
{
scripts/ab_sweep.sh 65536 "HK_CLASS_LANES=5555" "HK_CLASS_LANES=5544" "HK_CLASS_LANES=5533" "HK_CLASS_LANES=5543" "HK_CLASS_LANES=5534" "HK_CLASS_LANES=5433" "HK_CLASS_LANES=5545" "HK_CLASS_LANES=5554"
scripts/ab_sweep.sh 131072 "HK_CLASS_LANES=5555" "HK_CLASS_LANES=5544" "HK_CLASS_LANES=5533"
} > gpurun_out/ab_r1g.txt 2>&1
cat gpurun_out/ab_r1g.txt
