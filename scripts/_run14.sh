python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1q.txt 2>&1
grep "finish phase" gpurun_out/lane_trace_r1q.txt
scripts/ab_sweep.sh 32768 "HK_X=1" "HK_X=2" "HK_CLASS_WARPS=0" "HK_CLASS_WARPS=4222" "HK_X=3" > gpurun_out/ab_r1m.txt 2>&1; cat gpurun_out/ab_r1m.txt
