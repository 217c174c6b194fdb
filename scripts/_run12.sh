python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1o.txt 2>&1
grep -A17 "blocks with work" gpurun_out/lane_trace_r1o.txt | head -80
