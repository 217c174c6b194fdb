python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1i.txt 2>&1
HK_ENV_WARPS=5 HK_SLOW_BLOCK=160 python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1j.txt 2>&1
grep -A12 "blocks with work" gpurun_out/lane_trace_r1i.txt | head -60
echo ======
grep -A12 "blocks with work" gpurun_out/lane_trace_r1j.txt | head -30
