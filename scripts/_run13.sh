python scripts/parity_sweep.py 8192 300 > gpurun_out/parity_sweep_r1j.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1j.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1j.txt | head -3
{
scripts/ab_sweep.sh 65536 "HK_X=1" "HK_X=2" "HK_X=3"
scripts/ab_sweep.sh 131072 "HK_X=1" 
scripts/ab_sweep.sh 32768 "HK_X=1" 
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_X=1" 
} > gpurun_out/ab_r1l.txt 2>&1
cat gpurun_out/ab_r1l.txt
python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1p.txt 2>&1
grep -A4 "blocks with work" gpurun_out/lane_trace_r1p.txt | head -30
grep "slowest unit is" gpurun_out/lane_trace_r1p.txt | sort | uniq -c | sort -rn | head -5
