python scripts/parity_sweep.py 8192 300 > gpurun_out/parity_sweep_r1m.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1m.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1m.txt | head -3
{
scripts/ab_sweep.sh 65536 "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31" "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31"
scripts/ab_sweep.sh 32768 "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31"
scripts/ab_sweep.sh 131072 "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31"
scripts/ab_sweep.sh 262144 "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31" "HK_PHASE_SYNC=15" "HK_PHASE_SYNC=31"
} > gpurun_out/ab_r1t.txt 2>&1; cat gpurun_out/ab_r1t.txt
