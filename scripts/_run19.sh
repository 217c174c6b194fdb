D=$PWD/hockey_env_b200
python scripts/parity_sweep.py 8192 400 > gpurun_out/parity_sweep_r1l.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1l.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1l.txt | head -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
scripts/ab_sweep.sh 65536 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
scripts/ab_sweep.sh 131072 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
scripts/ab_sweep.sh 32768 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
} > gpurun_out/ab_r1r.txt 2>&1; cat gpurun_out/ab_r1r.txt
python scripts/phase_cycles.py 65536 | tail -1
