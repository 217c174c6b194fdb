D=$PWD/hockey_env_b200
python scripts/parity_sweep.py 8192 300 > gpurun_out/parity_sweep_r1n.txt 2>&1; grep -c " ok " gpurun_out/parity_sweep_r1n.txt; grep -i "mismatch\|error" gpurun_out/parity_sweep_r1n.txt | head -3
{
scripts/ab_sweep.sh 65536 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
scripts/ab_sweep.sh 131072 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new"
} > gpurun_out/ab_r1u.txt 2>&1; cat gpurun_out/ab_r1u.txt
