D=$PWD/hockey_env_b200
{
scripts/ab_sweep.sh 32768 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_CARVEOUT=0" "HK_CARVEOUT=1"
scripts/ab_sweep.sh 4096 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_CARVEOUT=0" "HK_CARVEOUT=1"
scripts/ab_sweep.sh 65536 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_CARVEOUT=0" "HK_CARVEOUT=1"
scripts/ab_sweep.sh 131072 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_CARVEOUT=0" "HK_CARVEOUT=1"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_CARVEOUT=0" "HK_CARVEOUT=1"
} > gpurun_out/ab_r1p.txt 2>&1; cat gpurun_out/ab_r1p.txt
