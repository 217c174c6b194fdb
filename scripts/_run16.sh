D=$PWD/hockey_env_b200
{
scripts/ab_sweep.sh 32768 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_varA.so" "HK_LIB_PATH=$D/libhockey_varB.so"
scripts/ab_sweep.sh 4096 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_varA.so" "HK_LIB_PATH=$D/libhockey_varB.so"
scripts/ab_sweep.sh 65536 "HK_LIB_PATH=$D/libhockey_prev.so" "HK_X=new" "HK_LIB_PATH=$D/libhockey_varA.so" "HK_LIB_PATH=$D/libhockey_varB.so"
} > gpurun_out/ab_r1o.txt 2>&1; cat gpurun_out/ab_r1o.txt
