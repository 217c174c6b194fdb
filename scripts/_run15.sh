P=$PWD/hockey_env_b200/libhockey_prev.so
{
scripts/ab_sweep.sh 32768 "HK_LIB_PATH=$P" "HK_X=new" "HK_LIB_PATH=$P" "HK_X=new"
scripts/ab_sweep.sh 65536 "HK_LIB_PATH=$P" "HK_X=new" "HK_LIB_PATH=$P" "HK_X=new"
scripts/ab_sweep.sh 4096 "HK_LIB_PATH=$P" "HK_X=new"
scripts/ab_sweep.sh 131072 "HK_LIB_PATH=$P" "HK_X=new"
} > gpurun_out/ab_r1n.txt 2>&1; cat gpurun_out/ab_r1n.txt
