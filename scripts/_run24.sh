{
scripts/ab_sweep.sh 131072 "HK_X=default" "HK_TOUCH=1" "HK_CLASS_LANES=5555" "HK_CLASS_LANES=5554" "HK_X=default"
scripts/ab_sweep.sh 262144 "HK_X=default" "HK_TOUCH=1" "HK_CLASS_LANES=5544"
scripts/ab_sweep.sh 524288 "HK_X=default" "HK_TOUCH=0"
} > gpurun_out/ab_r1w.txt 2>&1; cat gpurun_out/ab_r1w.txt
