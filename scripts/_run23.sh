python -m pytest tests -m gpu -x -q -k "variants" 2>&1 | tail -2
{
scripts/ab_sweep.sh 65536 "HK_FAST_BLOCK=128" "HK_FAST_BLOCK=64" "HK_FAST_BLOCK=96" "HK_FAST_BLOCK=32" "HK_FAST_BLOCK=128" "HK_FAST_BLOCK=64"
scripts/ab_sweep.sh 131072 "HK_FAST_BLOCK=128" "HK_FAST_BLOCK=64"
STEPS=50 scripts/ab_sweep.sh 1048576 "HK_FAST_BLOCK=128" "HK_FAST_BLOCK=64"
} > gpurun_out/ab_r1v.txt 2>&1; cat gpurun_out/ab_r1v.txt
