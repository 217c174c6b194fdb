python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_r1c.txt; cat gpurun_out/pytest_r1c.txt
python scripts/lane_trace.py 65536 > gpurun_out/lane_trace_r1m.txt 2>&1
grep "solve shapes" gpurun_out/lane_trace_r1m.txt
