#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/g_launches_fp32.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graphs > gpurun_out/g_ncu_fp32.log 2>&1; echo "ncu rc=$?" | tee gpurun_out/g_rc.txt
python tools/launch_summary.py gpurun_out/g_launches_fp32.csv 90 > gpurun_out/g_launch_summary_fp32.txt 2>&1
cat gpurun_out/g_launch_summary_fp32.txt | head -140
