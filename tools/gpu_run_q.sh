#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_profile.py > gpurun_out/q_train_profile.log 2>&1; echo "train_profile rc=$?" | tee gpurun_out/q_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/q_launches_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/q_rc.txt
python tools/launch_summary.py gpurun_out/q_launches_fp32.csv 100 > gpurun_out/q_launches_summary_fp32.txt 2>&1
head -40 gpurun_out/q_launches_summary_fp32.txt
tail -n 45 gpurun_out/q_train_profile.log | cut -c1-200
