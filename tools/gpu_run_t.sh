#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_profile.py > gpurun_out/t_train_profile.log 2>&1; echo "train_profile rc=$?" | tee gpurun_out/t_rc.txt
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -s > gpurun_out/t_train_test.log 2>&1; echo "train test rc=$?" | tee -a gpurun_out/t_rc.txt
tail -n 6 gpurun_out/t_train_test.log | cut -c1-250
