#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -s > gpurun_out/v_train_test.log 2>&1; echo "train test rc=$?" | tee gpurun_out/v_rc.txt
grep -n "largest relative\|AssertionError\|passed\|failed" gpurun_out/v_train_test.log | cut -c1-600
