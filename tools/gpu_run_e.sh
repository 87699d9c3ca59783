#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest "tests/test_gpu_window32.py::test_img_sample_win32_vs_fp32_kernel" -x -q > gpurun_out/e_img_blocking.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_window32.py::test_img_sample_win32_vs_fp32_kernel" -x -q > gpurun_out/e_img_memcheck.log 2>&1
timeout 300 python -m pytest tests/test_gpu_window32.py -q -k "not img_sample" > gpurun_out/e_w32.log 2>&1
tail -n 30 gpurun_out/e_img_blocking.log; grep -v "^$" gpurun_out/e_img_memcheck.log | head -60; tail -n 5 gpurun_out/e_w32.log
