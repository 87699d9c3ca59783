#!/bin/bash
mkdir -p gpurun_out
UB_X3_PAIR=1 timeout 180 python -m pytest tests/test_gpu_gemm.py -q -x -k "x3" > gpurun_out/i_gemm_pair.log 2>&1; echo "gemm pair rc=$?" | tee gpurun_out/i_rc.txt
UB_X3_PAIR=1 OPTS='1,1,2,0' timeout 180 python tools/bench_gemm_x3.py > gpurun_out/i_gemm_bench_pair.log 2>&1; echo "bench pair rc=$?" | tee -a gpurun_out/i_rc.txt
OPTS='1,1,2,0' timeout 180 python tools/bench_gemm_x3.py > gpurun_out/i_gemm_bench_single.log 2>&1
tail -n 25 gpurun_out/i_gemm_pair.log; cat gpurun_out/i_gemm_bench_pair.log gpurun_out/i_gemm_bench_single.log
