#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q > gpurun_out/j_gemm.log 2>&1; echo "gemm rc=$?" | tee gpurun_out/j_rc.txt
OPTS='1,1,2,0' timeout 240 python tools/bench_gemm_x3.py > gpurun_out/j_gemm_bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/j_rc.txt
UB_X3_PAIR=1 OPTS='1,1,2,0' timeout 240 python tools/bench_gemm_x3.py > gpurun_out/j_gemm_bench_pair.log 2>&1
tail -n 25 gpurun_out/j_gemm.log; cat gpurun_out/j_gemm_bench.log gpurun_out/j_gemm_bench_pair.log
