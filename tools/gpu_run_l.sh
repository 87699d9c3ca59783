#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k "x3" > gpurun_out/l_gemm.log 2>&1; echo "gemm rc=$?" | tee gpurun_out/l_rc.txt
UB_X3_PREFETCH=1 OPTS='1,1,2,0' timeout 240 python tools/bench_gemm_x3.py > gpurun_out/l_gemm_bench.log 2>&1
UB_X3_PREFETCH=0 OPTS='1,1,2,0' timeout 240 python tools/bench_gemm_x3.py > gpurun_out/l_gemm_bench_nopf.log 2>&1
tail -n 3 gpurun_out/l_gemm.log; cat gpurun_out/l_gemm_bench.log; echo "-- no prefetch"; cat gpurun_out/l_gemm_bench_nopf.log
