#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gemm.py -q -x -k "f16x3" > gpurun_out/j_gemm_f16x3.log 2>&1; echo "f16x3 rc=$?" | tee gpurun_out/j_rc.txt
OPTS='1,1,2,0;1,1,1,0' timeout 240 python tools/bench_gemm_x3.py > gpurun_out/j_gemm_bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/j_rc.txt
tail -n 25 gpurun_out/j_gemm_f16x3.log; cat gpurun_out/j_gemm_bench.log
