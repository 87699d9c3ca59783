#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:win32_kernel -c 8 -o gpurun_out/o_win32_b4 python tools/profile_window32.py 4 > gpurun_out/o_ncu_win32.log 2>&1; echo "ncu win32 rc=$?" | tee gpurun_out/o_rc.txt
ONCE=1 OPTS='1,1,2,0' timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -c 12 -o gpurun_out/o_gemm_b4 python tools/bench_gemm_x3.py > gpurun_out/o_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a gpurun_out/o_rc.txt
ls -la gpurun_out/o_*.ncu-rep
