#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window32.py tests/test_gpu_gemm.py -q > gpurun_out/c_unit.log 2>&1; echo "unit rc=$?" | tee gpurun_out/c_rc.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -s > gpurun_out/c_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/c_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/c_rc.txt
for m in x3 x3ln; do for cs in 1 4; do echo "== $m cs=$cs"; UB_TRACE_M=160000 timeout 120 python tools/trace_gemm.py 256 $cs $m; done; done > gpurun_out/c_trace.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_fp32.json 2> gpurun_out/c_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/c_rc.txt
UB_WIN32=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_fp32_tile.json 2> gpurun_out/c_bench_fp32_tile.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -c 6 -o gpurun_out/c_gemm_x3 env ONCE=1 CS=1 python tools/bench_gemm_x3.py > gpurun_out/c_ncu.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/c_rc.txt
tail -n 4 gpurun_out/c_unit.log gpurun_out/c_encoder.log gpurun_out/c_smoke.log
