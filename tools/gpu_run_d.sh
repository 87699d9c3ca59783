#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window32.py tests/test_gpu_gemm.py -q > gpurun_out/d_unit.log 2>&1; echo "unit rc=$?" | tee gpurun_out/d_rc.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -s > gpurun_out/d_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/d_rc.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_encoder.py --deselect tests/test_gpu_window32.py > gpurun_out/d_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/d_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/d_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/d_rc.txt
OPTS='1,0,1,0;0,0,1,0;0,1,1,0;0,1,1,3000;0,1,1,6000;0,1,2,0;0,1,4,0' timeout 600 python tools/bench_gemm_x3.py > gpurun_out/d_gemm_bench.log 2>&1; echo "gemmbench rc=$?" | tee -a gpurun_out/d_rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_fp32.json 2> gpurun_out/d_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/d_rc.txt
UB_WIN32=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_fp32_tile.json 2> gpurun_out/d_bench_fp32_tile.err
UB_X3=0,1,1,0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_fp32_x3opt.json 2> gpurun_out/d_bench_fp32_x3opt.err
tail -n 4 gpurun_out/d_unit.log gpurun_out/d_encoder.log gpurun_out/d_rest.log gpurun_out/d_smoke.log; cat gpurun_out/d_gemm_bench.log
