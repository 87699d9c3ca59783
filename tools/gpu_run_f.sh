#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window32.py tests/test_gpu_gemm.py -q > gpurun_out/f_unit.log 2>&1; echo "unit rc=$?" | tee gpurun_out/f_rc.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -s > gpurun_out/f_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/f_rc.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_encoder.py --deselect tests/test_gpu_window32.py > gpurun_out/f_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/f_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/f_rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_fp32.json 2> gpurun_out/f_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/f_rc.txt
UB_WIN32=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_fp32_tile.json 2> gpurun_out/f_bench_fp32_tile.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --precision fp16 > gpurun_out/f_bench_fp16.json 2> gpurun_out/f_bench_fp16.err; echo "bench16 rc=$?" | tee -a gpurun_out/f_rc.txt
tail -n 4 gpurun_out/f_unit.log gpurun_out/f_encoder.log gpurun_out/f_rest.log gpurun_out/f_smoke.log
