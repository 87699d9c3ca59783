#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q > gpurun_out/b_gemm.log 2>&1; echo "gemm rc=$?" | tee gpurun_out/b_rc.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -q -s > gpurun_out/b_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/b_rc.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_encoder.py > gpurun_out/b_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/b_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/b_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/b_rc.txt
CS=4,2,1 timeout 300 python tools/bench_gemm_x3.py > gpurun_out/b_gemm_bench.log 2>&1; echo "gemmbench rc=$?" | tee -a gpurun_out/b_rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/b_bench_fp32.json 2> gpurun_out/b_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/b_rc.txt
tail -n 3 gpurun_out/b_gemm.log gpurun_out/b_encoder.log gpurun_out/b_rest.log gpurun_out/b_smoke.log; cat gpurun_out/b_gemm_bench.log
