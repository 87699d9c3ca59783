#!/bin/bash
# round-2 GPU session A: 3xTF32 GEMM + precision classes
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/a_gemm.log 2>&1; echo "gemm rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q -s > gpurun_out/a_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_encoder.py > gpurun_out/a_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/a_bench_fp32.json 2> gpurun_out/a_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/a_rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 --precision fp16 --no-cpu-baseline > gpurun_out/a_bench_fp16.json 2> gpurun_out/a_bench_fp16.err; echo "bench16 rc=$?" | tee -a gpurun_out/a_rc.txt
tail -3 gpurun_out/a_gemm.log gpurun_out/a_encoder.log gpurun_out/a_rest.log gpurun_out/a_smoke.log
