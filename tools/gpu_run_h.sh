#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/h_rc.txt
timeout 300 python tools/ablation.py > gpurun_out/h_ablation.txt 2>&1; echo "ablation rc=$?" | tee -a gpurun_out/h_rc.txt
timeout 300 python tools/misc_numbers.py > gpurun_out/h_misc.json 2> gpurun_out/h_misc.err; echo "misc rc=$?" | tee -a gpurun_out/h_rc.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/h_bench_fp32.json 2> gpurun_out/h_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/h_rc.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/h_train_n1.json 2> gpurun_out/h_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/h_rc.txt
# compute-sanitizer at the small test shapes: window (fp16 / fp32), camera, GEMM, generic kernels
SEL="tests/test_gpu_window32.py tests/test_gpu_window.py::test_bev_sample_win_vs_oracle tests/test_gpu_window.py::test_img_sample_win_vs_fp32_kernel tests/test_gpu_gemm.py::test_linear_x3_gaussian tests/test_gpu_gemm.py::test_linear_x3_residual_strided_and_layernorm tests/test_gpu_gemm.py::test_linear_layernorm_exact_inputs tests/test_gpu_kernels.py"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SEL -q -x -k "not full_size and not 40000" > gpurun_out/h_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/h_rc.txt
done
tail -n 3 gpurun_out/h_all.log; cat gpurun_out/h_ablation.txt; for t in memcheck racecheck synccheck; do grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/h_sanitizer_$t.log | tail -3; done
