#!/bin/bash
mkdir -p gpurun_out
for m in h3; do for cs in 2; do echo "== $m cs=$cs"; UB_X3=1,1,$cs,0 UB_TRACE_M=160000 timeout 120 python tools/trace_gemm.py 256 $cs $m 2>&1 | grep -A3 "CTA 0:" ; done; done > gpurun_out/k_trace.log 2>&1
cat gpurun_out/k_trace.log
