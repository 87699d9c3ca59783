#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_bwd_kernel -c 3 -o gpurun_out/ee_sample_bwd python tools/profile_sample_bwd.py > gpurun_out/ee_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 5 gpurun_out/ee_ncu.log; ls -la gpurun_out/ee_sample_bwd.ncu-rep
