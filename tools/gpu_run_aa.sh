#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/aa_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/aa_rc.txt
tail -n 6 gpurun_out/aa_all.log
