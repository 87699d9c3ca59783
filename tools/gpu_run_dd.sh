#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_profile_shapes.py > gpurun_out/dd_shapes.log 2>&1; echo "rc=$?"
cut -c1-330 gpurun_out/dd_shapes.log | tail -45
