#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_sample_bwd.py -q -x > gpurun_out/u_ops.log 2>&1; echo "ops rc=$?" | tee gpurun_out/u_rc.txt
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_encoder.py -q -x -s -k "train or module_path or gradients" > gpurun_out/u_train_test.log 2>&1; echo "train test rc=$?" | tee -a gpurun_out/u_rc.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/u_train_n1.json 2> gpurun_out/u_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/u_rc.txt
timeout 600 python tools/train_profile.py > gpurun_out/u_train_profile.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_train_ops.py -q -x -k "not 80000" > gpurun_out/u_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/u_rc.txt
tail -n 15 gpurun_out/u_ops.log | cut -c1-220; tail -n 12 gpurun_out/u_train_test.log | cut -c1-300; tail -n 3 gpurun_out/u_memcheck.log
python - <<'PY'
import json
for f in ('u_train_n1.json',):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['config']['loss'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-600:])
PY
