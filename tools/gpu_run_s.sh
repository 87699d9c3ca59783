#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sample_bwd.py tests/test_gpu_train.py -q -x > gpurun_out/s_bwd.log 2>&1; echo "bwd rc=$?" | tee gpurun_out/s_rc.txt
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x -k "module_path or golden" > gpurun_out/s_enc.log 2>&1; echo "enc rc=$?" | tee -a gpurun_out/s_rc.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/s_train_n1.json 2> gpurun_out/s_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/s_rc.txt
UB_FUSED_TRAIN=0 timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/s_train_n1_unfused.json 2> gpurun_out/s_train_n1_unfused.err; echo "train unfused rc=$?" | tee -a gpurun_out/s_rc.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sample_bwd.py -q -x > gpurun_out/s_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/s_rc.txt
tail -n 30 gpurun_out/s_bwd.log | cut -c1-220; tail -n 8 gpurun_out/s_enc.log | cut -c1-220; tail -n 4 gpurun_out/s_memcheck.log
python - <<'PY'
import json
for f in ('s_train_n1.json','s_train_n1_unfused.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['config']['loss'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-600:])
PY
