#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -x -s > gpurun_out/w_train_test.log 2>&1; echo "train test rc=$?" | tee gpurun_out/w_rc.txt
timeout 600 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/w_train_n1.json 2> gpurun_out/w_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/w_rc.txt
timeout 600 python bench.py --mode train --steps 20 --warmup 3 --no-graphs > gpurun_out/w_train_n1_nograph.json 2> gpurun_out/w_train_n1_nograph.err; echo "train nograph rc=$?" | tee -a gpurun_out/w_rc.txt
grep -n "largest relative\|Error\|passed\|failed" gpurun_out/w_train_test.log | cut -c1-400 | tail -12
python - <<'PY'
import json
for f in ('w_train_n1.json','w_train_n1_nograph.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['config']['loss'], d['config'].get('cuda_graphs'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-1500:])
PY
