#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sample_bwd.py tests/test_gpu_train.py -q -x > gpurun_out/ff_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/ff_rc.txt
timeout 300 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/ff_train_n1.json 2> gpurun_out/ff_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/ff_rc.txt
tail -n 4 gpurun_out/ff_tests.log | cut -c1-300
python - <<'PY'
import json
for f in ('ff_train_n1.json',):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), round(d.get('ms_per_step',0),3), 'e2e',round(d['e2e']['value'],2), d.get('gpu_launches'), d['config']['loss'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-1500:])
PY
