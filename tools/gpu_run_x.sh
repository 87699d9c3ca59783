#!/bin/bash
# graphed training step on N GPUs (gradient exchange after the backward graph); every launch under a short timeout
mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --mode train --steps 20 --warmup 3 > gpurun_out/x_train_n$N.json 2> gpurun_out/x_train_n$N.err; echo "train N=$N graphs rc=$?" | tee gpurun_out/x_rc.txt
python - <<PY
import json
for f in ('x_train_n$N.json',):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['config']['params_in_sync_across_ranks'], d['config'].get('cuda_graphs'), d['config'].get('collective'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-2500:])
PY
