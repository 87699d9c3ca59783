#!/bin/bash
# 8-GPU box: graphed training step (configs[4]) at N = 8, 4; every launch under a short timeout
mkdir -p gpurun_out
rm -f gpurun_out/z_rc.txt
for N in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --mode train --steps 20 --warmup 3 > gpurun_out/z_train_n$N.json 2> gpurun_out/z_train_n$N.err; echo "train N=$N rc=$?" | tee -a gpurun_out/z_rc.txt
done
timeout 240 python bench.py --gpus 1 --mode train --steps 20 --warmup 3 > gpurun_out/z_train_n1.json 2> gpurun_out/z_train_n1.err; echo "train N=1 rc=$?" | tee -a gpurun_out/z_rc.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/z_train_n*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['config']['params_in_sync_across_ranks'], d['config'].get('collective'))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
