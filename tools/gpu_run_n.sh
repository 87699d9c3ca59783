#!/bin/bash
# 8-GPU box: inference (BASELINE metric) and training step (configs[4]) at N = 1 (train only), 2, 4, 8
mkdir -p gpurun_out
run() { # N mode tag extra
  local N=$1 MODE=$2 TAG=$3; shift 3
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --mode $MODE "$@" > gpurun_out/n_${TAG}_n1.json 2> gpurun_out/n_${TAG}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --mode $MODE "$@" > gpurun_out/n_${TAG}_n$N.json 2> gpurun_out/n_${TAG}_n$N.err
  fi
  echo "$TAG N=$N rc=$?" | tee -a gpurun_out/n_rc.txt
}
rm -f gpurun_out/n_rc.txt
nvidia-smi -L | wc -l
for N in 8 4 2; do run $N infer infer --steps 50 --warmup 5 --no-cpu-baseline; done
for N in 8 4 2 1; do run $N train train --steps 10 --warmup 3; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/n_*_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['e2e'].get('host_ceiling_frames_s'), d['config'].get('collective'))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
