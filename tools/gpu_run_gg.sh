#!/bin/bash
mkdir -p gpurun_out
for CS in 1 4 2; do
  UB_X3=1,1,$CS,0 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/gg_bench_cs$CS.json 2> gpurun_out/gg_bench_cs$CS.err
done
python - <<'PY'
import json
for cs in (1,4,2):
    f='gg_bench_cs%d.json'%cs
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), round(d.get('ms_per_step',0),3), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-500:])
PY
