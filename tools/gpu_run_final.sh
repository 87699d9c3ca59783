#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/final_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final_rc.txt
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" | tee -a gpurun_out/final_rc.txt
tail -n 3 gpurun_out/final_all.log; tail -n 5 gpurun_out/final_smoke.log
python - <<'PY'
import json
for f in ('final_bench.json',):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), round(d.get('ms_per_step',0),3), 'e2e',round(d['e2e']['value'],2), d.get('gpu_launches'), (d.get('clocks') or {}), {k:(round(v['avg_us'],1),round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()}, (d.get('cpu_baseline') or {}).get('value'), d['roofline'])
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-800:])
PY
