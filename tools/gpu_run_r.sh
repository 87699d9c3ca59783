#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/r_rc.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r_bench_fp32.json 2> gpurun_out/r_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/r_rc.txt
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r_train_n1.json 2> gpurun_out/r_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/r_rc.txt
tail -n 5 gpurun_out/r_all.log
python - <<'PY'
import json
for f in ('r_bench_fp32.json','r_train_n1.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['clocks']['sm_mhz'], {k:v['launches_per_step'] for k,v in (d.get('other_kernels') or {}).items()}, {k:(round(v['avg_us'],1),round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f,'ERR',e)
PY
