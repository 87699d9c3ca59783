#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/p_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/p_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/p_rc.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/p_bench_fp32.json 2> gpurun_out/p_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/p_rc.txt
tail -n 5 gpurun_out/p_all.log; tail -n 4 gpurun_out/p_smoke.log
python - <<'PY'
import json
for f in ('p_bench_fp32.json',):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value'],1), d['gpu_launches'], d['clocks']['sm_mhz'], {k:v['launches_per_step'] for k,v in d['other_kernels'].items()}, {k:(round(v['avg_us'],1),round(v['frac'],3)) for k,v in d['kernels'].items()})
    except Exception as e: print(f,'ERR',e)
PY
