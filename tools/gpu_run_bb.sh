#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window32.py tests/test_gpu_encoder.py -q -x > gpurun_out/bb_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/bb_rc.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bb_bench_nb3.json 2> gpurun_out/bb_bench_nb3.err; echo "bench nb3 rc=$?" | tee -a gpurun_out/bb_rc.txt
UB_WIN32_BUFFERS=2 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bb_bench_nb2.json 2> gpurun_out/bb_bench_nb2.err; echo "bench nb2 rc=$?" | tee -a gpurun_out/bb_rc.txt
tail -n 4 gpurun_out/bb_tests.log
python - <<'PY'
import json
for f in ('bb_bench_nb3.json','bb_bench_nb2.json'):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), round(d.get('ms_per_step',0),3), 'e2e',round(d['e2e']['value'],2), (d.get('clocks') or {}).get('sm_mhz'), {k:(round(v['avg_us'],1),round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()})
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-800:])
PY
