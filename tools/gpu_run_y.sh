#!/bin/bash
# final single-GPU evidence of the round
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/y_all.log 2>&1; echo "all rc=$?" | tee gpurun_out/y_rc.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/y_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 600 python bench.py > gpurun_out/y_bench_fp32.json 2> gpurun_out/y_bench_fp32.err; echo "bench32 rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 300 python bench.py --batch 1 --no-cpu-baseline > gpurun_out/y_bench_fp32_b1.json 2> gpurun_out/y_bench_fp32_b1.err; echo "bench32 b1 rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 300 python bench.py --precision fp16 --no-cpu-baseline > gpurun_out/y_bench_fp16.json 2> gpurun_out/y_bench_fp16.err; echo "bench16 rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 300 python bench.py --mode train --steps 20 --warmup 3 > gpurun_out/y_train_n1.json 2> gpurun_out/y_train_n1.err; echo "train rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/y_reference.json 2> gpurun_out/y_reference.err; echo "reference rc=$?" | tee -a gpurun_out/y_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/y_launches_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/y_ncu_bench.log 2>&1; echo "ncu list rc=$?" | tee -a gpurun_out/y_rc.txt
python tools/launch_summary.py gpurun_out/y_launches_fp32.csv 100 > gpurun_out/y_launches_summary_fp32.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:win32_kernel -c 8 -o gpurun_out/y_win32_b4 python tools/profile_window32.py 4 > gpurun_out/y_ncu_win32.log 2>&1; echo "ncu win32 rc=$?" | tee -a gpurun_out/y_rc.txt
tail -n 4 gpurun_out/y_all.log; tail -n 5 gpurun_out/y_smoke.log
python - <<'PY'
import json
for f in ('y_bench_fp32.json','y_bench_fp32_b1.json','y_bench_fp16.json','y_train_n1.json','y_reference.json'):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value'],2), round(d.get('ms_per_step',0),3), 'e2e',round(d['e2e']['value'],2), d.get('gpu_launches'), (d.get('clocks') or {}).get('sm_mhz'), {k:(round(v['avg_us'],1),round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()}, (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e, open('gpurun_out/'+f.replace('.json','.err')).read()[-800:])
PY
head -16 gpurun_out/y_launches_summary_fp32.txt
