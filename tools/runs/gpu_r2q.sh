#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2q_rc.log
timeout 300 python tools/lbs_quick.py > gpurun_out/r2q_lbs_quick.log 2>&1; echo "lbs quick rc=$?" >> gpurun_out/r2q_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2q_probe.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2q_bench_quick.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?" >> gpurun_out/r2q_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2q_bench_quick_bf16.json 2>> gpurun_out/r2q_bench.err
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2q_timeline_tf32_bf16.txt 2> gpurun_out/r2q_timeline.err
timeout 300 python tools/config4_timeline.py tf32 > gpurun_out/r2q_config4_timeline_tf32.txt 2>&1
cat gpurun_out/r2q_rc.log; tail -n 3 gpurun_out/r2q_pytest_all.log; grep -h "LBS_QUICK\|EXCHANGE_PROBE" gpurun_out/r2q_lbs_quick.log gpurun_out/r2q_probe.log; head -2 gpurun_out/r2q_config4_timeline_tf32.txt
python -c "
import json
for f in ('r2q_bench_quick','r2q_bench_quick_bf16'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
"
