#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2u_rc.log
SCAT_GEMM_NO_WIDE=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2u_probe_nowide.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2u_probe_wide.log 2>&1
SCAT_GEMM_NO_WIDE=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2u_probe_nowide2.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2u_probe_wide2.log 2>&1
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2u_timeline_tf32_bf16.txt 2> gpurun_out/r2u_timeline.err
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2u_bench_quick_bf16.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?" >> gpurun_out/r2u_rc.log
cat gpurun_out/r2u_rc.log; tail -n 12 gpurun_out/r2u_pytest_all.log; grep -H "EXCHANGE_PROBE" gpurun_out/r2u_*.log
python -c "
import json
for f in ('r2u_bench_quick_bf16',):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
"
