#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2t_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2t_probe_default.log 2>&1
SCAT_EXP_SKIP_GELU=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2t_probe_skipgelu.log 2>&1
timeout 300 python tools/gemm_timeline.py > gpurun_out/r2t_gemm_timeline.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2t_bench_quick.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?" >> gpurun_out/r2t_rc.log
cat gpurun_out/r2t_rc.log; tail -n 3 gpurun_out/r2t_pytest_all.log; grep -H "EXCHANGE_PROBE" gpurun_out/r2t_*.log; cat gpurun_out/r2t_gemm_timeline.txt | cut -c1-400
python -c "
import json
for f in ('r2t_bench_quick',):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
"
