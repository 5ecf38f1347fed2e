#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2k_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2k_bench_quick.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?" >> gpurun_out/r2k_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2k_bench_quick_bf16.json 2>> gpurun_out/r2k_bench.err
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2k_timeline_tf32_bf16.txt 2> gpurun_out/r2k_timeline.err
cat gpurun_out/r2k_rc.log; tail -n 6 gpurun_out/r2k_pytest_all.log; tail -n 3 gpurun_out/r2k_bench.err
python -c "
import json
for f in ('r2k_bench_quick','r2k_bench_quick_bf16'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
"
