#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ad_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2ad_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2ad_bench_bf16.json 2> gpurun_out/r2ad_bench_bf16.err; echo "bench bf16 rc=$?" >> gpurun_out/r2ad_rc.log
cat gpurun_out/r2ad_rc.log; tail -n 3 gpurun_out/r2ad_pytest_all.log
python -c "
import json
d=json.load(open('gpurun_out/r2ad_bench_bf16.json')); print('bf16', d['value'], d['ms_per_step'])
"
