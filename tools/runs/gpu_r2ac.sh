#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ac_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2ac_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision fp32 --seam fp32 > gpurun_out/r2ac_bench_fp32.json 2> gpurun_out/r2ac_bench_fp32.err; echo "bench fp32 rc=$?" >> gpurun_out/r2ac_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2ac_probe.log 2>&1
cat gpurun_out/r2ac_rc.log; tail -n 3 gpurun_out/r2ac_pytest_all.log; grep -h EXCHANGE_PROBE gpurun_out/r2ac_probe.log
python -c "
import json
d=json.load(open('gpurun_out/r2ac_bench_fp32.json')); print('fp32', d['value'], d['ms_per_step'])
"
