#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_optim.py tests/test_gpu_ops.py -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2g_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2g_bench_quick.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?" >> gpurun_out/r2g_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2g_bench_quick_bf16.json 2>> gpurun_out/r2g_bench.err; echo "bench bf16 rc=$?" >> gpurun_out/r2g_rc.log
cat gpurun_out/r2g_rc.log; tail -n 8 gpurun_out/r2g_pytest.log; tail -n 3 gpurun_out/r2g_bench.err
python -c "
import json
for f in ('r2g_bench_quick','r2g_bench_quick_bf16'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
"
