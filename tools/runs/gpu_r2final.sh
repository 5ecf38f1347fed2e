#!/bin/bash
# final 1-GPU record: tests, smoke, full bench (tf32 headline + bf16), reference arm, timeline, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2f_rc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_rc.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r2f_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2f_bench_1gpu_bf16.json 2> gpurun_out/r2f_bench_bf16.err; echo "bench bf16 rc=$?" >> gpurun_out/r2f_rc.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "bench ref rc=$?" >> gpurun_out/r2f_rc.log
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2f_timeline_tf32_bf16.txt 2> gpurun_out/r2f_timeline.err
timeout 300 python tools/lbs_quick.py > gpurun_out/r2f_lbs_quick.log 2>&1
python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2f_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches.txt 2>&1
cat gpurun_out/r2f_rc.log; tail -n 3 gpurun_out/r2f_pytest_all.log; tail -n 2 gpurun_out/r2f_smoke.log; tail -n 3 gpurun_out/r2f_bench_1gpu.err; cat gpurun_out/r2f_lbs_quick.log | tail -1
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','vs_baseline')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline'])
"
