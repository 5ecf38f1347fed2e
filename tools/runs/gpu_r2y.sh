#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2y_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2y_probe_1rank.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2y_probe_1rank_b.log 2>&1
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2y_timeline_tf32_bf16.txt 2> gpurun_out/r2y_timeline.err
cat gpurun_out/r2y_rc.log; tail -n 3 gpurun_out/r2y_pytest_all.log; grep -h EXCHANGE_PROBE gpurun_out/r2y_probe_1rank.log gpurun_out/r2y_probe_1rank_b.log
grep -n "attention_fwd" gpurun_out/r2y_timeline_tf32_bf16.txt | head -4
