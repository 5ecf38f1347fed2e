#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2x_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2x_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2x_probe_1rank.log 2>&1
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2x_timeline_tf32_bf16.txt 2> gpurun_out/r2x_timeline.err
cat gpurun_out/r2x_rc.log; tail -n 3 gpurun_out/r2x_pytest_all.log; grep -h EXCHANGE_PROBE gpurun_out/r2x_probe_1rank.log
grep -n "conv_fwd\|conv_weight_prep\|round_copy\|regressor_hoist" gpurun_out/r2x_timeline_tf32_bf16.txt | head -8
