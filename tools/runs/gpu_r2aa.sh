#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2aa_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2aa_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2aa_probe_1rank.log 2>&1
timeout 300 python tools/lbs_quick.py > gpurun_out/r2aa_lbs_quick.log 2>&1
timeout 300 python tools/config4_timeline.py tf32 > gpurun_out/r2aa_config4_timeline_tf32.txt 2>&1
cat gpurun_out/r2aa_rc.log; tail -n 8 gpurun_out/r2aa_pytest_all.log; grep -h "EXCHANGE_PROBE\|LBS_QUICK" gpurun_out/r2aa_probe_1rank.log gpurun_out/r2aa_lbs_quick.log; head -1 gpurun_out/r2aa_config4_timeline_tf32.txt
