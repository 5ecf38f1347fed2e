#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2z_rc.log
for w in 0 1 2 4 7; do
SCAT_EXP_WIDE_BWD=$w timeout 300 python tools/exchange_probe.py > gpurun_out/r2z_probe_wide$w.log 2>&1
done
timeout 300 python tools/exchange_probe.py > gpurun_out/r2z_probe_wide0b.log 2>&1
cat gpurun_out/r2z_rc.log; tail -n 3 gpurun_out/r2z_pytest_all.log; grep -H EXCHANGE_PROBE gpurun_out/r2z_probe_wide*.log
