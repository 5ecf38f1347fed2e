#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2r_rc.log
timeout 300 python tools/lbs_quick.py > gpurun_out/r2r_lbs_quick.log 2>&1; echo "lbs quick rc=$?" >> gpurun_out/r2r_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2r_probe_default.log 2>&1
SCAT_CARVEOUT=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2r_probe_carveout.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2r_probe_default2.log 2>&1
SCAT_CARVEOUT=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2r_probe_carveout2.log 2>&1
SCAT_CARVEOUT=1 timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2r_timeline_carveout.txt 2> gpurun_out/r2r_timeline.err
SCAT_CARVEOUT=1 timeout 300 python tools/lbs_quick.py > gpurun_out/r2r_lbs_quick_carveout.log 2>&1
cat gpurun_out/r2r_rc.log; tail -n 3 gpurun_out/r2r_pytest_all.log; grep -H "LBS_QUICK\|EXCHANGE_PROBE" gpurun_out/r2r_*.log
