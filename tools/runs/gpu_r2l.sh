#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbs.py -x -q -m gpu > gpurun_out/r2l_pytest_lbs.log 2>&1; echo "pytest lbs rc=$?" > gpurun_out/r2l_rc.log
timeout 300 python tools/lbs_quick.py > gpurun_out/r2l_lbs_quick.log 2>&1
SCAT_LBS_PIPELINE=0 timeout 300 python tools/lbs_quick.py > gpurun_out/r2l_lbs_quick_nopipe.log 2>&1
cat gpurun_out/r2l_rc.log; tail -n 2 gpurun_out/r2l_pytest_lbs.log; grep -h LBS_QUICK gpurun_out/r2l_lbs_quick.log gpurun_out/r2l_lbs_quick_nopipe.log
