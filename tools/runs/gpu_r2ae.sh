#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2ae_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2ae_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2ae_probe.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2ae_probe2.log 2>&1
timeout 300 python tools/grad_error_report.py > gpurun_out/r2ae_grad_errors.log 2>&1
cat gpurun_out/r2ae_rc.log; tail -n 3 gpurun_out/r2ae_pytest_all.log; grep -h EXCHANGE_PROBE gpurun_out/r2ae_probe.log gpurun_out/r2ae_probe2.log; grep -A3 "precision tf32" gpurun_out/r2ae_grad_errors.log | head -8
