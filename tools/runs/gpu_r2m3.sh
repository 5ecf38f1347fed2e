#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2m3_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2m3_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2m3_probe_1rank.log 2>&1; echo "probe1 rc=$?" >> gpurun_out/r2m3_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_probe.py gpurun_out/r2m3_timeline_2rank.txt > gpurun_out/r2m3_probe_2rank.log 2>&1; echo "probe2 rc=$?" >> gpurun_out/r2m3_rc.log
cat gpurun_out/r2m3_rc.log; tail -n 4 gpurun_out/r2m3_pytest_all.log; grep EXCHANGE_PROBE gpurun_out/r2m3_probe_1rank.log gpurun_out/r2m3_probe_2rank.log; tail -n 5 gpurun_out/r2m3_probe_2rank.log
