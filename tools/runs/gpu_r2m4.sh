#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2m4_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2m4_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_probe.py gpurun_out/r2m4_timeline_2rank.txt > gpurun_out/r2m4_probe_2rank.log 2>&1; echo "probe2 rc=$?" >> gpurun_out/r2m4_rc.log
SCAT_GEMM_SHALLOW=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2m4_probe_1rank_shallow.log 2>&1; echo "probe1 shallow rc=$?" >> gpurun_out/r2m4_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2m4_probe_1rank_deep.log 2>&1; echo "probe1 deep rc=$?" >> gpurun_out/r2m4_rc.log
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2m4_timeline_tf32_bf16.txt 2> gpurun_out/r2m4_timeline.err
cat gpurun_out/r2m4_rc.log; tail -n 4 gpurun_out/r2m4_pytest_all.log; grep -h EXCHANGE_PROBE gpurun_out/r2m4_probe_*.log; tail -n 5 gpurun_out/r2m4_probe_2rank.log
