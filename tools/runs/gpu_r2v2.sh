#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2v2_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2v2_rc.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/peer_worker.py > gpurun_out/r2v2_peer_worker_2rank.log 2>&1; echo "peer_worker rc=$?" >> gpurun_out/r2v2_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_probe.py gpurun_out/r2v2_timeline_2rank.txt > gpurun_out/r2v2_probe_2rank.log 2>&1; echo "probe2 rc=$?" >> gpurun_out/r2v2_rc.log
timeout 300 python tools/exchange_probe.py > gpurun_out/r2v2_probe_1rank.log 2>&1
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2v2_timeline_tf32_bf16.txt 2> gpurun_out/r2v2_timeline.err
cat gpurun_out/r2v2_rc.log; tail -n 3 gpurun_out/r2v2_pytest_all.log; tail -n 1 gpurun_out/r2v2_peer_worker_2rank.log; grep -h EXCHANGE_PROBE gpurun_out/r2v2_probe_2rank.log gpurun_out/r2v2_probe_1rank.log | sort -u
head -4 gpurun_out/r2v2_timeline_tf32_bf16.txt | cut -c1-120
