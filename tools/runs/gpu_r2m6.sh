#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2m6_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2m6_rc.log
for nb in 32 64 148; do
SCAT_PEER_BLOCKS=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_probe.py gpurun_out/r2m6_timeline_2rank_nb$nb.txt > gpurun_out/r2m6_probe_2rank_nb$nb.log 2>&1; echo "probe2 nb=$nb rc=$?" >> gpurun_out/r2m6_rc.log
done
timeout 300 python tools/step_timeline.py tf32 bf16 > gpurun_out/r2m6_timeline_tf32_bf16.txt 2> gpurun_out/r2m6_timeline.err
cat gpurun_out/r2m6_rc.log; tail -n 3 gpurun_out/r2m6_pytest_all.log; grep -H EXCHANGE_PROBE gpurun_out/r2m6_probe_2rank_nb*.log
grep -n "peer_allreduce\|conv_dgrad\|conv_wgrad\|pl_loss\|conv_weight_prep" gpurun_out/r2m6_timeline_2rank_nb64.txt | head -9
