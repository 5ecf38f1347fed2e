#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q -m gpu > gpurun_out/r2m5_pytest_peer.log 2>&1; echo "pytest peer rc=$?" > gpurun_out/r2m5_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/exchange_probe.py gpurun_out/r2m5_timeline_2rank.txt > gpurun_out/r2m5_probe_2rank.log 2>&1; echo "probe2 rc=$?" >> gpurun_out/r2m5_rc.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2m5_bench_2gpu.json 2> gpurun_out/r2m5_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/r2m5_rc.log
cat gpurun_out/r2m5_rc.log; tail -n 3 gpurun_out/r2m5_pytest_peer.log; grep -h EXCHANGE_PROBE gpurun_out/r2m5_probe_2rank.log | tail -1
grep -n "peer_allreduce\|conv_dgrad\|conv_wgrad" gpurun_out/r2m5_timeline_2rank.txt | head -6
python -c "
import json
d=json.loads(open('gpurun_out/r2m5_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_check'))
"
