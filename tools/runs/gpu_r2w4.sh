#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 tools/exchange_probe.py gpurun_out/r2w4_timeline_4rank.txt > gpurun_out/r2w4_probe_4rank.log 2>&1; echo "probe4 rc=$?" > gpurun_out/r2w4_rc.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 5 --quick > gpurun_out/r2w4_bench_4gpu.json 2> gpurun_out/r2w4_bench_4gpu.err; echo "bench4 rc=$?" >> gpurun_out/r2w4_rc.log
cat gpurun_out/r2w4_rc.log; grep -h EXCHANGE_PROBE gpurun_out/r2w4_probe_4rank.log | tail -1
python -c "
import json
d=json.loads(open('gpurun_out/r2w4_bench_4gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_check'), d['gpu_launches_per_step'])
"
