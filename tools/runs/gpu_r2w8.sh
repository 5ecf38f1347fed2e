#!/bin/bash
mkdir -p gpurun_out
for nb in 64 128 0; do
SCAT_PEER_BLOCKS=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/exchange_probe.py gpurun_out/r2w8_timeline_8rank_nb$nb.txt > gpurun_out/r2w8_probe_8rank_nb$nb.log 2>&1; echo "probe8 nb=$nb rc=$?" >> gpurun_out/r2w8_rc.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/r2w8_bench_8gpu.json 2> gpurun_out/r2w8_bench_8gpu.err; echo "bench8 rc=$?" >> gpurun_out/r2w8_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 5 --quick > gpurun_out/r2w8_bench_4gpu.json 2> gpurun_out/r2w8_bench_4gpu.err; echo "bench4 rc=$?" >> gpurun_out/r2w8_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2w8_bench_2gpu.json 2> gpurun_out/r2w8_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/r2w8_rc.log
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2w8_bench_1gpu.json 2> gpurun_out/r2w8_bench_1gpu.err; echo "bench1 rc=$?" >> gpurun_out/r2w8_rc.log
cat gpurun_out/r2w8_rc.log; grep -H EXCHANGE_PROBE gpurun_out/r2w8_probe_8rank_nb*.log | sort -u
python -c "
import json
for f in ('r2w8_bench_8gpu','r2w8_bench_4gpu','r2w8_bench_2gpu','r2w8_bench_1gpu'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_check'))
"
