#!/bin/bash
# 8 GPUs of one node: peer exchange vs NCCL (peer_worker), the exchange probe, bench at N = 8 (and the reference-free N = 4)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s8_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tests/peer_worker.py > gpurun_out/r2s8_peer_worker_8rank.log 2>&1; echo "peer_worker8 rc=$?" > gpurun_out/r2s8_rc.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/exchange_probe.py gpurun_out/r2s8_timeline_8rank.txt > gpurun_out/r2s8_probe_8rank.log 2>&1; echo "probe8 rc=$?" >> gpurun_out/r2s8_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 20 --warmup 5 --quick > gpurun_out/r2s8_bench_8gpu.json 2> gpurun_out/r2s8_bench_8gpu.err; echo "bench8 rc=$?" >> gpurun_out/r2s8_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 5 --quick > gpurun_out/r2s8_bench_4gpu.json 2> gpurun_out/r2s8_bench_4gpu.err; echo "bench4 rc=$?" >> gpurun_out/r2s8_rc.log
cat gpurun_out/r2s8_rc.log; tail -n 2 gpurun_out/r2s8_peer_worker_8rank.log; grep -h EXCHANGE_PROBE gpurun_out/r2s8_probe_8rank.log | tail -1; tail -n 3 gpurun_out/r2s8_bench_8gpu.err
python -c "
import json
for f in ('r2s8_bench_8gpu','r2s8_bench_4gpu'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_check'))
"
