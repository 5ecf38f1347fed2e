#!/bin/bash
# two GPUs: peer all-reduce test (vs NCCL), peer_worker log, bench at N=2 (reference arm not needed)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q -m gpu > gpurun_out/r2m2_pytest_peer.log 2>&1; echo "pytest peer rc=$?" > gpurun_out/r2m2_rc.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/peer_worker.py > gpurun_out/r2m2_peer_worker_2rank.log 2>&1; echo "peer_worker rc=$?" >> gpurun_out/r2m2_rc.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2m2_bench_2gpu.json 2> gpurun_out/r2m2_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/r2m2_rc.log
cat gpurun_out/r2m2_rc.log; tail -n 5 gpurun_out/r2m2_pytest_peer.log; tail -n 12 gpurun_out/r2m2_peer_worker_2rank.log; tail -n 5 gpurun_out/r2m2_bench_2gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m2_bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d.get('allreduce_check'), d.get('host_binding'))
"
