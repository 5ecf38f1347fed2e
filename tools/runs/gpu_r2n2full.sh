#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2n2_bench_ref.json 2> gpurun_out/r2n2_bench_ref.err; echo "ref rc=$?" > gpurun_out/r2n2_rc.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench_full.json 2> gpurun_out/r2n2_bench_full.err; echo "full rc=$?" >> gpurun_out/r2n2_rc.log
cat gpurun_out/r2n2_rc.log; tail -n 3 gpurun_out/r2n2_bench_full.err; tail -c 400 gpurun_out/r2n2_bench_ref.json
python -c "
import json
d=json.loads(open('gpurun_out/r2n2_bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce_check'), d['gpu_launches_per_step'], d['roofline']['frac'], d.get('soak'))
"
