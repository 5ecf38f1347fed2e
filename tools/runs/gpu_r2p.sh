#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2p_pytest_all.log 2>&1; echo "pytest all rc=$?" > gpurun_out/r2p_rc.log
SCAT_GEMM_NO_B_PREFETCH=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2p_probe_noprefetch.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2p_probe_prefetch.log 2>&1
SCAT_GEMM_NO_B_PREFETCH=1 timeout 300 python tools/exchange_probe.py > gpurun_out/r2p_probe_noprefetch2.log 2>&1
timeout 300 python tools/exchange_probe.py > gpurun_out/r2p_probe_prefetch2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2p_bench_quick.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?" >> gpurun_out/r2p_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick --precision bf16 > gpurun_out/r2p_bench_quick_bf16.json 2>> gpurun_out/r2p_bench.err
cat gpurun_out/r2p_rc.log; tail -n 3 gpurun_out/r2p_pytest_all.log; grep -H EXCHANGE_PROBE gpurun_out/r2p_probe_*.log
python -c "
import json
for f in ('r2p_bench_quick','r2p_bench_quick_bf16'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline']['frac'])
"
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:"attention_fwd_tc128|gemm_simt|gemm_tc|lbs_|peer_allreduce|proj_loss|regressor_" -s 0 -c 45 -o /tmp/misc python tools/prof_misc.py > gpurun_out/r2p_ncu_misc.log 2>&1
python tools/ncu_metrics.py /tmp/misc.ncu-rep > gpurun_out/r2p_misc_metrics.txt 2>&1
python tools/ncu_top_stalls.py /tmp/misc.ncu-rep 10 > gpurun_out/r2p_misc_stalls.txt 2>&1
grep -c "^==" gpurun_out/r2p_misc_metrics.txt
