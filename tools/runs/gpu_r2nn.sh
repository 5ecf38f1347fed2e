#!/bin/bash
# re-capture of the kernels that changed after the first round-2 ncu pass: LayerNorm forward, attention fwd/bwd, LBS skinning + blend GEMM
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_kernels.py > gpurun_out/r2nn_plain_kernels.log 2>&1 &&
timeout 600 $NCU -k regex:"attention|layernorm" -s 8 -c 4 -o /tmp/kernels2 python tools/prof_kernels.py > gpurun_out/r2nn_ncu_kernels.log 2>&1
python tools/ncu_metrics.py /tmp/kernels2.ncu-rep > gpurun_out/r2nn_kernels_metrics.txt 2>&1
python tools/prof_misc.py > gpurun_out/r2nn_plain_misc.log 2>&1 &&
timeout 900 $NCU -k regex:"lbs_tc_skin|lbs_tc_setup|gemm_tc" -s 6 -c 6 -o /tmp/misc2 python tools/prof_misc.py > gpurun_out/r2nn_ncu_misc.log 2>&1
python tools/ncu_metrics.py /tmp/misc2.ncu-rep > gpurun_out/r2nn_misc_metrics.txt 2>&1
python tools/ncu_top_stalls.py /tmp/misc2.ncu-rep 10 > gpurun_out/r2nn_misc_stalls.txt 2>&1
grep -c "^==" gpurun_out/r2nn_kernels_metrics.txt gpurun_out/r2nn_misc_metrics.txt; grep -E "^==|gpu__time_duration" gpurun_out/r2nn_kernels_metrics.txt gpurun_out/r2nn_misc_metrics.txt | cut -c1-140
