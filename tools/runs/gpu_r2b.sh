#!/bin/bash
# ncu --set full of the three conv_tc kernels, both seams; summaries only come back
mkdir -p gpurun_out
for s in fp32 bf16; do
  python tools/prof_conv.py 96 tc $s > gpurun_out/r2b_plain_$s.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_.*tc_kernel -s 6 -c 3 -o /tmp/prof_$s python tools/prof_conv.py 96 tc $s > gpurun_out/r2b_ncu_$s.log 2>&1
  python tools/ncu_metrics.py /tmp/prof_$s.ncu-rep > gpurun_out/r2b_conv_${s}_metrics.txt 2>&1
  python tools/ncu_top_stalls.py /tmp/prof_$s.ncu-rep 16 > gpurun_out/r2b_conv_${s}_stalls.txt 2>&1
done
ls -la /tmp/*.ncu-rep
