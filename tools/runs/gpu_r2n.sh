#!/bin/bash
# round-2 ncu evidence: --set full captures (summaries come back, the .ncu-rep files stay on the box) + the launch list
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for s in fp32 bf16; do
  python tools/prof_conv.py 96 tc $s > gpurun_out/r2n_plain_conv_$s.log 2>&1 &&
  timeout 600 $NCU -k regex:conv_.*tc_kernel -s 6 -c 3 -o /tmp/conv_$s python tools/prof_conv.py 96 tc $s > gpurun_out/r2n_ncu_conv_$s.log 2>&1
  python tools/ncu_metrics.py /tmp/conv_$s.ncu-rep > gpurun_out/r2n_conv_${s}_metrics.txt 2>&1
  python tools/ncu_top_stalls.py /tmp/conv_$s.ncu-rep 12 > gpurun_out/r2n_conv_${s}_stalls.txt 2>&1
done
python tools/prof_kernels.py > gpurun_out/r2n_plain_kernels.log 2>&1 &&
timeout 600 $NCU -k regex:"gemm_tc|attention|layernorm" -s 14 -c 7 -o /tmp/kernels python tools/prof_kernels.py > gpurun_out/r2n_ncu_kernels.log 2>&1
python tools/ncu_metrics.py /tmp/kernels.ncu-rep > gpurun_out/r2n_kernels_metrics.txt 2>&1
python tools/prof_misc.py > gpurun_out/r2n_plain_misc.log 2>&1 &&
timeout 900 $NCU -k regex:scat -s 0 -c 40 -o /tmp/misc python tools/prof_misc.py > gpurun_out/r2n_ncu_misc.log 2>&1
python tools/ncu_metrics.py /tmp/misc.ncu-rep > gpurun_out/r2n_misc_metrics.txt 2>&1
python tools/ncu_top_stalls.py /tmp/misc.ncu-rep 10 > gpurun_out/r2n_misc_stalls.txt 2>&1
python bench.py --steps 2 --warmup 3 --quick > /dev/null 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2n_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2n_launches.csv > gpurun_out/r2n_launches.txt 2>&1
ls -la /tmp/*.ncu-rep; tail -n 3 gpurun_out/r2n_plain_*.log; grep -c "^==" gpurun_out/r2n_*_metrics.txt; head -30 gpurun_out/r2n_launches.txt
timeout 300 python tools/config4_timeline.py tf32 > gpurun_out/r2n_config4_timeline_tf32.txt 2>&1
timeout 300 python tools/config4_timeline.py bf16 > gpurun_out/r2n_config4_timeline_bf16.txt 2>&1
head -3 gpurun_out/r2n_config4_timeline_tf32.txt
