#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbs.py -x -q > gpurun_out/r2c_pytest_lbs.log 2>&1; echo "pytest lbs rc=$?" > gpurun_out/r2c_rc.log
for s in fp32 bf16; do
  timeout 180 python tools/prof_conv.py 96 tc $s > gpurun_out/r2c_conv_$s.log 2>&1; echo "prof_conv $s rc=$?" >> gpurun_out/r2c_rc.log
done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_head.py -x -q > gpurun_out/r2c_pytest_ops_head.log 2>&1; echo "pytest ops+head rc=$?" >> gpurun_out/r2c_rc.log
python tools/prof_conv.py 96 tc fp32 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_.*tc_kernel -s 6 -c 3 --csv --log-file gpurun_out/r2c_ncu_conv_fp32.csv python tools/prof_conv.py 96 tc fp32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_.*tc_kernel -s 6 -c 3 --csv --log-file gpurun_out/r2c_ncu_conv_bf16.csv python tools/prof_conv.py 96 tc bf16 > /dev/null 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_rc.log
cat gpurun_out/r2c_rc.log; tail -n 12 gpurun_out/r2c_pytest_lbs.log; tail -n 6 gpurun_out/r2c_pytest_ops_head.log
grep -h "conv_" gpurun_out/r2c_ncu_conv_fp32.csv gpurun_out/r2c_ncu_conv_bf16.csv | cut -d, -f5,13- | head -20
