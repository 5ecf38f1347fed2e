#!/bin/bash
# round 2, first GPU call: new conv kernels, full GPU suite, LBS blockings, gradient error report, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
for s in fp32 bf16; do
  timeout 180 python tools/prof_conv.py 96 tc $s > gpurun_out/r2a_conv_$s.log 2>&1; echo "prof_conv $s rc=$?" >> gpurun_out/r2a_rc.log
done
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "conv or proj_loss or attention_fwd_bwd" > gpurun_out/r2a_pytest_conv.log 2>&1; echo "pytest conv rc=$?" >> gpurun_out/r2a_rc.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest all rc=$?" >> gpurun_out/r2a_rc.log
timeout 600 python tools/grad_error_report.py > gpurun_out/r2a_grad_errors.log 2>&1; echo "grad report rc=$?" >> gpurun_out/r2a_rc.log
SCAT_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_lbs.py -k experimental -s -q > gpurun_out/r2a_lbs_v2.log 2>&1; echo "lbs v2 rc=$?" >> gpurun_out/r2a_rc.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_rc.log
cat gpurun_out/r2a_rc.log
tail -5 gpurun_out/r2a_conv_fp32.log gpurun_out/r2a_conv_bf16.log
tail -15 gpurun_out/r2a_pytest_conv.log
tail -15 gpurun_out/r2a_pytest_all.log
